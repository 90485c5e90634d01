"""SASS evidence for the tcgen05 / TMA / bulk-copy claims: per kernel of libtt_b200.so, the number of tensor-core MMA (UTCHMMA),
tensor-memory load / store (LDTM / STTM), tcgen05 commit (UTCBAR), bulk-copy (UBLKCP), cp.async (LDGSTS) and FP64 FMA (DFMA)
instructions.    python profiles/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "ddpg-trucktrailer_b200", "libtt_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
ops = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "LDGSTS", "SYNCS", "DFMA", "FFMA2", "F2FP")
kern, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(anonymous namespace\)::", "", kern).split("(")[0]
        counts[kern] = collections.Counter()
        continue
    if kern:
        for op in ops:
            if re.search(r"\b" + op + r"\b", line) or (op + ".") in line:
                counts[kern][op] += 1
                break
print(f"{'kernel':70s} " + " ".join(f"{o:>8s}" for o in ops))
for k, c in counts.items():
    print(f"{k[:70]:70s} " + " ".join(f"{c[o]:8d}" for o in ops))
print(f"\n{len(counts)} kernel entries in {os.path.basename(lib)}")
