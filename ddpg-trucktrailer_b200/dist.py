"""Multi-GPU plumbing: one process per GPU, environments sharded by global id, no data-path collective.

The path shards trivially: rank r owns global envs [r*n, (r+1)*n); the Philox streams
are keyed by (seed, global id), so a given environment's trajectory is the same for any GPU count.  The
only collectives are latency-bound and go through ``torch.distributed`` (NCCL over NVLink on the GPU box,
gloo in the CPU tests): a sum all-reduce of the 16-word statistics vector per iteration and a broadcast of
the 131 601 actor parameters from the learner rank after each update.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from ._lib import STAT_NAMES, TT_NSTATS

ACTOR_KEYS = ("fc1.weight", "fc1.bias", "bn1.weight", "bn1.bias", "fc2.weight", "fc2.bias", "bn2.weight", "bn2.bias",
              "mu.weight", "mu.bias")


def init_from_env(backend: str | None = None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns (rank, world,
    local_rank).  Single-process runs (no WORLD_SIZE) return (0, 1, 0) without creating a group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def shard(num_envs_total: int, rank: int, world: int):
    """Global env range of `rank`: (offset, count); remainders go to the lowest ranks."""
    base, rem = divmod(int(num_envs_total), world)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def all_reduce_stats(stats: torch.Tensor) -> torch.Tensor:
    """Sum the per-rank statistics vector (float64[16]) over ranks, in place."""
    assert stats.numel() == TT_NSTATS
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def summarize(stats: torch.Tensor) -> dict:
    v = stats.detach().cpu().tolist()
    d = dict(zip(STAT_NAMES, v))
    ep = max(d["episodes"], 1.0)
    d["mean_return"] = d["return_sum"] / ep
    d["std_return"] = max(d["return_sq_sum"] / ep - d["mean_return"] ** 2, 0.0) ** 0.5
    d["success_rate"] = d["successes"] / ep
    return d


def flatten_actor(sd: dict, device=None) -> torch.Tensor:
    return torch.cat([torch.as_tensor(sd[k], dtype=torch.float32, device=device).reshape(-1) for k in ACTOR_KEYS])


def unflatten_actor(flat: torch.Tensor, like: dict) -> dict:
    out, o = {}, 0
    for k in ACTOR_KEYS:
        n = like[k].numel()
        out[k] = flat[o:o + n].reshape(like[k].shape)
        o += n
    return out


def broadcast_actor(sd: dict, src: int = 0, device=None) -> dict:
    """Broadcast an actor state_dict from `src` as ONE flat float32 message (526 404 B)."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return sd
    flat = flatten_actor(sd, device=device)
    dist.broadcast(flat, src=src)
    return unflatten_actor(flat, sd)


class OverlappedSync:
    """The two per-iteration collectives of the rollout, issued asynchronously on a side stream so
    that they overlap the next iteration's kernels instead of sitting on the rollout stream:

    * ``push_stats(stats)``: sum all-reduce of the 16-double statistics vector.  Returns the all-reduced vector of the
      PREVIOUS push (None the first time): the consumer is one iteration behind, the rollout stream never waits for the
      network.
    * ``push_policy(flat)`` / ``wait_policy()``: one-message broadcast of the flat actor parameter vector from the learner
      rank (the vector ``tt_learn_step`` updates in place: no torch.cat / unflatten round trip); ``wait_policy`` makes the
      CALLING stream wait for it -- call it on the stream that re-packs the parameters into the spare packed actor.

    Works with NCCL (side CUDA stream) and with gloo (CPU tests: the same call sequence, ``wait`` blocks the host)."""

    def __init__(self, device=None, src: int = 0):
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.cuda = self.device.type == "cuda"
        self.on = dist.is_initialized() and dist.get_world_size() > 1
        self.src = src
        self._stats = [torch.zeros(TT_NSTATS, dtype=torch.float64, device=self.device) for _ in range(2)]
        self._swork = [None, None]
        self._i = 0
        self._pwork = None
        if self.cuda:
            self.stream = torch.cuda.Stream(device=self.device, priority=-1)
            self._ev = [torch.cuda.Event() for _ in range(2)]
            self._pev = torch.cuda.Event()

    def _side(self):
        import contextlib
        return torch.cuda.stream(self.stream) if self.cuda else contextlib.nullcontext()

    def push_stats(self, stats: torch.Tensor):
        b = self._i & 1
        self._i += 1
        prev = None
        if self._swork[b ^ 1] is not None:
            w = self._swork[b ^ 1]
            if w is not True:
                w.wait()                       # NCCL: the current stream waits (the collective finished an iteration ago)
            prev = self._stats[b ^ 1]
        self._stats[b].copy_(stats)
        if not self.on:
            self._swork[b] = True
            return prev
        if self.cuda:
            self._ev[b].record()
            with self._side():
                self.stream.wait_event(self._ev[b])
                self._swork[b] = dist.all_reduce(self._stats[b], op=dist.ReduceOp.SUM, async_op=True)
        else:
            self._swork[b] = dist.all_reduce(self._stats[b], op=dist.ReduceOp.SUM, async_op=True)
        return prev

    def flush_stats(self):
        """The all-reduced vector of the LAST push (waits for it)."""
        b = (self._i - 1) & 1
        w = self._swork[b]
        if w is None:
            return None
        if w is not True:
            w.wait()
        return self._stats[b]

    def push_policy(self, flat: torch.Tensor):
        """Start the broadcast of ``flat`` (in place) from the learner rank; everything queued on the current stream so far
        (the learner step that produced it) is ordered before it."""
        if not self.on:
            self._pwork = True
            return
        if self.cuda:
            self._pev.record()
            with self._side():
                self.stream.wait_event(self._pev)
                self._pwork = dist.broadcast(flat, src=self.src, async_op=True)
        else:
            self._pwork = dist.broadcast(flat, src=self.src, async_op=True)

    def wait_policy(self):
        if self._pwork is not None and self._pwork is not True:
            self._pwork.wait()
        self._pwork = None
