"""Empty stand-in for matplotlib (absent from this image): simv2.py imports it at module level but uses it only in
render() (simv2.py:376-457, 547-605), which the rollout loop never calls.  Baseline harness only."""
