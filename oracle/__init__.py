"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the B200 truck-trailer rollout kernels.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  The product package never does.
"""
