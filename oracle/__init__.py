"""TEST INFRASTRUCTURE ONLY -- the CPU oracle for the AIT detection-head hot path.

Nothing under `ait_b200/` may import this package.  Allowed importers: `tests/`,
`__graft_entry__.smoke()` (as the checker) and `bench.py`'s CPU-baseline / `--impl reference`
legs (as the thing timed on the host cores, never as the product).
"""
