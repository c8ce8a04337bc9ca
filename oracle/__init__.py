"""TEST INFRASTRUCTURE ONLY.  Nothing under `oracle/` is product code: only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s CPU-baseline / `--impl reference` legs may import it.  parity unpinned: the reference ships no
golden vectors; the goldens in `tests/golden/` were produced here by the reference's own forward
(`oracle/make_goldens.py`)."""
