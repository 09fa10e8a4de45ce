"""CPU oracle for the box-level hot path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()`, `bench.py`'s CPU-baseline leg and the
golden-vector generator may import anything from this package; the product
package (`rodet_b200`) never does and fails loudly when its CUDA library is
missing.  See `oracle/README.md`.
"""
