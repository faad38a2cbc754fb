"""ORACLE — CPU restatement of the reference algorithms on the scoring path.

This package is TEST INFRASTRUCTURE.  It is imported only by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs, and only as the checker or the timed CPU baseline.
Nothing under ``facet_b200/`` imports it; the product path fails loudly when
the CUDA library is missing instead of falling back to this code.
"""
