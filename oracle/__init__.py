"""ORACLE — test infrastructure only (CPU restatements of the reference's hot path).

Allowed importers: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / `--impl reference` legs. The product package
`magcache_b200` must never import from here (tests/test_layout.py enforces it).
"""
