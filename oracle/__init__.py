"""TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's hot path (zrc0622/HIRL4UCAV).  Nothing in
``hirl4ucav_b200`` may import from this package; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs do.  See ``oracle/README.md``.
"""
