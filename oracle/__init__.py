"""Oracle package: CPU checkers for the Hamming matching hot path.

TEST INFRASTRUCTURE ONLY.  Importable from ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline legs; never from ``slam_experiments_b200``.
"""
