"""ORACLE -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.

CPU restatement of the sparse U-ResNet hot path (the scn module surface used by
/root/reference/uresnet/models/uresnet_sparse.py:19-24).  See oracle/README.md.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.
"""
