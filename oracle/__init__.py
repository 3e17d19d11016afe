"""ORACLE — test infrastructure only (see oracle/README.md).

CPU restatements of the reference algorithms for the training-step hot path. Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package;
the product path (torchseg_b200/) never does.
"""
