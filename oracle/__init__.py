"""CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product package (eth-lc-plonky2_b200/) never does.  PARITY UNPINNED except the
Poseidon permutation (see oracle/gl.h).
"""
