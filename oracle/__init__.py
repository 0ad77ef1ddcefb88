"""CPU parity oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (rs-vio_b200/) never does.
PARITY UNPINNED -- see oracle/rsvio_oracle.h.
"""
