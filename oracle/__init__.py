"""CPU oracle for the DTW hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  See oracle/apd_oracle.h for the parity status
("parity unpinned" by the reference: it ships no tests and cannot be built here).
"""
