"""CPU oracle for the dense-simplex hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package (see oracle/simplex_oracle.c for the contract).
"""
