"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (see the header of turbomesh_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this package.
"""
