"""CPU oracle for the batched 2048 step path — TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  See g2048_oracle.c for the parity status.
"""
