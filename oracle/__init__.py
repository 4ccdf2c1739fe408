"""TEST INFRASTRUCTURE.  CPU oracle for the selective-recompute search path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package; the product (leann_b200/) never does.
"""
