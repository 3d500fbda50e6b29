"""Test infrastructure: CPU oracle for the mimpy MFD hot path (see mimpy_oracle.py header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package. The product (mimpy_b200/) never does.
"""
