"""TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's algorithm for the pose-generation hot path.  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import anything from this package; genpose2_b200/ never does.
"""
