"""CPU restatements of the reference's algorithms for the hot path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs as the checker. Nothing under evfly_b200/ imports it and no
product path may route through it.
"""
