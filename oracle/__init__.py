"""oracle/ -- TEST INFRASTRUCTURE (CPU restatement of the reference algorithm).

Nothing under cmflow_b200/ may import this package.  Allowed importers: tests/,
__graft_entry__.smoke(), and bench.py's cpu_baseline / --impl reference legs.
"""
