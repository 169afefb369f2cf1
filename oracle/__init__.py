"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU checkers for the hot path: ``oracle.port`` (plain-C restatement,
oracle/port/*.c -> oracle/libnboracle.so) and ``oracle.ref`` (the unmodified
reference compiled from /root/reference into oracle/_ref/libnbots_ref.so).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this package; nbots_b200/ never does.
"""
