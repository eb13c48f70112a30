"""ORACLE package: CPU restatements of the reference algorithms, used ONLY as the checker by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
