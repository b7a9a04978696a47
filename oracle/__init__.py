"""TEST INFRASTRUCTURE -- the CPU oracle for the phyNGSC hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this package; the product (phyngsc_b200) never does.
"""
