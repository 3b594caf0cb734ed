"""TEST INFRASTRUCTURE ONLY.

CPU restatements ("oracles") of the reference algorithms on the AV-HuBERT hot path.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  The product (multimodalvc_b200/) never does.
"""
