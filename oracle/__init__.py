"""CPU oracle: restatement of the reference's hot path (TEST INFRASTRUCTURE ONLY).

PARITY UNPINNED — the reference ships no tests/golden vectors for this path and
its arithmetic lives in TensorFlow 1.14, which is not installable here; see the
header of oracle/knn_oracle.c and DESIGN.md.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this package.
"""
