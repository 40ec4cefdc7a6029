"""CPU oracle: restatement of the reference's hot path (TEST INFRASTRUCTURE ONLY).

PINNING STATUS -- the reference ships no tests / golden vectors for this path and its arithmetic lives in
TensorFlow 1.14, which is not installable here.  What pins this oracle:
  * tests/golden/ref_*.npz -- outputs of the REFERENCE'S OWN PYTHON (tf_util, DGCNN_S3DIS, DGCNN_ShapeNet,
    transform_nets, SmoothConstraint, Tool, ProbLabelPropagation, the trainers' defineNetwork/WeakSupLoss), imported
    unmodified from /root/reference and executed on an eager op-level TF-1.14 stand-in
    (tests/golden/tf1_shim, generator tests/golden/make_reference_golden.py); tests/test_reference_golden_cpu.py holds
    the oracle to them (logits / losses / gradients / BN statistics / Adam / label propagation);
  * NOT pinned: the bit-level arithmetic of TensorFlow's binary kernels (matmul summation order, TopK tie order,
    MaxPoolGrad / reduce_max tie rules) -- those follow the published semantics listed in SURVEY.md App. A.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
"""
