"""Shared by tests/golden/make_reference_golden.py (generator) and the tests that read its fixtures: seeded
weights in the reference's variable naming and the fixed subsampling applied to large gradient tensors (keeps the
committed fixtures small).  numpy only."""
import numpy as np

S3DIS_LAYERS = [("adj_conv1", 18, 64, True), ("adj_conv2", 64, 64, True), ("adj_conv3", 128, 64, True),
                ("adj_conv4", 64, 64, True), ("adj_conv5", 128, 64, True), ("adj_conv7", 192, 1024, True),
                ("seg/conv1", 1216, 512, True), ("seg/conv2", 512, 256, True), ("seg/conv3", 256, 13, False)]
SHAPENET_LAYERS = [("transform_net1/tconv1", 6, 64, True), ("transform_net1/tconv2", 64, 128, True),
                   ("transform_net1/tconv3", 128, 1024, True), ("transform_net1/tfc1", 1024, 512, True),
                   ("transform_net1/tfc2", 512, 256, True),
                   ("adj_conv1", 6, 64, True), ("adj_conv2", 64, 64, True), ("adj_conv3", 128, 64, True),
                   ("adj_conv4", 64, 64, True), ("adj_conv5", 128, 64, True), ("adj_conv7", 192, 1024, True),
                   ("one_hot_label_expand", 16, 64, True), ("seg/conv1", 1280, 256, True), ("seg/conv2", 256, 256, True),
                   ("seg/conv3", 256, 128, True), ("seg/conv4", 128, 50, False)]
# Networks/dgcnn/models/dgcnn.py:20-98 (classification net, non-dist batch norm: pop_* = EMA shadows)
CLS_LAYERS = [("transform_net1/tconv1", 6, 64, True), ("transform_net1/tconv2", 64, 128, True),
              ("transform_net1/tconv3", 128, 1024, True), ("transform_net1/tfc1", 1024, 512, True),
              ("transform_net1/tfc2", 512, 256, True), ("dgcnn1", 6, 64, True), ("dgcnn2", 128, 64, True),
              ("dgcnn3", 128, 64, True), ("dgcnn4", 128, 128, True), ("agg", 320, 1024, True), ("fc1", 1024, 512, True),
              ("fc2", 512, 256, True), ("fc3", 256, 40, False)]
MAX_KEEP = 8192          # gradient / weight tensors above this size are stored as a strided subsample


def xavier_params(layers, seed, tnet_seed=None):
    """Xavier-uniform weights (tf_util.py:43-47) with NON-trivial biases / BN parameters / population statistics so
    every term of the graph is exercised; numpy, so reference-on-shim, oracle and CUDA engine load identical bits."""
    from collections import OrderedDict
    rng = np.random.default_rng(seed)
    p = OrderedDict()
    for scope, cin, cout, has_bn in layers:
        lim = np.sqrt(6.0 / (cin + cout))
        p[scope + "/weights"] = rng.uniform(-lim, lim, (cin, cout)).astype(np.float32)
        p[scope + "/biases"] = rng.normal(0, 0.02, (cout,)).astype(np.float32)
        if has_bn:
            p[scope + "/bn/beta"] = rng.normal(0, 0.05, (cout,)).astype(np.float32)
            p[scope + "/bn/gamma"] = (1 + rng.normal(0, 0.05, (cout,))).astype(np.float32)
            p[scope + "/bn/pop_mean"] = rng.normal(0, 0.1, (cout,)).astype(np.float32)
            p[scope + "/bn/pop_var"] = rng.uniform(0.5, 1.5, (cout,)).astype(np.float32)
    if tnet_seed is not None:      # transform_nets.py:42-55 starts at W=0, b=0 (+I); use a non-identity transform
        r2 = np.random.default_rng(tnet_seed)
        p["transform_net1/transform_XYZ/weights"] = r2.normal(0, 0.02, (256, 9)).astype(np.float32)
        p["transform_net1/transform_XYZ/biases"] = r2.normal(0, 0.05, (9,)).astype(np.float32)
    return p


def subsample(a):
    """(values, stride): the flattened tensor itself, or every stride-th element when it is larger than MAX_KEEP."""
    f = np.asarray(a).reshape(-1)
    stride = max(1, -(-f.size // MAX_KEEP))
    return f[::stride].copy(), stride
