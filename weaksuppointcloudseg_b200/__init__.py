"""B200-native (sm_100a) hot path for weakly-supervised point-cloud segmentation.

Drop-in for the DGCNN EdgeConv stack and the weak-supervision losses of
alex-xun-xu/WeakSupPointCloudSeg.  Host side is Python/PyTorch (device memory,
streams, torch.distributed); all arithmetic runs in hand-written CUDA kernels
behind the C ABI in include/wspc.h (libwspc.so, built for sm_100a only).
"""
__version__ = "0.1.0"
