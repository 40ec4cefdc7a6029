"""Parity of the fused CUDA kNN (through the C ABI) against the CPU oracle: bit-exact indices.

Reference call sites: tf_util.pairwise_distance/knn (Networks/dgcnn/utils/tf_util.py:638-671),
SmoothConstraint Dmat/top_k (Util/SmoothConstraint.py:141-154).
"""
import numpy as np
import pytest
import torch

from oracle import knn as ok

pytestmark = pytest.mark.gpu


def _cloud(rng, B, N, D, dup_frac=0.05, kind="uniform"):
    if kind == "uniform":
        x = rng.uniform(-1, 1, (B, N, D)).astype(np.float32)
    elif kind == "relu":  # feature-like: non-negative, clustered
        x = np.maximum(rng.standard_normal((B, N, D)), 0).astype(np.float32)
    else:  # coarse grid -> massive exact ties
        x = rng.integers(0, 4, (B, N, D)).astype(np.float32) * 0.25
    ndup = int(N * dup_frac)
    if ndup:
        for b in range(B):
            src = rng.integers(0, N, ndup)
            dst = rng.integers(0, N, ndup)
            x[b, dst] = x[b, src]
    return x


CASES = [
    # B, N, D, k, flavour, kind
    (2, 128, 3, 20, 0, "uniform"),
    (3, 333, 3, 20, 0, "uniform"),      # ragged N (not a tile multiple)
    (2, 1024, 3, 20, 0, "uniform"),
    (2, 1024, 6, 10, 1, "uniform"),     # smooth flavour (clamped), k=10
    (2, 1024, 64, 20, 0, "relu"),
    (1, 777, 64, 20, 0, "relu"),
    (1, 512, 128, 20, 0, "relu"),
    (1, 600, 9, 40, 0, "uniform"),      # k > 32 -> two list slots
    (2, 512, 3, 20, 0, "grid"),         # many exact ties -> lower index first
    (2, 512, 6, 10, 1, "grid"),
    (1, 64, 3, 64, 0, "uniform"),       # k == N
    (1, 20, 3, 20, 0, "uniform"),       # N < tile, k == N
    (4, 2048, 3, 20, 0, "uniform"),
    (1, 4096, 64, 20, 0, "relu"),       # one full-size S3DIS cloud
    # 24 < k <= 64 on the tensor-core kernel: bisection threshold, three candidates per lane, rank-ordered output
    (2, 1024, 3, 25, 0, "uniform"),
    (2, 1024, 64, 32, 0, "relu"),
    (2, 777, 6, 33, 1, "uniform"),      # ragged N, clamped flavour
    (2, 1024, 64, 40, 0, "relu"),       # cfg-4's k
    (1, 2048, 3, 64, 0, "uniform"),     # the largest k of the kernel
    (2, 512, 3, 40, 0, "grid"),         # massive exact ties: rows overflow the candidate slots -> k > 32 fallback kernel
    (1, 96, 3, 64, 0, "uniform"),       # N < tile
]


@pytest.mark.parametrize("B,N,D,k,flavour,kind", CASES)
def test_knn_fused_bit_exact(cuda, B, N, D, k, flavour, kind):
    from weaksuppointcloudseg_b200 import ops

    rng = np.random.default_rng(1234 + N + D)
    x = _cloud(rng, B, N, D, kind=kind)
    ref_idx, ref_d = ok.knn(x, k, flavour, return_dist=True)
    xd = torch.from_numpy(x).to(cuda)
    idx, dist = ops.knn_fused(xd, k, flavour, return_dist=True)
    torch.cuda.synchronize()
    assert np.array_equal(idx.cpu().numpy(), ref_idx)
    assert np.array_equal(dist.cpu().numpy(), ref_d)  # distances are bit-exact too


def test_knn_channel_window(cuda):
    """kNN on channels 6:9 of a 9-channel S3DIS cloud (DGCNN_S3DIS.py:32) without a copy."""
    from weaksuppointcloudseg_b200 import ops

    rng = np.random.default_rng(7)
    x = rng.uniform(0, 1, (2, 700, 9)).astype(np.float32)
    ref = ok.knn(x, 20, 0, coff=6, D=3)
    got = ops.knn_fused(torch.from_numpy(x).to(cuda), 20, 0, coff=6, D=3)
    assert np.array_equal(got.cpu().numpy(), ref)


def test_self_is_rank0_and_zero(cuda):
    """SURVEY §4 invariant 2: without duplicates each point is its own nearest neighbour at d == 0."""
    from weaksuppointcloudseg_b200 import ops

    rng = np.random.default_rng(3)
    x = torch.from_numpy(rng.uniform(-1, 1, (2, 1000, 64)).astype(np.float32)).to(cuda)
    idx, dist = ops.knn_fused(x, 20, return_dist=True)
    assert torch.equal(idx[:, :, 0].cpu(), torch.arange(1000, dtype=torch.int32).expand(2, -1))
    assert float(dist[:, :, 0].abs().max()) == 0.0


@pytest.mark.parametrize("flavour", [0, 1])
def test_unfused_pair_matches_oracle(cuda, flavour):
    from weaksuppointcloudseg_b200 import ops

    rng = np.random.default_rng(11)
    x = _cloud(rng, 2, 300, 6)
    adj_ref = ok.pairwise_distance(x, flavour)
    adj = ops.pairwise_distance(torch.from_numpy(x).to(cuda), flavour)
    assert np.array_equal(adj.cpu().numpy(), adj_ref)
    idx, vals = ops.topk_rows(adj, 20, return_vals=True)
    ridx, rvals = ok.topk_rows(adj_ref, 20, return_vals=True)
    assert np.array_equal(idx.cpu().numpy(), ridx)
    assert np.array_equal(vals.cpu().numpy(), rvals)
    # fused == unfused
    assert torch.equal(ops.knn_fused(torch.from_numpy(x).to(cuda), 20, flavour), idx)


def test_full_size_properties(cuda):
    """cfg-3-sized batch slice: sortedness + self-inclusion properties (size-independent checks)."""
    from weaksuppointcloudseg_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.rand((8, 4096, 64), device=cuda, generator=g)
    idx, dist = ops.knn_fused(x, 20, return_dist=True)
    assert bool((dist[..., 1:] >= dist[..., :-1]).all())
    assert bool((idx >= 0).all()) and bool((idx < 4096).all())
    assert bool((idx[..., 0] == torch.arange(4096, device=cuda, dtype=torch.int32)).all())
    # no duplicate neighbour within a row
    s = torch.sort(idx, dim=-1).values
    assert bool((s[..., 1:] != s[..., :-1]).all())


def test_errors_are_loud(cuda):
    from weaksuppointcloudseg_b200 import ops
    from weaksuppointcloudseg_b200._lib import WspcError

    x = torch.zeros((1, 16, 3), device=cuda)
    with pytest.raises(WspcError):
        ops.knn_fused(x, 20)  # k > N
    with pytest.raises(WspcError):
        ops.knn_fused(x.cpu(), 4)  # CPU tensor: no fallback


@pytest.mark.parametrize("path", [0, 1], ids=["tcgen05+refine", "cuda-core"])
@pytest.mark.parametrize("B,N,D,kind", [(2, 1024, 64, "relu"), (1, 777, 64, "relu"), (2, 512, 32, "uniform"),
                                        (2, 640, 64, "grid"), (1, 300, 16, "relu"), (1, 2048, 64, "clustered"),
                                        (2, 1500, 3, "uniform"), (2, 900, 6, "uniform"), (1, 2048, 3, "clustered"),
                                        (1, 1111, 9, "relu"), (3, 256, 3, "grid"), (2, 4096, 3, "uniform")])
def test_wide_feature_knn_both_device_paths(cuda, path, B, N, D, kind):
    """D <= 64: tensor-core distances + exact re-scoring must reproduce the oracle bit for bit, including
    exact ties (grid), duplicated points and tight clusters far from the origin (stress for the error margin)."""
    from weaksuppointcloudseg_b200 import _lib as L, ops

    rng = np.random.default_rng(99 + N + D)
    if kind == "clustered":
        centres = rng.normal(0, 1, (8, D)) * 5 + 20.0
        x = (centres[rng.integers(0, 8, (B, N))] + rng.normal(0, 0.01, (B, N, D))).astype(np.float32)
    else:
        x = _cloud(rng, B, N, D, kind=kind)
    ref_idx, ref_d = ok.knn(x, 20, 0, return_dist=True)
    old = L.lib().wspc_set_knn_path(path)
    try:
        idx, dist = ops.knn_fused(torch.from_numpy(x).to(cuda), 20, 0, return_dist=True)
        torch.cuda.synchronize()
    finally:
        L.lib().wspc_set_knn_path(old)
    assert np.array_equal(idx.cpu().numpy(), ref_idx)
    assert np.array_equal(dist.cpu().numpy(), ref_d)


@pytest.mark.parametrize("seed", range(6))
def test_error_margin_stress_bounded_fallback(cuda, seed):
    """tf_util.py:638-671 on hostile data for the tensor-core distance pass: feature-like clouds (D = 64) far from the origin
    (norms up to ~1e4, so |x|^2 is ~1e5 times the neighbour distances) whose points sit in clusters spaced by 2^-15 .. 2^-9
    relative to the coordinates.  The bf16 x 3 distances then cannot separate neighbours; the kernel must notice (proven error
    margin), re-score exactly, and send only a BOUNDED share of rows to the exact per-row fallback.  Results stay bit-exact."""
    import ctypes
    from weaksuppointcloudseg_b200 import _lib as L
    rng = np.random.default_rng(900 + seed)
    B, N, D, k = 2, 1024, 64, 20
    offset = rng.uniform(5.0, 100.0)
    spacing = 2.0 ** rng.uniform(-15.0, -9.0)
    centres = rng.uniform(-1, 1, (B, 64, D)).astype(np.float32) + np.float32(offset)
    member = rng.integers(0, 64, (B, N))
    jitter = rng.integers(-8, 9, (B, N, D)).astype(np.float32) * np.float32(spacing * offset)
    x = (np.take_along_axis(centres, member[..., None].repeat(D, -1), 1) + jitter).astype(np.float32)
    ref_idx, ref_d = ok.knn(x, k, 0, return_dist=True)
    xd = torch.from_numpy(x).to(cuda)
    idx = torch.empty((B, N, k), dtype=torch.int32, device=cuda)
    dist = torch.empty((B, N, k), dtype=torch.float32, device=cuda)
    ws = torch.empty(L.lib().wspc_knn_workspace_bytes(B, N, D), dtype=torch.uint8, device=cuda)
    L.check(L.lib().wspc_knn_fused(L.ptr(xd), B, N, D, 0, D, k, 0, L.ptr(idx), L.ptr(dist), L.ptr(ws), ws.numel(), L.stream()))
    rows = ctypes.c_int(-1)
    L.check(L.lib().wspc_knn_fallback_rows(L.ptr(ws), B, N, D, ctypes.byref(rows)))
    assert np.array_equal(idx.cpu().numpy(), ref_idx)
    assert np.array_equal(dist.cpu().numpy(), ref_d)
    print(f"offset {offset:.1f} spacing 2^{np.log2(spacing):.1f}: {rows.value} of {B * N} rows took the exact fallback")
    assert 0 <= rows.value <= B * N
