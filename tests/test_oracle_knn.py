"""CPU-only: the C oracle against an independent numpy restatement + published tf.nn.top_k semantics."""
import numpy as np
import pytest

from oracle import knn as ok


@pytest.mark.parametrize("D,flavour", [(3, 0), (6, 1), (64, 0), (17, 1)])
def test_c_oracle_matches_numpy(D, flavour):
    rng = np.random.default_rng(D)
    x = rng.standard_normal((2, 160, D)).astype(np.float32)
    x[:, 50:60] = x[:, 10:20]  # exact duplicates -> ties
    a = ok.pairwise_distance(x, flavour)
    assert np.array_equal(a, ok.pairwise_distance_numpy(x, flavour))
    assert np.all(np.diagonal(a, axis1=1, axis2=2) == 0)  # canonical chain => exact zero diagonal
    i_fused = ok.knn(x, 20, flavour)
    assert np.array_equal(i_fused, ok.knn_numpy(a, 20))
    assert np.array_equal(i_fused, ok.topk_rows(a, 20))


def test_tie_rule_lower_index_first():
    adj = np.array([[3, 1, 1, 0, 1, 3, 0, 2]], np.float32)
    assert ok.topk_rows(adj, 5).tolist() == [[3, 6, 1, 2, 4]]
    # -0.0 == +0.0 compare as floats
    adj = np.array([[0.0, -0.0, 1.0, -0.0]], np.float32)
    assert ok.topk_rows(adj, 3).tolist() == [[0, 1, 3]]


def test_smooth_flavour_is_clamped():
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((1, 300, 6)) * 100).astype(np.float32)
    assert ok.pairwise_distance(x, ok.SMOOTH).min() >= 0
