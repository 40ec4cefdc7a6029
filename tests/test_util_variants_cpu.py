"""Util/Loss.py drop-in (row a19) against what the REFERENCE'S OWN Util/Loss.py returns on the tf1_shim
(tests/golden/make_util_golden.py -> ref_util_variants.npz).  The functions are plain tensor expressions, so they run on CPU."""
import os

import numpy as np
import pytest
import torch

from weaksuppointcloudseg_b200 import Loss

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_util_variants.npz"))
t = lambda n: torch.from_numpy(G[n])   # noqa: E731


def close(a, name, tol=2e-6):
    ref = G[name]
    a = a.detach().numpy()
    assert a.shape == ref.shape, (name, a.shape, ref.shape)
    assert np.abs(a - ref).max() <= tol * max(1.0, np.abs(ref).max()), (name, np.abs(a - ref).max())


def test_fixture_is_not_vacuous():
    for n in ("OverwhelmLoss_v1", "OverwhelmLoss_v2", "OverwhelmLoss"):
        assert float(G[n]) > 1e-2, n
    assert G["sm_mask_fraction"][0] > 0.05


def test_focal_losses():
    close(Loss.focal_loss(t("ls_L"), t("ls_Ypt")), "focal_loss")
    close(Loss.focal_loss(t("ls_L"), t("ls_Ypt"), alpha=0.4, gamma=3), "focal_loss_a4_g3")
    close(Loss.focal_loss_v1(t("ls_L"), t("ls_Ypt")), "focal_loss_v1")
    close(Loss.focal_loss_v1(t("ls_L"), t("ls_Ypt"), alpha=t("ls_alpha")), "focal_loss_v1_alpha")


def test_class_weighted_ce_and_self_entropy():
    close(Loss.class_weighted_CE_loss(t("ls_L")[:, :1], t("ls_Ypt")[:, :1], t("ls_pw"), t("ls_nw")), "class_weighted_CE_loss")
    close(Loss.SelfEntropy(t("ls_L")), "SelfEntropy")


def test_overwhelm_losses():
    close(Loss.OverwhelmLoss_v1(t("ls_L"), t("ls_Ycl")), "OverwhelmLoss_v1")
    l2, p2, n2 = Loss.OverwhelmLoss_v2(t("ls_L"), t("ls_Ycl"))
    close(l2, "OverwhelmLoss_v2"); close(p2, "OverwhelmLoss_v2_pos"); close(n2, "OverwhelmLoss_v2_neg")
    l3, f3 = Loss.OverwhelmLoss(t("ls_L"), t("ls_Ycl"))
    close(l3, "OverwhelmLoss"); close(f3, "OverwhelmLoss_full")
