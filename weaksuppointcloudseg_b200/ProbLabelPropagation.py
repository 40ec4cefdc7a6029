"""Drop-in for Util/ProbLabelPropagation.py (reference :3-62): closed-form label propagation.

Same class / method names; `sess` arguments are accepted and ignored.  The solve runs on the GPU through
wspc_lp_blocks (SPD system solved with preconditioned CG instead of the reference's dense inverse)."""
from __future__ import annotations

import torch

from . import ops


class LabelPropagation_TF():
    '''Propagates per-point class probabilities over a graph Laplacian by solving the regularised linear system
    (alpha L + beta diag(w) + 1e-5 I) Y = beta diag(w) G once; w is the confidence 1 - normalised entropy of G.
    Mirrors the reference class of the same name (constructor and method signatures).'''

    def __init__(self, alpha, beta, K):
        self.alpha = alpha
        self.beta = beta
        # self.K = K   (the reference ignores K as well, ProbLabelPropagation.py:11)

    def set_alpha(self, alpha):
        self.alpha = alpha

    def set_beta(self, beta):
        self.beta = beta

    def SolveLabelProp(self, sess, L, G):
        '''One propagation.  L: (N,N) symmetric normalised Laplacian, G: (N,K) network probabilities (numpy or tensor).
        Returns numpy Y (N,K), Y_prob = Y / row sums, w (N,) -- the fetch list of the reference's sess.run (:44-57).'''
        Y, Yp, w = ops.lp_solve(L, G, float(self.alpha), float(self.beta))
        self.Y_val, self.Y_prob_val, self.w_val = Y.cpu().numpy(), Yp.cpu().numpy(), w.cpu().numpy()
        return self.Y_val, self.Y_prob_val, self.w_val

    def EvalWeight4EachPoint(self, sess, G):
        """w = 1 - H_2(G)/log_2 K for every point (reference :59-62)"""
        G = torch.as_tensor(G, dtype=torch.float32)
        N, K = G.shape
        _, _, w = ops.lp_solve(torch.zeros((N, N)), G, 0.0, 1.0, max_iter=1)
        self.w_val = [w.cpu().numpy()]
        return self.w_val
