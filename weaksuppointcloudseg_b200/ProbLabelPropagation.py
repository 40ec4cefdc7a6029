"""Drop-in for Util/ProbLabelPropagation.py (reference :3-62): closed-form label propagation.

Same class / method names; `sess` arguments are accepted and ignored.  The solve runs on the GPU through
wspc_lp_solve (SPD system solved with preconditioned CG instead of the reference's dense inverse)."""
from __future__ import annotations

import torch

from . import ops


class LabelPropagation_TF():
    '''
    The baseline method for label propagation. The closed-form solution is adopted for label propagation.
    '''

    def __init__(self, alpha, beta, K):
        self.alpha = alpha
        self.beta = beta
        # self.K = K   (the reference ignores K as well, ProbLabelPropagation.py:11)

    def set_alpha(self, alpha):
        self.alpha = alpha

    def set_beta(self, beta):
        self.beta = beta

    def SolveLabelProp(self, sess, L, G):
        '''
        Solve label propagation with closed-form solution
        :param L:   Laplacian matrix N*N
        :param G:   network prediction N*K (dense)
        :return: Y, Y_prob, w   (numpy, like sess.run)
        '''
        Y, Yp, w = ops.lp_solve(L, G, float(self.alpha), float(self.beta))
        self.Y_val, self.Y_prob_val, self.w_val = Y.cpu().numpy(), Yp.cpu().numpy(), w.cpu().numpy()
        return self.Y_val, self.Y_prob_val, self.w_val

    def EvalWeight4EachPoint(self, sess, G):
        G = torch.as_tensor(G, dtype=torch.float32)
        N, K = G.shape
        Lz = torch.zeros((N + (-N) % 8, N + (-N) % 8))
        Gp = torch.full((Lz.shape[0], K), 1.0 / K)
        Gp[:N] = G
        _, _, w = ops.lp_solve(Lz, Gp, 0.0, 1.0, max_iter=1)
        self.w_val = [w[:N].cpu().numpy()]
        return self.w_val
