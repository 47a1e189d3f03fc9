"""CPU restatement of the reference's `MaxMargin_coot.forward` (trainer/loss.py:29-41) and the gradients autograd
derives from it.  TEST INFRASTRUCTURE ONLY (tests/ and __graft_entry__.smoke may import it; the product never does).

    scores = im @ s^T                                   (:30, `cosine_sim` :7-15 is a plain mm -- no normalisation)
    d_i    = scores_ii                                  (:31-33)
    cost_s[i][j]  = max(0, margin + scores_ij - d_i)    (:34)   row-wise hinge
    cost_im[i][j] = max(0, margin + scores_ij - d_j)    (:35)   column-wise hinge
    both with the diagonal zeroed (:36-40);  loss = (sum cost_s + sum cost_im) / (B * B)     (:41)

Pinned by tests/golden/maxmargin/*.npz, generated from the unmodified reference by oracle/make_goldens_maxmargin.py
(the reference class cannot be constructed -- its ctor names a class that does not exist, trainer/loss.py:24 -- so the
generator calls the unbound `forward` with a stand-in `self`; the arithmetic is the reference's own).
"""
import numpy as np


def maxmargin_loss_and_grads(im, s, margin=0.1, grad_out=1.0):
    """loss, dL/dim, dL/ds in float64."""
    im = np.asarray(im, dtype=np.float64)
    s = np.asarray(s, dtype=np.float64)
    B = im.shape[0]
    scores = im @ s.T
    d = np.diag(scores)
    cs = np.maximum(0.0, margin + scores - d[:, None])
    ci = np.maximum(0.0, margin + scores - d[None, :])
    off = ~np.eye(B, dtype=bool)
    loss = float((cs[off].sum() + ci[off].sum()) / (B * B))
    # d loss / d scores: indicators off the diagonal; the diagonal collects -(row count of cs) - (column count of ci)
    c = grad_out / (B * B)
    G = ((cs > 0) & off).astype(np.float64) + ((ci > 0) & off).astype(np.float64)
    rowcnt = ((cs > 0) & off).sum(1)
    colcnt = ((ci > 0) & off).sum(0)
    G[np.arange(B), np.arange(B)] = -(rowcnt + colcnt)
    G *= c
    return loss, G @ s, G.T @ im
