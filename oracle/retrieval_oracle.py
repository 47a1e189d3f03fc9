"""CPU restatement of the retrieval ranks over `cosine_sim(im, s)` (trainer/loss.py:7-15: a plain mm).  TEST INFRASTRUCTURE
ONLY (tests/ may import it; the product never does).  The reference has no retrieval code of its own (only
figures/qual_retriv.png); the definition follows the hinge indicators of its MaxMargin_coot.forward at margin 0
(trainer/loss.py:34-35): a candidate counts against the partner only when it scores strictly higher.
"""
import numpy as np


def retrieval_ranks(im, s):
    """(rank_im2s, rank_s2im, gap): 0-based ranks in float64 arithmetic, and per direction the smallest |score - partner
    score| over the candidates of each query (how far each rank is from flipping under rounding)."""
    im = np.asarray(im, dtype=np.float64)
    s = np.asarray(s, dtype=np.float64)
    B = im.shape[0]
    scores = im @ s.T
    d = np.diag(scores)
    off = ~np.eye(B, dtype=bool)
    r_a = ((scores > d[:, None]) & off).sum(1)
    r_b = ((scores > d[None, :]) & off).sum(0)
    big = np.where(off, 0.0, np.inf)
    gap_a = (np.abs(scores - d[:, None]) + big).min(1)
    gap_b = (np.abs(scores - d[None, :]) + big).min(0)
    return r_a.astype(np.int64), r_b.astype(np.int64), (gap_a, gap_b)


def recall_at_k(ranks, ks=(1, 5, 10)):
    ranks = np.asarray(ranks)
    return {int(k): float((ranks < k).mean()) for k in ks}
