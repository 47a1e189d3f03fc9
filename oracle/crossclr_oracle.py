"""CPU oracle for the CrossCLR (intra-modality) criterion -- TEST INFRASTRUCTURE ONLY.

This file is a numpy/float64 restatement of the reference algorithm
(`/root/reference/trainer/loss.py:68-114`, `CrossCLR_onlyIntraModality.forward`, plus the
backward that PyTorch autograd derives from it).  It is the *checker*: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` leg may import it.
The product path (`crossmodal_contrastive_learning_b200`) never calls into this directory.

Parity status: PINNED.  The reference ships no tests, so the pin is the reference itself, run in
the build container: `oracle/make_goldens.py` imports the unmodified `/root/reference/trainer/loss.py`
(only shim: `Tensor.cuda` -> identity, because `.cuda()` is hard-coded at loss.py:66,103,104 and the
container has no GPU) and writes `tests/golden/*.npz`; `tests/test_oracle.py` checks this file
against those vectors and against the closed-form known-answer cases.

The restatement is row-blocked so memory is O(row_block * B) instead of the reference's ~216*B^2
bytes; that is what lets it check B = 16384 ... 131072 on sampled rows.

Formulas (loss.py line numbers on the right):

    v^_i = v_i / max(||v_i||, 1e-12), t^ likewise                       :79-80  (F.normalize eps)
    a    = V^ T^t / tau          (logits_per_text is a^t)               :83-84,:90-91
    cv   = V^ V^t / tau, ct = T^ T^t / tau                               :87-88,:92-93
    intra-modal logits: w * c * (1 - eye)  -> the diagonal stays as logit 0 (e^0 = 1)   :95-100
    row i of the video softmax: [a_i1..a_iB, w cv_i1..w cv_iB], positive = column i     :99,:102-111
    row i of the text  softmax: [a_1i..a_Bi, w ct_i1..w ct_iB], positive = column i     :100,:112
    L = (mean_i(-log p_v,ii) + mean_i(-log p_t,ii)) / 2                  :60,:114
"""
from __future__ import annotations

import numpy as np

EPS = 1e-12  # torch.nn.functional.normalize default eps (loss.py:79-80)


def normalize_rows(x: np.ndarray):
    """loss.py:79-80 -- returns (x_hat, clamped_norm) in float64."""
    x = np.asarray(x, dtype=np.float64)
    n = np.sqrt((x * x).sum(axis=1))
    nc = np.maximum(n, EPS)
    return x / nc[:, None], nc


def _row_stats(fh_rows, row_mod, row_idx, vh, th, tau, w):
    """For a block of rows of the stacked matrix [V^;T^] return (logZ, pos_logit) per row.

    row_mod: 0 for video rows, 1 for text rows; row_idx: sample index of each row.
    Uses a max-shifted log-sum-exp in float64 (the reference uses softmax->log, loss.py:60, which
    is the same number until its float64 softmax underflows at tau ~ 1e-3).
    """
    own = vh if row_mod == 0 else th
    other = th if row_mod == 0 else vh
    inter = fh_rows @ other.T / tau                     # a (video rows) or a^t (text rows)
    intra = w * (fh_rows @ own.T) / tau                 # w * cv or w * ct
    r = np.arange(fh_rows.shape[0])
    pos = inter[r, row_idx].copy()
    intra[r, row_idx] = 0.0                             # (1 - eye) mask: logit 0, still in the softmax
    m = np.maximum(inter.max(axis=1), intra.max(axis=1))
    z = np.exp(inter - m[:, None]).sum(axis=1) + np.exp(intra - m[:, None]).sum(axis=1)
    return m + np.log(z), pos, inter, intra


def loss_only(v, t, temperature=0.03, negative_weight=0.8, row_block=1024):
    """Loss value (python float), loss.py:68-114."""
    vh, _ = normalize_rows(v)
    th, _ = normalize_rows(t)
    B = vh.shape[0]
    acc = 0.0
    for mod, fh in ((0, vh), (1, th)):
        for r0 in range(0, B, row_block):
            idx = np.arange(r0, min(B, r0 + row_block))
            logz, pos, _, _ = _row_stats(fh[idx], mod, idx, vh, th, temperature, negative_weight)
            acc += float((logz - pos).sum())
    return acc / (2.0 * B)


def row_logz(v, t, temperature=0.03, negative_weight=0.8, row_block=1024):
    """Per-row log-normalisers (logZv[B], logZt[B]) and positive logits a_ii[B] in float64."""
    vh, _ = normalize_rows(v)
    th, _ = normalize_rows(t)
    B = vh.shape[0]
    out = np.zeros((2, B))
    pos_all = np.zeros(B)
    for mod, fh in ((0, vh), (1, th)):
        for r0 in range(0, B, row_block):
            idx = np.arange(r0, min(B, r0 + row_block))
            logz, pos, _, _ = _row_stats(fh[idx], mod, idx, vh, th, temperature, negative_weight)
            out[mod, idx] = logz
            if mod == 0:
                pos_all[idx] = pos
    return out[0], out[1], pos_all


def loss_and_grads(v, t, temperature=0.03, negative_weight=0.8, row_block=1024, rows=None,
                   grad_scale=1.0):
    """Loss and dL/dv, dL/dt (float64) -- what autograd produces for loss.py:79-114.

    rows: optional 1-D index array; if given, gradients are produced only for those sample rows
    (returned arrays have len(rows) rows) -- used to spot-check very large batches.

    Derivation: with c = 1/(2B),
        dL/da_ij  = c (e^{a_ij}/Zv_i + e^{a_ij}/Zt_j - 2 delta_ij)
        dL/dcv_ij = c w e^{w cv_ij} (1/Zv_i + 1/Zv_j)   (i != j, 0 on the diagonal), ct likewise
        dV^_i = (1/tau) [ sum_j dL/da_ij t^_j + sum_j dL/dcv_ij v^_j ]
        dT^_i = (1/tau) [ sum_j dL/da_ji v^_j + sum_j dL/dct_ij t^_j ]
        dv_i  = (dV^_i - (dV^_i . v^_i) v^_i) / max(||v_i||, eps)     (no projection if ||v_i|| < eps)
    """
    tau, w = float(temperature), float(negative_weight)
    vh, nv = normalize_rows(v)
    th, nt = normalize_rows(t)
    v64 = np.asarray(v, dtype=np.float64)
    t64 = np.asarray(t, dtype=np.float64)
    B = vh.shape[0]
    logzv, logzt, pos = row_logz(v, t, tau, w, row_block)
    loss = float(((logzv - pos).sum() + (logzt - pos).sum()) / (2.0 * B))
    c = grad_scale / (2.0 * B)
    sel = np.arange(B) if rows is None else np.asarray(rows)
    dv = np.zeros((len(sel), vh.shape[1]))
    dt = np.zeros((len(sel), vh.shape[1]))
    for b0 in range(0, len(sel), row_block):
        idx = sel[b0:b0 + row_block]
        r = np.arange(len(idx))
        # ---- video rows: a_ij (i in idx, all j) and cv_ij
        a = vh[idx] @ th.T / tau
        ga = np.exp(a - logzv[idx, None]) + np.exp(a - logzt[None, :])
        ga[r, idx] -= 2.0
        cv = w * (vh[idx] @ vh.T) / tau
        gv = w * (np.exp(cv - logzv[idx, None]) + np.exp(cv - logzv[None, :]))
        gv[r, idx] = 0.0
        dvh = (c / tau) * (ga @ th + gv @ vh)
        # ---- text rows: a_ji (i in idx as the text index, all j as the video index) and ct_ij
        at = th[idx] @ vh.T / tau                      # at[i, j] = a_ji
        gat = np.exp(at - logzt[idx, None]) + np.exp(at - logzv[None, :])
        gat[r, idx] -= 2.0
        ct = w * (th[idx] @ th.T) / tau
        gt = w * (np.exp(ct - logzt[idx, None]) + np.exp(ct - logzt[None, :]))
        gt[r, idx] = 0.0
        dth = (c / tau) * (gat @ vh + gt @ th)
        # ---- F.normalize backward (x / clamp_min(||x||, eps))
        dv[b0:b0 + len(idx)] = _normalize_backward(dvh, v64[idx], vh[idx], nv[idx])
        dt[b0:b0 + len(idx)] = _normalize_backward(dth, t64[idx], th[idx], nt[idx])
    return loss, dv, dt


def _normalize_backward(dxh, x, xh, nc):
    raw = np.sqrt((x * x).sum(axis=1))
    proj = (dxh * xh).sum(axis=1)
    proj = np.where(raw >= EPS, proj, 0.0)            # clamp active -> no norm gradient
    return (dxh - proj[:, None] * xh) / nc[:, None]


def sharded_loss_and_grads(v, t, world_size, temperature=0.03, negative_weight=0.8):
    """CPU emulation of the row-sharded multi-GPU scheme.

    Rank r owns rows [r*B/P, (r+1)*B/P); it sees all features (all-gather), computes the
    per-row stats of its own rows, all ranks exchange stats (second all-gather), and each rank then
    produces the gradient of the GLOBAL loss w.r.t. its own rows.  Returns (loss, [dv_r], [dt_r]).
    """
    B = np.asarray(v).shape[0]
    assert B % world_size == 0
    bl = B // world_size
    outs_v, outs_t = [], []
    loss = None
    for r in range(world_size):
        rows = np.arange(r * bl, (r + 1) * bl)
        l, dv, dt = loss_and_grads(v, t, temperature, negative_weight, rows=rows)
        loss = l if loss is None else loss
        assert abs(l - loss) < 1e-12
        outs_v.append(dv)
        outs_t.append(dt)
    return loss, outs_v, outs_t


# --------------------------------------------------------------------------------------------
# Closed-form known-answer tests.  Each returns (v, t, tau, w, expected_loss).
def kat_identity(n, tau, w):
    """v = t = I_n  ->  L = log1p((2n-1) e^{-1/tau})."""
    v = np.eye(n)
    return v, v.copy(), tau, w, float(np.log1p((2 * n - 1) * np.exp(-1.0 / tau)))


def kat_collinear(n, tau, w):
    """all rows (3,4), t = 2v -> L = log(n e^{1/tau} + (n-1) e^{w/tau} + 1) - 1/tau."""
    v = np.tile(np.array([[3.0, 4.0]]), (n, 1))
    return v, 2 * v, tau, w, float(np.log(n * np.exp(1 / tau) + (n - 1) * np.exp(w / tau) + 1) - 1 / tau)


def kat_antipodal(n, tau, w):
    """v = I_n, t = -v -> L = log(e^{-1/tau} + 2n - 1) + 1/tau."""
    v = np.eye(n)
    return v, -v, tau, w, float(np.log(np.exp(-1 / tau) + 2 * n - 1) + 1 / tau)
