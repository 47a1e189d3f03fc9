"""World-size-2 gloo test (CPU) of the criterion's multi-rank orchestration: row sharding, stacked-row layout,
the two all-gathers, the global loss and the per-rank gradients of the GLOBAL loss.

The product only ever runs `_NativeOps` (CUDA).  Here the private `_forward_impl` / `_backward_impl` are driven with
a checker-backed stand-in that implements the C-ABI contract of include/crossclr_b200.h in numpy, so the host
logic (rank -> row_begin, segment order, in-place all-gathers, what is saved for backward) is covered without a GPU.
The result is compared with the oracle on the concatenated batch.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

LOG2E = 1.4426950408889634


class NumpyOps:
    """The contract of include/crossclr_b200.h, restated in numpy float64 (test stand-in for _NativeOps).  `pad` > 1 mimics the
    tensor-core layout: every segment of the stacked matrix / stats / coef is padded with zero rows to a multiple of `pad`."""

    def __init__(self, pad=1):
        self.pad = pad

    def plan(self, prob, in_dtype, exact, force=None):
        from crossmodal_contrastive_learning_b200 import _native as N
        return N.PATH_SIMT, torch.float32, prob.dim

    def seg_rows(self, code, bseg):
        return (bseg + self.pad - 1) // self.pad * self.pad

    def _real(self, prob):
        """padded stacked indices of the real rows, in stacked order"""
        S = self.seg_rows(None, prob.bseg)
        return (np.arange(prob.nseg)[:, None] * S + np.arange(prob.bseg)[None, :]).reshape(-1)

    @staticmethod
    def _shift(prob):
        return max(0.0, LOG2E * max(1.0, abs(prob.negative_weight)) / prob.temperature - 96.0)

    def pack2(self, v, t, feat_out, rnorm_out, code=None):
        B = v.shape[0]
        feat_out.zero_()
        for k, x in enumerate((v, t)):
            x64 = x.double().numpy()
            n = np.maximum(np.sqrt((x64 * x64).sum(1)), 1e-12)
            feat_out[k, :B].copy_(torch.from_numpy(x64 / n[:, None]).to(feat_out.dtype))
            rnorm_out[k * B:(k + 1) * B].copy_(torch.from_numpy(1.0 / n).float())

    @staticmethod
    def _logits(prob, F, g_rows):
        """log2-domain shifted logits of the given (unpadded) stacked rows against all rows, plus masks."""
        bseg = prob.bseg
        R = F.shape[0]
        g_all = np.arange(R)
        mod = (g_all // bseg) & 1
        samp = (g_all // bseg >> 1) * bseg + g_all % bseg
        G = F[g_rows] @ F.T
        same_mod = mod[g_rows][:, None] == mod[None, :]
        same_samp = samp[g_rows][:, None] == samp[None, :]
        k = np.where(same_mod, prob.negative_weight, 1.0) * LOG2E / prob.temperature
        return G * k - NumpyOps._shift(prob), same_mod, same_samp

    def fwd(self, prob, code, feat_all, stats):
        real = self._real(prob)
        F = feat_all.reshape(-1, prob.dim).double().numpy()[real]
        rows = np.arange(prob.row_begin, prob.row_begin + prob.row_count)
        x, same_mod, same_samp = self._logits(prob, F, rows)
        e = np.exp2(x)
        X = np.where(same_samp, 0.0, e).sum(1) + 2.0 ** (-self._shift(prob))     # + the masked intra diagonal (logit 0)
        xpos = x[same_samp & ~same_mod]
        stats[real[rows]] = torch.from_numpy(np.stack([X, xpos], 1)).float()

    def finalize(self, prob, code, stats, coef, loss, scal):
        real = self._real(prob)
        s = stats.double().numpy()[real]
        X, xp = s[:, 0], s[:, 1]
        Z = X + np.exp2(xp)
        coef.zero_()
        coef[real] = torch.from_numpy(np.stack([1.0 / Z, X / Z], 1)).float()
        loss.copy_(torch.tensor(np.log1p(X * np.exp2(-xp)).sum() / len(X), dtype=torch.float64))
        scal.copy_(torch.tensor([1.0, 1.0, float((X / Z).max()), 0.0]))

    def forward_single(self, prob, code, v, t, feat_all, rnorm, stats, coef, scal, loss):
        self.pack2(v, t, feat_all, rnorm)
        self.fwd(prob, code, feat_all, stats)
        self.finalize(prob, code, stats, coef, loss, scal)

    def bwd(self, prob, code, feat_all, rnorm, coef, scal, grad_out, grad_scale, dv, dt, dscale=None, logit_scale=1.0):
        real = self._real(prob)
        F = feat_all.reshape(-1, prob.dim).double().numpy()[real]
        R, bseg = F.shape[0], prob.bseg
        rows = np.arange(prob.row_begin, prob.row_begin + prob.row_count)
        x, same_mod, same_samp = self._logits(prob, F, rows)
        c = coef.double().numpy()[real]
        iz, rho = c[:, 0], c[:, 1]
        P = np.exp2(x) * (iz[rows][:, None] + iz[None, :]) * np.where(same_mod, prob.negative_weight, 1.0)
        P[same_samp] = 0.0
        partner = np.where((rows // bseg) & 1, rows - bseg, rows + bseg)
        h = P @ F - (rho[rows] + rho[partner])[:, None] * F[partner]
        rn = rnorm.double().numpy()
        dot = (h * F[rows]).sum(1)
        dot = np.where(rn >= 1e12, 0.0, dot)
        m = float(grad_out) * grad_scale / prob.temperature / R
        out = m * rn[:, None] * (h - dot[:, None] * F[rows])
        dv.copy_(torch.from_numpy(out[:bseg]).to(dv.dtype))
        dt.copy_(torch.from_numpy(out[bseg:]).to(dt.dtype))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, D, tau, w, grad_scale, out, pad=1):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from crossmodal_contrastive_learning_b200 import loss as L
        g = torch.Generator().manual_seed(123)
        v = torch.randn(B, D, generator=g)
        t = v + 2.0 * torch.randn(B, D, generator=g)
        bl = B // world
        vl, tl = v[rank * bl:(rank + 1) * bl].contiguous(), t[rank * bl:(rank + 1) * bl].contiguous()
        ops = NumpyOps(pad)
        loss, prob, code, saved = L._forward_impl(ops, vl, tl, tau, w, "auto", dist.group.WORLD)
        assert (prob.nseg, prob.bseg, prob.row_begin, prob.row_count) == (2 * world, bl, 2 * rank * bl, 2 * bl)
        dv, dt = L._backward_impl(ops, prob, code, saved, torch.tensor(1.0, dtype=torch.float64), grad_scale,
                                  torch.float32)
        out[rank] = (float(loss), dv.numpy().copy(), dt.numpy().copy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,B,D,grad_scale,pad", [(2, 64, 32, 1.0, 1), (2, 96, 24, 2.0, 1), (2, 72, 24, 1.0, 16)],
                         ids=["b64", "b96", "b72_padded_segments"])
def test_sharded_matches_global_oracle(world, B, D, grad_scale, pad):
    from oracle import crossclr_oracle as O
    tau, w = 0.05, 0.8
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, B, D, tau, w, grad_scale, out, pad), nprocs=world, join=True)
    g = torch.Generator().manual_seed(123)
    v = torch.randn(B, D, generator=g)
    t = v + 2.0 * torch.randn(B, D, generator=g)
    rloss, rdv, rdt = O.loss_and_grads(v.numpy(), t.numpy(), tau, w, grad_scale=grad_scale)
    sl, sdv, sdt = O.sharded_loss_and_grads(v.numpy(), t.numpy(), world, tau, w)
    assert abs(sl - rloss) < 1e-12
    bl = B // world
    for r in range(world):
        loss, dv, dt = out[r]
        assert abs(loss - rloss) <= 1e-6 * abs(rloss), (r, loss, rloss)          # identical global loss on every rank
        sl_ = slice(r * bl, (r + 1) * bl)
        for a, b in ((dv, rdv[sl_]), (dt, rdt[sl_])):
            assert np.linalg.norm(a - b) <= 2e-5 * np.linalg.norm(b), (r, np.linalg.norm(a - b) / np.linalg.norm(b))
        assert np.linalg.norm(dv - grad_scale * sdv[r]) <= 2e-5 * np.linalg.norm(sdv[r]) * grad_scale


def test_single_rank_orchestration_matches_oracle():
    """world = 1 goes through the fused forward entry point; same contract."""
    from oracle import crossclr_oracle as O
    from crossmodal_contrastive_learning_b200 import loss as L
    g = torch.Generator().manual_seed(5)
    v = torch.randn(48, 20, generator=g)
    t = torch.randn(48, 20, generator=g)
    ops = NumpyOps()
    loss, prob, code, saved = L._forward_impl(ops, v, t, 0.03, 0.8, "auto", None)
    dv, dt = L._backward_impl(ops, prob, code, saved, torch.tensor(0.5, dtype=torch.float64), 1.0, torch.float32)
    rloss, rdv, rdt = O.loss_and_grads(v.numpy(), t.numpy(), 0.03, 0.8, grad_scale=0.5)
    assert abs(float(loss) - rloss) <= 1e-6 * abs(rloss)
    assert np.linalg.norm(dv.numpy() - rdv) <= 2e-5 * np.linalg.norm(rdv)
    assert np.linalg.norm(dt.numpy() - rdt) <= 2e-5 * np.linalg.norm(rdt)
