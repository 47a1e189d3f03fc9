"""Host-side mirror of the reference's `MaxMargin_coot` (`trainer/loss.py:17-41`), SURVEY.md section 8 row f1.

Same constructor signature `(use_cuda: bool, margin: float = 0.1)`, attributes (`margin`, `sim`, `use_cuda`) and
`forward(im, s)` -> 0-dim loss in the input dtype, gradients in the input dtype.  One documented deviation: the reference
class cannot be constructed at all (`trainer/loss.py:24` calls `super(ContrastiveLoss_coot, self)`, a name that does not
exist, so `MaxMargin_coot(...)` raises NameError); this one constructs.  The arithmetic is `trainer/loss.py:29-41`
(`cosine_sim`, :7-15, is a plain `mm`: no normalisation) behind the C ABI (`crossclr_maxmargin_fwd/bwd`): fp16 / bf16
inputs run on the tensor cores (tcgen05 score tiles with a hinge epilogue, csrc/maxmargin_tc.cu), fp32 inputs and small
problems on exact fp32 CUDA-core kernels; there is no CPU path.
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn

from . import _native as N
from .loss import _DTYPE_CODE, _ptr, _rowmajor, _stream


def cosine_sim(emb1, emb2):
    """trainer/loss.py:7-15 (kept for attribute parity; MaxMargin_coot.forward does not call it)."""
    return emb1.mm(emb2.t())


class _MaxMarginFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, im, s, margin):
        if im.dim() != 2 or s.dim() != 2:
            raise RuntimeError(f"MaxMargin_coot expects 2-D [B, D] embeddings, got {tuple(im.shape)} and {tuple(s.shape)}")
        if im.shape[1] != s.shape[1]:
            raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied ({tuple(im.shape)} and {tuple(s.t().shape)})")
        if im.shape[0] != s.shape[0]:
            raise RuntimeError(f"batch sizes differ: {im.shape[0]} vs {s.shape[0]} (the score matrix must be square)")
        if not (im.is_cuda and s.is_cuda) or im.device != s.device:
            raise RuntimeError("MaxMargin_coot (B200-native) needs CUDA tensors on one device: there is no CPU path")
        if im.dtype != s.dtype or im.dtype not in _DTYPE_CODE:
            raise RuntimeError(f"unsupported / mismatched dtypes {im.dtype} and {s.dtype}")
        lib = N.load()
        a, b = _rowmajor(im.detach()), _rowmajor(s.detach())
        B, D = a.shape
        with torch.cuda.device(a.device):
            ws_bytes = int(lib.crossclr_maxmargin_workspace_bytes(B, D, _DTYPE_CODE[a.dtype]))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=a.device)
            loss = torch.empty((), dtype=torch.float64, device=a.device)
            N.check(lib.crossclr_maxmargin_fwd(_ptr(a), _ptr(b), _DTYPE_CODE[a.dtype], a.stride(0), b.stride(0), B, D,
                                               float(margin), _ptr(ws), ws_bytes, _ptr(loss), _stream()),
                    "crossclr_maxmargin_fwd")
        ctx.save_for_backward(a, b, ws)
        ctx.margin = float(margin)
        return loss.to(im.dtype)                  # the reference returns the input dtype (trainer/loss.py:41)

    @staticmethod
    def backward(ctx, grad_out):
        a, b, ws = ctx.saved_tensors
        B, D = a.shape
        lib = N.load()
        with torch.cuda.device(a.device):
            go = grad_out.detach().to(device=a.device, dtype=torch.float64).contiguous()
            da, db = torch.empty_like(a), torch.empty_like(b)
            N.check(lib.crossclr_maxmargin_bwd(_ptr(a), _ptr(b), _DTYPE_CODE[a.dtype], a.stride(0), b.stride(0), B, D,
                                               ctx.margin, _ptr(ws), ws.numel(), _ptr(go), _ptr(da), da.stride(0), _ptr(db),
                                               db.stride(0), _DTYPE_CODE[a.dtype], _stream()), "crossclr_maxmargin_bwd")
        return da, db, None


class MaxMargin_coot(nn.Module):
    """Bidirectional max-margin (hinge) ranking loss over the B x B score matrix of two [B, D] embedding blocks (the COOT
    loss); drop-in for the reference class of the same name (`trainer/loss.py:17-41`)."""

    def __init__(self, use_cuda: bool, margin: float = 0.1):
        super().__init__()                        # (the reference's super(ContrastiveLoss_coot, ...) is a NameError)
        self.margin = margin                      # trainer/loss.py:25
        self.sim = cosine_sim                     # :26
        self.use_cuda = use_cuda                  # :27

    def forward(self, im, s):
        return _MaxMarginFunction.apply(im, s, self.margin)
