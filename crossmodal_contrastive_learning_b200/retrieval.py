"""Retrieval-side consumers of the cross-modal score tiles.

The reference only pictures retrieval (`figures/qual_retriv.png`) and scores two embedding blocks with the plain `mm` of
`cosine_sim` (`trainer/loss.py:7-15`).  What every paper table built on it reports -- R@1 / R@5 / R@10, median and mean
rank, in both directions -- needs one number per query: how many candidates score above the query's own partner.  That is
the indicator count the `MaxMargin_coot` forward (`trainer/loss.py:34-35`) already forms at margin 0, so the ranks come
out of the same tcgen05 score tiles (`crossclr_retrieval_ranks`, csrc/maxmargin_tc.cu) without a B x B matrix, a sort or
a top-k.  Pass L2-normalised rows for cosine ranking; there is no CPU path.
"""
from __future__ import annotations

import torch

from . import _native as N
from .loss import _DTYPE_CODE, _ptr, _rowmajor, _stream


def retrieval_ranks(im: torch.Tensor, s: torch.Tensor):
    """(rank_im2s, rank_s2im): int32 `[B]` tensors, 0-based rank of each query's own partner among all candidates
    (`rank_im2s[i] = #{j != i : im_i . s_j > im_i . s_i}`; ties do not count against the partner)."""
    if im.dim() != 2 or s.dim() != 2 or im.shape != s.shape:
        raise RuntimeError(f"retrieval_ranks expects two [B, D] blocks of one shape, got {tuple(im.shape)} and {tuple(s.shape)}")
    if not (im.is_cuda and s.is_cuda) or im.device != s.device:
        raise RuntimeError("retrieval_ranks (B200-native) needs CUDA tensors on one device: there is no CPU path")
    if im.dtype != s.dtype or im.dtype not in _DTYPE_CODE:
        raise RuntimeError(f"unsupported / mismatched dtypes {im.dtype} and {s.dtype}")
    lib = N.load()
    a, b = _rowmajor(im.detach()), _rowmajor(s.detach())
    B, D = a.shape
    with torch.cuda.device(a.device):
        ws_bytes = int(lib.crossclr_maxmargin_workspace_bytes(B, D, _DTYPE_CODE[a.dtype]))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=a.device)
        r_a = torch.empty(B, dtype=torch.int32, device=a.device)
        r_b = torch.empty(B, dtype=torch.int32, device=a.device)
        N.check(lib.crossclr_retrieval_ranks(_ptr(a), _ptr(b), _DTYPE_CODE[a.dtype], a.stride(0), b.stride(0), B, D,
                                             _ptr(ws), ws_bytes, _ptr(r_a), _ptr(r_b), _stream()),
                "crossclr_retrieval_ranks")
    return r_a, r_b


def recall_at_k(ranks: torch.Tensor, ks=(1, 5, 10)):
    """{k: fraction of queries whose partner is among the top k} from 0-based ranks."""
    return {int(k): float((ranks < k).double().mean()) for k in ks}


def retrieval_metrics(im: torch.Tensor, s: torch.Tensor, ks=(1, 5, 10)):
    """The usual video-text retrieval table for one batch of paired embeddings: R@k, median rank (1-based) and mean rank
    (1-based) for `im -> s` and `s -> im`."""
    out = {}
    for name, r in zip(("im2s", "s2im"), retrieval_ranks(im, s)):
        rec = recall_at_k(r, ks)
        for k in ks:
            out[f"{name}_R@{int(k)}"] = rec[int(k)]
        out[f"{name}_MedR"] = float(r.double().median()) + 1.0
        out[f"{name}_MeanR"] = float(r.double().mean()) + 1.0
    return out
