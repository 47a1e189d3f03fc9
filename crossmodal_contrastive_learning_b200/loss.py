"""Host-side mirror of the reference criterion module for the CrossCLR hot path.

`CrossCLR_onlyIntraModality` keeps the surface of the reference `nn.Module`
(`trainer/loss.py:44-114`): same constructor `(temperature=0.03, negative_weight=0.8, logger=None)`
(`:50`), same attributes (`temperature`, `negative_w`, `logger`, the dormant `logit_scale` Parameter and the
unused `criterion` child, `:52-56`), same `forward(video_features, text_features)` (`:68`) returning a 0-dim
float64 tensor on the inputs' device with gradients in the input dtype, and plain `RuntimeError` on shape /
device mismatches.  Everything between the inputs and the loss runs in libcrossclr_b200.so (hand-written
sm_100a kernels behind the C ABI of include/crossclr_b200.h); torch only owns device memory, the stream and,
for more than one rank, the two NCCL all-gathers.

There is no CPU / eager fallback: without the native library, or with CPU tensors, `forward` raises.

Multi-GPU (keyword-only extension; the reference is single-device): rank r holds rows
[r*B, (r+1)*B) of the global batch.  Normalised rows are all-gathered, every rank reduces the statistics of
its own rows, the per-row statistics are all-gathered, and every rank returns the identical GLOBAL loss and
the gradient of that global loss w.r.t. its own rows (times `grad_scale`) -- no gradient collective.
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn

from . import _native as N

_DTYPE_CODE = {torch.float32: N.F32, torch.float16: N.F16, torch.bfloat16: N.BF16}
_FEAT_TORCH = {N.F32: torch.float32, N.F16: torch.float16, N.F16X2: torch.float16}
_PATH_CODE = {"auto": N.PATH_AUTO, "simt": N.PATH_SIMT, "tc": N.PATH_TC, "split": N.PATH_TC_SPLIT}
_FORCED = {"tc": N.PATH_TC, "split": N.PATH_TC_SPLIT}       # "simt" goes through crossclr_choose_path(exact=1)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check_inputs(v, t):
    # The reference validates nothing itself; torch raises RuntimeError from inside its ops
    # (trainer/loss.py:83 for ndim / D mismatch, :97 for B mismatch, :66 for CPU tensors).  Same class here.
    if v.dim() != 2 or t.dim() != 2:
        raise RuntimeError(f"CrossCLR expects 2-D [B, D] features, got {tuple(v.shape)} and {tuple(t.shape)}")
    if v.shape[1] != t.shape[1]:
        raise RuntimeError(f"feature dims differ: {v.shape[1]} vs {t.shape[1]} (mat1 and mat2 shapes cannot be multiplied)")
    if v.shape[0] != t.shape[0]:
        raise RuntimeError(f"batch sizes differ: {v.shape[0]} vs {t.shape[0]}")
    if v.shape[0] < 1 or v.shape[1] < 1:
        raise RuntimeError("CrossCLR needs at least one row and one feature")
    if not (v.is_cuda and t.is_cuda):
        raise RuntimeError("CrossCLR_onlyIntraModality (B200-native) needs CUDA tensors: the criterion has no CPU path")
    if v.device != t.device:
        raise RuntimeError(f"Expected all tensors to be on the same device, got {v.device} and {t.device}")
    if v.dtype != t.dtype:
        raise RuntimeError(f"expected video and text features of the same dtype, got {v.dtype} and {t.dtype}")
    if v.dtype not in _DTYPE_CODE and v.dtype != torch.float64:
        raise RuntimeError(f"unsupported feature dtype {v.dtype}")


def _rowmajor(x):
    return x if (x.stride(1) == 1 and x.stride(0) >= x.shape[1]) else x.contiguous()


class _NativeOps:
    """The C ABI of include/crossclr_b200.h on torch tensors (device pointers + the current stream).

    This is the only implementation the criterion ever uses.  `_forward_impl` / `_backward_impl` take the ops object
    as an argument so the rank / shard / all-gather orchestration can be exercised on CPU (gloo) by tests/ with a
    checker-backed stand-in; nothing in the package selects anything but `_NativeOps`.
    """

    def __init__(self):
        self.lib = N.load()

    def plan(self, prob, in_dtype, exact, force=None):
        """(path code, stacked-row dtype, stacked-row pitch in elements) the library picks for this problem; `force`: a path
        code to take as is (a shape or temperature outside its range is refused at launch)."""
        if force:
            code = force
        else:
            code = self.lib.crossclr_choose_path(ctypes.byref(prob), _DTYPE_CODE[in_dtype], 1 if exact else 0)
            if code < 0:
                N.check(code, "crossclr_choose_path")
        return code, _FEAT_TORCH[self.lib.crossclr_feature_dtype(code)], int(self.lib.crossclr_feature_pitch(code, prob.dim))

    def seg_rows(self, code, bseg):
        """Rows per segment of the stacked matrix / stats / coef of a path (the tensor-core layout pads segments to 128 rows)."""
        return int(self.lib.crossclr_segment_rows(code, bseg))

    def _feat_code(self, feat_out, code):
        return _DTYPE_CODE[feat_out.dtype] if code is None else self.lib.crossclr_feature_dtype(code)

    def pack(self, x, feat_out, rnorm_out, code=None):
        B, D = x.shape
        N.check(self.lib.crossclr_pack(_ptr(x), _DTYPE_CODE[x.dtype], x.stride(0), B, D, _ptr(feat_out),
                                       self._feat_code(feat_out, code), _ptr(rnorm_out), _stream()), "crossclr_pack")

    def pack2(self, v, t, feat_out, rnorm_out, code=None):
        """`code`: the path the rows are for (fixes the row layout); None: plain rows of feat_out's dtype."""
        B, D = v.shape
        N.check(self.lib.crossclr_pack2(_ptr(v), _ptr(t), _DTYPE_CODE[v.dtype], v.stride(0), t.stride(0), B, D,
                                        _ptr(feat_out), self._feat_code(feat_out, code), _ptr(rnorm_out), _stream()),
                "crossclr_pack2")

    def forward_single(self, prob, code, v, t, feat_all, rnorm, stats, coef, scal, loss):
        """pack2 + fwd + finalize in one C call (single rank)."""
        N.check(self.lib.crossclr_forward(ctypes.byref(prob), code, _ptr(v), _ptr(t), _DTYPE_CODE[v.dtype], v.stride(0),
                                          t.stride(0), _ptr(feat_all), _ptr(rnorm), _ptr(stats), _ptr(coef), _ptr(scal),
                                          _ptr(loss), _stream()), "crossclr_forward")

    def fwd(self, prob, code, feat_all, stats):
        N.check(self.lib.crossclr_fwd(ctypes.byref(prob), code, _ptr(feat_all), _ptr(stats), None, 0, _stream()),
                "crossclr_fwd")

    def finalize(self, prob, code, stats, coef, loss, scal):
        N.check(self.lib.crossclr_finalize(ctypes.byref(prob), code, _ptr(stats), _ptr(coef), _ptr(loss), _ptr(scal),
                                           _stream()), "crossclr_finalize")

    def bwd(self, prob, code, feat_all, rnorm, coef, scal, grad_out, grad_scale, dv, dt, dscale=None, logit_scale=1.0):
        """dscale: optional 0-dim float64 tensor receiving d loss / d logit_scale (the problem then carries the effective
        temperature tau / logit_scale); the gradient kernel, the scale gradient and the finish stage run as separate calls."""
        ws_bytes = int(self.lib.crossclr_workspace_bytes(ctypes.byref(prob), code))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=feat_all.device)
        if dscale is None:
            N.check(self.lib.crossclr_bwd(ctypes.byref(prob), code, _ptr(feat_all), _ptr(rnorm), _ptr(coef), _ptr(scal),
                                          _ptr(grad_out), grad_scale, _ptr(dv), dv.stride(0), _ptr(dt), dt.stride(0),
                                          _DTYPE_CODE[dv.dtype], _ptr(ws), ws_bytes, _stream()), "crossclr_bwd")
            return
        N.check(self.lib.crossclr_bwd_accumulate(ctypes.byref(prob), code, _ptr(feat_all), _ptr(coef), _ptr(scal), _ptr(ws),
                                                 ws_bytes, _stream()), "crossclr_bwd_accumulate")
        N.check(self.lib.crossclr_bwd_scale_grad(ctypes.byref(prob), code, _ptr(feat_all), _ptr(coef), _ptr(scal), _ptr(grad_out),
                                                 grad_scale, logit_scale, _ptr(ws), _ptr(dscale), _stream()),
                "crossclr_bwd_scale_grad")
        N.check(self.lib.crossclr_bwd_finish(ctypes.byref(prob), code, _ptr(feat_all), _ptr(rnorm), _ptr(coef), _ptr(scal),
                                             _ptr(grad_out), grad_scale, _ptr(dv), dv.stride(0), _ptr(dt), dt.stride(0),
                                             _DTYPE_CODE[dv.dtype], _ptr(ws), _stream()), "crossclr_bwd_finish")


def _group_info(group):
    if group is None:
        return 1, 0
    import torch.distributed as dist
    return dist.get_world_size(group), dist.get_rank(group)


def _forward_impl(ops, v, t, temperature, negative_weight, path, group, exchange="nccl"):
    """pack -> [all-gather features] -> row statistics of the owned rows -> [all-gather stats] -> finalize.

    Rank r owns stacked rows [2 r B, 2 (r+1) B): its video rows then its text rows (include/crossclr_b200.h).
    exchange: "nccl" (two `all_gather_into_tensor`) | "peer" (one node: `crossclr_peer_exchange`, peer stores over NVLink +
    a flag barrier in one kernel per exchange; the stacked matrix and the statistics live in an IPC-mapped buffer, peer.py).
    Returns (loss, prob, code, saved tensors)."""
    B, D = v.shape
    dev = v.device
    world, rank = _group_info(group)
    prob = N.Problem(2 * world, B, D, 2 * rank * B, 2 * B, float(temperature), float(negative_weight))
    code, feat_dtype, pitch = ops.plan(prob, v.dtype, path == "simt", _FORCED.get(path))
    S = ops.seg_rows(code, B)                                   # rows per segment in the path's layout (>= B: zero padding)
    rows = 2 * world * S
    plan = None
    if world > 1 and exchange == "peer":
        from . import peer
        esz = torch.empty((), dtype=feat_dtype).element_size()
        plan = peer.plan_for(group, dev, 2 * world * S * pitch * esz, rows * 8)
        feat_all, stats = plan.feat((2 * world, S, pitch), feat_dtype), plan.stats(rows)
        plan.generation += 1                 # the saved rows are a view of the plan's buffer: a later forward overwrites them
    else:
        feat_all = torch.empty((2 * world, S, pitch), dtype=feat_dtype, device=dev)
        stats = torch.empty((rows, 2), dtype=torch.float32, device=dev)
    rnorm = torch.empty(2 * B, dtype=torch.float32, device=dev)
    coef = torch.empty((rows, 2), dtype=torch.float32, device=dev)          # separate allocations: the custom op returns
    scal = torch.empty(4, dtype=torch.float32, device=dev)                  # coef and scal, and outputs must not alias
    loss = torch.empty((), dtype=torch.float64, device=dev)
    if world == 1:
        ops.forward_single(prob, code, v, t, feat_all, rnorm, stats, coef, scal, loss)
        return loss, prob, code, (feat_all, rnorm, coef, scal)
    import torch.distributed as dist
    feat_loc = feat_all[2 * rank:2 * rank + 2]
    if plan is not None:
        # the entry barrier of the first exchange: no peer stores into this rank's rows before it is done with the previous
        # step (its backward reads them), and pack below must not run ahead of the peers' previous readers either -- pack
        # writes LOCAL memory only, which no peer reads, so it may go first
        ops.pack2(v, t, feat_loc, rnorm, code)
        fb = feat_loc.numel() * feat_loc.element_size()
        plan.exchange(rank * fb, fb, True, _stream())
        ops.fwd(prob, code, feat_all, stats)
        sb = 2 * S * 8
        plan.exchange(plan.feat_bytes + rank * sb, sb, False, _stream())
        ops.finalize(prob, code, stats, coef, loss, scal)
        return loss, prob, code, (feat_all, rnorm, coef, scal)
    ops.pack2(v, t, feat_loc, rnorm, code)
    # in-place all-gather: rank r's block already sits at its slot of the output
    dist.all_gather_into_tensor(feat_all.view(-1), feat_loc.reshape(-1), group=group)
    ops.fwd(prob, code, feat_all, stats)
    dist.all_gather_into_tensor(stats.view(-1), stats[2 * rank * S:2 * (rank + 1) * S].reshape(-1), group=group)
    ops.finalize(prob, code, stats, coef, loss, scal)
    return loss, prob, code, (feat_all, rnorm, coef, scal)


def _backward_impl(ops, prob, code, saved, grad_out, grad_scale, out_dtype, logit_scale=None):
    """logit_scale: None, or the float value s the forward ran at (prob.temperature = tau / s); then a third result,
    d loss / d s as a 0-dim float64 tensor, is returned."""
    feat_all, rnorm, coef, scal = saved
    dev = feat_all.device
    go = grad_out.detach().to(device=dev, dtype=torch.float64).contiguous()
    dv = torch.empty((prob.bseg, prob.dim), dtype=out_dtype, device=dev)
    dt = torch.empty((prob.bseg, prob.dim), dtype=out_dtype, device=dev)
    if logit_scale is None:
        ops.bwd(prob, code, feat_all, rnorm, coef, scal, go, float(grad_scale), dv, dt)
        return dv, dt
    ds = torch.empty((), dtype=torch.float64, device=dev)
    ops.bwd(prob, code, feat_all, rnorm, coef, scal, go, float(grad_scale), dv, dt, ds, float(logit_scale))
    return dv, dt, ds


_native_ops = None


def _ops():
    global _native_ops
    if _native_ops is None:
        _native_ops = _NativeOps()
    return _native_ops


class _CrossCLRFunction(torch.autograd.Function):
    """logit_scale: None, or the module's `logit_scale` Parameter (opt-in learnable temperature): the kernels then run at the
    effective temperature tau / logit_scale and the Parameter receives d loss / d logit_scale."""

    @staticmethod
    def forward(ctx, video, text, temperature, negative_weight, path, group, grad_scale, logit_scale=None, exchange="nccl"):
        ops = _ops()
        _check_inputs(video, text)
        in_dtype = video.dtype
        if in_dtype == torch.float64:          # kernels compute in fp32; the reference's f64 inputs are down-cast
            video, text = video.float(), text.float()
        v, t = _rowmajor(video.detach()), _rowmajor(text.detach())
        ctx.scale = None
        if logit_scale is not None:
            ctx.scale = float(logit_scale.detach())      # one host read per step: the kernels take the temperature by value
            if not (ctx.scale > 0.0):
                raise RuntimeError(f"logit_scale must stay positive (got {ctx.scale})")
            ctx.scale_dtype = logit_scale.dtype
            temperature = float(temperature) / ctx.scale
        with torch.cuda.device(v.device):
            loss, prob, code, saved = _forward_impl(ops, v, t, temperature, negative_weight, path, group, exchange)
        ctx.peer_plan = ctx.peer_generation = None
        if group is not None and exchange == "peer" and _group_info(group)[0] > 1:
            from . import peer
            esz = saved[0].element_size()
            ctx.peer_plan = peer.plan_for(group, v.device, saved[0].numel() * esz, saved[2].shape[0] * 8)
            ctx.peer_generation = ctx.peer_plan.generation
        ctx.save_for_backward(*saved)
        ctx.prob, ctx.code, ctx.in_dtype, ctx.grad_scale = prob, code, in_dtype, float(grad_scale)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        saved = ctx.saved_tensors
        if ctx.peer_plan is not None and ctx.peer_plan.generation != ctx.peer_generation:
            raise RuntimeError("exchange='peer': this forward's stacked rows were overwritten by a later forward on the same "
                               "process group (one outstanding forward per group; use exchange='nccl' otherwise)")
        out_dtype = torch.float32 if ctx.in_dtype == torch.float64 else ctx.in_dtype
        with torch.cuda.device(saved[0].device):
            res = _backward_impl(_ops(), ctx.prob, ctx.code, saved, grad_out, ctx.grad_scale, out_dtype, ctx.scale)
        dv, dt = res[0], res[1]
        if ctx.in_dtype == torch.float64:
            dv, dt = dv.double(), dt.double()
        ds = res[2].to(ctx.scale_dtype) if ctx.scale is not None else None
        return dv, dt, None, None, None, None, None, ds, None


# ---------------------------------------------------------------------------------------------------
# torch.library registration: the single-rank criterion as two opaque custom ops, so that
# torch.compile traces straight through it (no graph break at the ctypes calls) and AMP / functorch see ordinary ops.
#   crossclr_b200::forward(video, text, temperature, negative_weight, path) -> (loss, feat, rnorm, coef, scal)
#   crossclr_b200::backward(feat, rnorm, coef, scal, grad_out, batch, dim, temperature, negative_weight, path,
#                           grad_scale, out_dtype) -> (dvideo, dtext)
# Multi-rank calls (process_group=...) keep using the autograd.Function above: a ProcessGroup is not an op argument.
_OUT_DTYPE = {0: torch.float32, 1: torch.float16, 2: torch.bfloat16}


def _plan_py(B, D, path, in_dtype, temperature, negative_weight):
    """(dtype, pitch, segment rows) of the stacked matrix libcrossclr_b200 uses for a single-rank problem, asked of the library
    itself (pure functions of the problem and dtype: no device needed), so that shape inference can never drift from the
    kernels' rule."""
    prob = N.Problem(2, B, D, 0, 2 * B, float(temperature), float(negative_weight))
    ops = _ops()
    code, fdt, pitch = ops.plan(prob, torch.float32 if in_dtype == torch.float64 else in_dtype, path == "simt", _FORCED.get(path))
    return fdt, pitch, ops.seg_rows(code, B)


@torch.library.custom_op("crossclr_b200::forward", mutates_args=())
def _op_forward(video: torch.Tensor, text: torch.Tensor, temperature: float, negative_weight: float,
                path: str) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    _check_inputs(video, text)
    v, t = _rowmajor(video.detach()), _rowmajor(text.detach())
    with torch.cuda.device(v.device):
        loss, _, _, (feat_all, rnorm, coef, scal) = _forward_impl(_ops(), v, t, temperature, negative_weight, path, None)
    return loss, feat_all, rnorm, coef, scal


@_op_forward.register_fake
def _(video, text, temperature, negative_weight, path):
    B, D = video.shape
    fdt, pitch, S = _plan_py(B, D, path, video.dtype, temperature, negative_weight)
    dev = video.device
    return (torch.empty((), dtype=torch.float64, device=dev), torch.empty((2, S, pitch), dtype=fdt, device=dev),
            torch.empty(2 * B, dtype=torch.float32, device=dev), torch.empty((2 * S, 2), dtype=torch.float32, device=dev),
            torch.empty(4, dtype=torch.float32, device=dev))


@torch.library.custom_op("crossclr_b200::backward", mutates_args=())
def _op_backward(feat: torch.Tensor, rnorm: torch.Tensor, coef: torch.Tensor, scal: torch.Tensor, grad_out: torch.Tensor,
                 batch: int, dim: int, temperature: float, negative_weight: float, path: str, grad_scale: float,
                 out_dtype: int) -> tuple[torch.Tensor, torch.Tensor]:
    prob = N.Problem(2, batch, dim, 0, 2 * batch, float(temperature), float(negative_weight))
    if feat.dtype == torch.float32:
        code = N.PATH_SIMT
    else:                                  # the saved rows tell which tensor-core layout the forward used
        split_pitch = int(_ops().lib.crossclr_feature_pitch(N.PATH_TC_SPLIT, dim))
        code = N.PATH_TC_SPLIT if path == "split" or (path == "auto" and feat.shape[-1] == split_pitch) else N.PATH_TC
    with torch.cuda.device(feat.device):
        return _backward_impl(_ops(), prob, code, (feat, rnorm, coef, scal), grad_out, grad_scale,
                              _OUT_DTYPE[out_dtype])


@_op_backward.register_fake
def _(feat, rnorm, coef, scal, grad_out, batch, dim, temperature, negative_weight, path, grad_scale, out_dtype):
    dt = _OUT_DTYPE[out_dtype]
    return (torch.empty((batch, dim), dtype=dt, device=feat.device), torch.empty((batch, dim), dtype=dt, device=feat.device))


def _op_setup_context(ctx, inputs, output):
    video, _, temperature, negative_weight, path = inputs
    _, feat, rnorm, coef, scal = output
    ctx.save_for_backward(feat, rnorm, coef, scal)
    ctx.set_materialize_grads(False)          # no zero-filled gradients for the four saved-state outputs
    ctx.meta = (int(video.shape[0]), int(video.shape[1]), float(temperature), float(negative_weight), path,
                _DTYPE_CODE[video.dtype])


def _op_backward_formula(ctx, g_loss, g_feat, g_rnorm, g_coef, g_scal):
    feat, rnorm, coef, scal = ctx.saved_tensors
    B, D, tau, w, path, dcode = ctx.meta
    if g_loss is None:                        # the loss itself was not used downstream
        return None, None, None, None, None
    dv, dt = torch.ops.crossclr_b200.backward(feat, rnorm, coef, scal, g_loss, B, D, tau, w, path, 1.0, dcode)
    return dv, dt, None, None, None


_op_forward.register_autograd(_op_backward_formula, setup_context=_op_setup_context)


def _criterion(video, text, temperature, negative_weight, path, group, grad_scale, logit_scale=None, exchange="nccl"):
    """Dispatch: custom ops on a single rank (traceable), the autograd.Function with a process group or a learnable scale."""
    if exchange not in ("nccl", "peer"):
        raise ValueError("exchange must be 'nccl' or 'peer'")
    if group is not None or logit_scale is not None:
        return _CrossCLRFunction.apply(video, text, temperature, negative_weight, path, group, grad_scale, logit_scale,
                                       exchange)
    _check_inputs(video, text)
    if video.dtype == torch.float64:          # kernels compute in fp32; autograd casts the gradients back
        video, text = video.float(), text.float()
    loss = torch.ops.crossclr_b200.forward(video, text, float(temperature), float(negative_weight), path)[0]
    return loss if grad_scale == 1.0 else _ScaleGrad.apply(loss, float(grad_scale))


class _ScaleGrad(torch.autograd.Function):
    """identity forward, gradient times `scale` (the grad_scale knob on a single rank)"""

    @staticmethod
    def forward(ctx, x, scale):
        ctx.scale = scale
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g * ctx.scale, None


class CrossCLR_onlyIntraModality(nn.Module):
    """CrossCLR Loss between 2 groups of embeddings - Only Intra Modality alignment.

    Drop-in for the reference module of the same name (`trainer/loss.py:44-114`).  Keyword-only extensions
    (defaults reproduce the reference's single-device behaviour):
      process_group  torch.distributed group whose ranks each hold a row shard of the global batch
      grad_scale     multiplies the returned gradients (set to world_size under DDP's gradient averaging)
      exchange       "nccl" (default): two NCCL all-gathers per step | "peer": ranks of ONE node store their row shard
                     straight into every rank's stacked matrix over NVLink (CUDA-IPC-mapped buffers, one kernel per exchange
                     with the cross-rank barrier inside, csrc/peer.cu).  The first step maps the buffers (collective, outside
                     graph capture); one forward may be outstanding per group (its saved rows live in the shared buffer)
      path           "auto" | "tc" (tcgen05 kernels, fp16 operands) | "split" (tcgen05 kernels, fp32 inputs as fp16 hi + lo
                     pairs) | "simt" (exact-fp32 CUDA-core kernels).  "auto": 16-bit inputs -> "tc"; fp32 inputs -> "split"
                     where it applies (>= 1024 global samples, D <= 1024), else "simt"; temperatures below ~0.0073 and tiny
                     shapes -> "simt"
      learnable_temperature  False (the reference: `logit_scale` is registered at trainer/loss.py:52 and never used) | True:
                     every logit is multiplied by the `logit_scale` Parameter (effective temperature
                     `temperature / logit_scale`) and the Parameter receives its gradient.  Costs one host read of the
                     scalar per step (the kernels take the temperature by value), so such a step is not graph-capturable;
                     with a process group each rank holds its own rows' share of the gradient, which DDP's all-reduce sums
                     like any other parameter gradient (times `grad_scale`).
    """

    def __init__(self, temperature=0.03, negative_weight=0.8, logger=None, *, process_group=None, grad_scale=1.0,
                 path="auto", learnable_temperature=False, exchange="nccl"):
        super().__init__()
        self.logit_scale = nn.Parameter(torch.ones([]))            # trainer/loss.py:52 (registered, never used)
        self.criterion = torch.nn.CrossEntropyLoss(reduction='none')  # :53 (registered, never used)
        self.temperature = temperature                             # :54
        self.logger = logger                                       # :55
        self.negative_w = negative_weight                          # :56 (attribute name differs from the arg)
        if path not in _PATH_CODE:
            raise ValueError(f"path must be one of {sorted(_PATH_CODE)}")
        self.process_group = process_group
        self.grad_scale = grad_scale
        self.path = path
        self.learnable_temperature = bool(learnable_temperature)
        if exchange not in ("nccl", "peer"):
            raise ValueError("exchange must be 'nccl' or 'peer'")
        self.exchange = exchange

    def forward(self, video_features, text_features):
        """
        Inputs shape (batch, embed_dim)
        Args:
            video_features: Video embeddings (batch, embed_dim)
            text_features: Text embeddings (batch, embed_dim)
        Returns: 0-dim float64 loss (trainer/loss.py:114)
        """
        # temperature / negative_w are read per call (trainer/loss.py:90-93, :99-100)
        return _criterion(video_features, text_features, self.temperature, self.negative_w, self.path,
                          self.process_group, self.grad_scale, self.logit_scale if self.learnable_temperature else None,
                          self.exchange)


def crossclr_loss(video_features, text_features, temperature=0.03, negative_weight=0.8, *, process_group=None,
                  grad_scale=1.0, path="auto", exchange="nccl"):
    """Functional form of `CrossCLR_onlyIntraModality.forward`."""
    return _criterion(video_features, text_features, temperature, negative_weight, path, process_group, grad_scale,
                      exchange=exchange)
