"""CUDA-graph capture of the criterion's forward + backward ("caller integration").

At the headline shape (B=4096, D=512) one fwd+bwd is ~0.2 ms of kernels issued through ~10 launches and half a
dozen allocations; the Python / ctypes / autograd launch path costs about as much again, so an eager training loop
is launch-bound.  `GraphedCrossCLR` captures the whole step once -- every kernel of libcrossclr_b200.so, the
memsets, the workspace -- into two CUDA graphs (forward, backward; the layout `torch.cuda.make_graphed_callables`
uses) and replays them:

    crit = CrossCLR_onlyIntraModality(0.03, 0.8).cuda()
    step = GraphedCrossCLR(crit, batch=4096, dim=512, dtype=torch.bfloat16)
    loss = step(video, text)          # copies into the static inputs (skipped if they ARE the static inputs), replays
    loss.backward()                   # replays the backward graph; video.grad / text.grad as usual

Semantics are those of the eager module (same kernels, same numbers); what changes is ownership: the returned loss
and the gradients are views of static buffers that the next call overwrites, and `temperature` / `negative_w` are
frozen at capture time (re-capture after changing them).  With a process group the two NCCL all-gathers are captured
with the kernels (every rank must construct and replay its `GraphedCrossCLR` in the same order, as with any
collective).
"""
from __future__ import annotations

import torch

from .loss import CrossCLR_onlyIntraModality


class _Replay(torch.autograd.Function):
    @staticmethod
    def forward(ctx, runner, video, text):
        if video.data_ptr() != runner.video.data_ptr():
            runner.video.detach().copy_(video, non_blocking=True)
        if text.data_ptr() != runner.text.data_ptr():
            runner.text.detach().copy_(text, non_blocking=True)
        runner._fwd.replay()
        ctx.runner = runner
        return runner._loss.detach()

    @staticmethod
    def backward(ctx, grad_out):
        r = ctx.runner
        r._grad_out.copy_(grad_out, non_blocking=True)
        r._bwd.replay()
        return None, r._gv.detach(), r._gt.detach()


class GraphedCrossCLR:
    """fwd+bwd of a `CrossCLR_onlyIntraModality` captured in CUDA graphs for one (batch, dim, dtype).

    Attributes `video`, `text` are the static input buffers ([batch, dim], requires_grad): fill them in place (e.g. as
    the target of the H2D copy) and call `step(step.video, step.text)` to skip the device-to-device copy."""

    def __init__(self, criterion: CrossCLR_onlyIntraModality, batch: int, dim: int, dtype=torch.bfloat16,
                 device=None, warmup: int = 3):
        if not torch.cuda.is_available():
            raise RuntimeError("GraphedCrossCLR needs a CUDA device (the criterion has no CPU path)")
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.criterion = criterion
        self.video = torch.zeros(batch, dim, dtype=dtype, device=dev).normal_().requires_grad_()
        self.text = torch.zeros(batch, dim, dtype=dtype, device=dev).normal_().requires_grad_()
        self._grad_out = torch.ones((), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            # Warm up on a side stream (the usual capture recipe): first-call work that must not happen under capture
            # -- loading the library, binding the primary context on the autograd thread, occupancy queries, function
            # attributes -- happens here.
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):
                    loss = criterion(self.video, self.text)
                    torch.autograd.grad(loss, (self.video, self.text), grad_outputs=self._grad_out)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(dev)
            self._fwd = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._fwd):
                self._loss = criterion(self.video, self.text)
            self._bwd = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._bwd, pool=self._fwd.pool()):
                self._gv, self._gt = torch.autograd.grad(self._loss, (self.video, self.text),
                                                         grad_outputs=self._grad_out)

    def __call__(self, video_features, text_features):
        """Same contract as `criterion(video_features, text_features)`; the result is a static buffer (see module doc)."""
        if video_features.shape != self.video.shape or text_features.shape != self.text.shape:
            raise RuntimeError(f"captured for {tuple(self.video.shape)} features, got {tuple(video_features.shape)} and "
                               f"{tuple(text_features.shape)}")
        if video_features.dtype != self.video.dtype or text_features.dtype != self.text.dtype:
            raise RuntimeError(f"captured for {self.video.dtype} features, got {video_features.dtype} and {text_features.dtype}")
        return _Replay.apply(self, video_features, text_features)
