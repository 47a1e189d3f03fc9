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


class HostFedCrossCLR:
    """A whole criterion step -- host-to-device copy of the inputs, forward, backward, device-to-host copy of the loss --
    for training loops whose embeddings arrive in pinned host memory: ONE CUDA graph launch (pack + forward + backward)
    plus copy-engine transfers on a side stream per step.

    Two buffer sets alternate.  `step()` number k
      * enqueues, on a copy stream, the upload of the pinned `host_video[(k + 1) & 1]` / `host_text[(k + 1) & 1]` into the
        other set's device inputs (it starts once step k - 1, the last reader of that set, has finished), and
      * replays the graph of set k & 1: forward + backward on the inputs uploaded during step k - 1; its loss is copied into
        the pinned `loss_host[k & 1]` at the head of step k + 1's transfers (or by `drain()` after the last step),
    so every step pays its own H2D copy and its own D2H read, but the copy of step k + 1 hides under the kernels of step k
    and the host issues one graph launch per step (the eager module, or even two graph replays threaded through autograd,
    cost more host time per step than the kernels take at B = 4096).  The transfers are deliberately NOT graph nodes: memcpy
    nodes captured with these kernels -- persistent, one CTA per SM, all of shared memory -- ran 3-4x slower than the same
    copies issued on a stream (measured on B200: 668 / 479 vs 171 us per step).

        pipe = HostFedCrossCLR(crit, batch=4096, dim=512)
        pipe.host_video[0].copy_(v0); pipe.host_text[0].copy_(t0); pipe.prime()      # upload of step 0
        for k in range(steps):
            fill pipe.host_video[(k + 1) & 1], pipe.host_text[(k + 1) & 1]            # next step's embeddings
            s = pipe.step()
            ...  pipe.grad_video[s], pipe.grad_text[s] (device; consume them on the current stream before the next step)
        pipe.drain()                                                                   # last loss; loss_host[s] is valid
                                                                                       # after pipe.done[s].synchronize()

    `feed="device"` captures forward + backward only (inputs already resident in `video[s]` / `text[s]`, loss left in
    `loss[s]`): the device-resident variant of the same single-launch step.

    Same kernels and numbers as the eager module; `temperature` / `negative_w` are frozen at capture time.  With a process
    group the NCCL all-gathers are captured as well (every rank must build and step its pipeline in lock-step).
    """

    def __init__(self, criterion: CrossCLR_onlyIntraModality, batch: int, dim: int, dtype=torch.bfloat16, device=None,
                 warmup: int = 3, feed: str = "host"):
        if not torch.cuda.is_available():
            raise RuntimeError("HostFedCrossCLR needs a CUDA device (the criterion has no CPU path)")
        if feed not in ("host", "device"):
            raise ValueError("feed must be 'host' (inputs uploaded from pinned memory every step) or 'device' (already resident)")
        self.feed = feed
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.criterion = criterion
        # Both blocks of a set share ONE pinned buffer and ONE device buffer, so a step's inputs cross PCIe as a single copy
        # (a copy's fixed cost is far from negligible on some hosts: two 4 MiB copies were measured at 12 GB/s where one
        # large one reaches 45 GB/s); `host_video` / `host_text` / `video` / `text` are views of the halves.
        half = (batch * dim + 127) // 128 * 128          # elements per half: the text half stays 256-byte aligned
        self._host_in = [torch.zeros(2 * half, dtype=dtype).pin_memory() for _ in range(2)]
        self.host_video = [h[:batch * dim].view(batch, dim) for h in self._host_in]
        self.host_text = [h[half:half + batch * dim].view(batch, dim) for h in self._host_in]
        self.loss_host = [torch.zeros((), dtype=torch.float64).pin_memory() for _ in range(2)]
        self._dev_in = [torch.zeros(2 * half, dtype=dtype, device=dev).normal_() for _ in range(2)]
        self.video = [d[:batch * dim].view(batch, dim).detach().requires_grad_() for d in self._dev_in]
        self.text = [d[half:half + batch * dim].view(batch, dim).detach().requires_grad_() for d in self._dev_in]
        self.grad_video, self.grad_text, self.loss = [None, None], [None, None], [None, None]
        self.done = [torch.cuda.Event(), torch.cuda.Event()]
        self._grad_out = torch.ones((), dtype=torch.float64, device=dev)
        self._k = 0
        self.h2d_bytes_per_step = 2 * batch * dim * self.host_video[0].element_size() if feed == "host" else 0
        self.d2h_bytes_per_step = 8 if feed == "host" else 0
        with torch.cuda.device(dev):
            side = torch.cuda.Stream()
            self._copy_stream = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):
                    for s in range(2):
                        loss = criterion(self.video[s], self.text[s])
                        torch.autograd.grad(loss, (self.video[s], self.text[s]), grad_outputs=self._grad_out)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(dev)
            self._graphs = []
            pool = None
            for s in range(2):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool):
                    loss = criterion(self.video[s], self.text[s])
                    self.grad_video[s], self.grad_text[s] = torch.autograd.grad(loss, (self.video[s], self.text[s]),
                                                                                grad_outputs=self._grad_out)
                    self.loss[s] = loss.detach()
                pool = g.pool()
                self._graphs.append(g)

    def prime(self):
        """Upload the inputs of the very next step from `host_video[k & 1]` / `host_text[k & 1]` (pipeline prologue)."""
        s = self._k & 1
        self._dev_in[s].copy_(self._host_in[s], non_blocking=True)

    def step(self) -> int:
        """The current set's forward + backward + loss read-back (one graph launch) beside the upload of the next set.
        Returns the set index."""
        s = self._k & 1
        cur = torch.cuda.current_stream()
        if self.feed == "host":
            self._copy_stream.wait_stream(cur)              # the previous step (last reader of the other set) is done
            with torch.cuda.stream(self._copy_stream):
                if self._k > 0:                             # its loss goes home first: 8 bytes
                    self.loss_host[s ^ 1].copy_(self.loss[s ^ 1], non_blocking=True)
                    self.done[s ^ 1].record()
                self._dev_in[s ^ 1].copy_(self._host_in[s ^ 1], non_blocking=True)
        self._graphs[s].replay()
        if self.feed == "host":
            cur.wait_stream(self._copy_stream)              # the next step computes on what was just uploaded
        else:
            self.done[s].record()
        self._k += 1
        return s

    def drain(self):
        """Read back the loss of the most recent step (inside the loop the read-back of step k rides on step k + 1)."""
        if self.feed != "host" or self._k == 0:
            return
        s = (self._k - 1) & 1
        cur = torch.cuda.current_stream()
        self._copy_stream.wait_stream(cur)
        with torch.cuda.stream(self._copy_stream):
            self.loss_host[s].copy_(self.loss[s], non_blocking=True)
            self.done[s].record()
        cur.wait_stream(self._copy_stream)
