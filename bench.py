#!/usr/bin/env python
"""CrossCLR fwd+bwd throughput benchmark (BASELINE.json metric: pairs/s at [B, D]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c3|c4|c5]

One step = one forward + one backward of the criterion over one synthetic [B, D] video/text batch
(a "pair" = one (video_i, text_i) row pair).  N > 1 is launched by torchrun, one rank
per GPU; the GLOBAL batch is sharded by rows (strong scaling: the global workload is fixed and the same for every N).

Workloads (BASELINE.json `configs`; tau = 0.03, w = 0.8):
    c2  B=4096   D=512   bf16   (configs[1]; the headline `value` / `e2e` at EVERY N, the config the reference arm can run)
    c3  B=16384  D=1024  bf16   (configs[2]; recorded under `workloads` at N = 1)
    c4  B=65536  D=512   bf16   (configs[3]; global batch, recorded under `workloads` at every N)
    c5  B=131072 D=1024  bf16   (configs[4]; global batch, recorded under `workloads` at every N)

The step is ONE CUDA-graph launch (`HostFedCrossCLR`, the package's public capture API; `--no-graph` times the eager
module instead).  `value`: inputs resident in HBM, graph = pack + forward + backward.  `e2e`: the same step fed from pinned
host memory -- the H2D copy of step i+1 is captured on a second stream beside step i's kernels and the loss is copied back
to pinned memory inside the graph, so every step pays its own H2D and D2H inside the timed region.

Printed JSON keys follow the bench contract: value, e2e, roofline (dominant kernel = the backward kernel, named by the
library, per-kernel CUDA events from the library's timing hooks), cpu_baseline (the reference criterion on this box's host
cores), clocks, gpu_launches; plus `parity` (each rank's loss / gradients against the CPU oracle, MAX over ranks) and
`workloads` (the other BASELINE configs, measured in the same process).

`--impl reference` times the reference's own CPU implementation (baseline/_ref/trainer/loss.py, an untracked
verbatim copy, when present; else the numpy oracle port) on the host cores; nothing of the product runs there.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    "c2": dict(B=4096, D=512, desc="BASELINE configs[1]: B=4096 D=512 bf16"),
    "c3": dict(B=16384, D=1024, desc="BASELINE configs[2]: B=16384 D=1024 bf16"),
    "c4": dict(B=65536, D=512, desc="BASELINE configs[3]: global B=65536 D=512 bf16"),
    "c5": dict(B=131072, D=1024, desc="BASELINE configs[4]: global B=131072 D=1024 bf16"),
}
TAU, W = 0.03, 0.8
METRIC = "crossclr_fwd_bwd_pairs_per_sec"
UNIT = "pairs/s"


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops"]), float(p["bf16_tflops_sustained"]), "measured"
    except Exception:
        return 1590.0, 1400.0, "fallback"


def ncu_traffic(workload, kernel):
    """DRAM bytes per launch of the dominant kernel, from the committed `ncu --set full` capture of this workload -- only if
    that capture is of the kernel the library launches now (else None)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            rec = json.load(f)[workload]
        return rec["bytes"] if rec.get("kernel") == kernel else None
    except Exception:
        return None


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference criterion on the host cores
def _reference_runner():
    """Returns (kind, fn(v, t) -> loss) running fwd+bwd on CPU fp32 tensors."""
    import torch
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if os.path.exists(os.path.join(ref_dir, "trainer", "loss.py")):
        import importlib.util
        spec = importlib.util.spec_from_file_location("_crossclr_reference_loss", os.path.join(ref_dir, "trainer", "loss.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        crit = mod.CrossCLR_onlyIntraModality(temperature=TAU, negative_weight=W)

        def step(v, t):
            # placement shim only: the reference hard-codes .cuda() (trainer/loss.py:66,103,104)
            import torch as _t
            orig = _t.Tensor.cuda
            _t.Tensor.cuda = lambda self, *a, **k: self
            try:
                v = v.detach().requires_grad_()
                t = t.detach().requires_grad_()
                loss = crit(v, t)
                loss.backward()
                return float(loss)
            finally:
                _t.Tensor.cuda = orig
        return "reference", step

    from oracle import crossclr_oracle as O

    def step(v, t):
        loss, _, _ = O.loss_and_grads(v.numpy(), t.numpy(), TAU, W)
        return loss
    return "port", step


def time_reference(B, D, steps, warmup, max_seconds=150.0):
    """pairs/s of the reference criterion on the host cores (all threads), fp32, bounded by max_seconds."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kind, step = _reference_runner()
    # bound the sample: the reference materialises ~216*B^2 bytes and takes ~1.8 s at B=4096 on 8 cores
    Bs = min(B, 4096)
    g = torch.Generator().manual_seed(0)
    v = torch.randn(Bs, D, generator=g)
    t = torch.randn(Bs, D, generator=g)
    times = []
    t_begin = time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step(v, t)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if time.perf_counter() - t_begin > max_seconds and len(times) >= 1:
            break
    best = min(times)
    mean = sum(times) / len(times)
    sample = (f"{len(times)} timed fwd+bwd steps of B={Bs} D={D} fp32 on {cores} threads ({cpu_model()}); "
              f"mean {mean * 1e3:.1f} ms, best {best * 1e3:.1f} ms"
              + ("" if Bs == B else f"; batch capped at {Bs} of {B}: the reference needs ~216*B^2 bytes"))
    return dict(value=Bs / mean, unit=UNIT, cores=cores, kind=kind, sample=sample), mean, len(times)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    cb, mean, n = time_reference(wl["B"], wl["D"], args.steps, args.warmup, max_seconds=900.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": n,
        "warmup": args.warmup, "ms_per_step": mean * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {wl['desc']} (reference CPU path, fp32 features)", "B": wl["B"],
                   "D": wl["D"], "temperature": TAU, "negative_weight": W},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / throttle reasons during the timed region, sampled by an `nvidia-smi -lms` child process (the
    recipe of B200_PROFILING.md).  A child process, not a thread: NVML polling from this interpreter holds the GIL
    and slows the Python launch path that is being timed.  Samples are filtered to the [begin(), end()] window."""

    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=5):
        import subprocess
        self.t0 = self.t1 = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-i", str(index), "-lms", str(period_ms)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def summary(self):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.05)
        self.proc.terminate()
        try:
            text, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            return out
        clocks, allclocks, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in text.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) != 7:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                mhz, mx = int(f[1]), int(f[2])
            except ValueError:
                continue
            out["sm_max_mhz"] = mx
            allclocks.append(mhz)
            if self.t0 is not None and self.t0 - 0.005 <= ts <= self.t1 + 0.005:
                clocks.append(mhz)
                for n, v in zip(names, f[3:]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        use = clocks or allclocks
        use.sort()
        out.update(sm_mhz=(use[len(use) // 2] if use else None), reasons=sorted(reasons), samples=len(clocks),
                   window_ms=round((self.t1 - self.t0) * 1e3, 1) if self.t0 else None,
                   window="value + e2e timed regions + 0.4 s of the same step (untimed) so that nvidia-smi yields samples")
        return out


def synthetic_shard(Bg, D, lo, hi, seed=0):
    """Rows [lo, hi) of the global synthetic batch (same seed on every rank), bf16, pinned."""
    import torch
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(2):
        dst = torch.empty(hi - lo, D, dtype=torch.bfloat16).pin_memory()
        chunk = 8192
        for r0 in range(0, Bg, chunk):
            blk = torch.randn(min(chunk, Bg - r0), D, generator=g)
            a, b = max(r0, lo), min(r0 + blk.shape[0], hi)
            if a < b:
                dst[a - lo:b - lo] = blk[a - r0:b - r0].to(torch.bfloat16)
        out.append(dst)
    return out


EXCHANGE = ["nccl"]          # how the ranks of this run exchange row shards (set once by run_b200_arm)


def kernel_times_eager(crit, v_dev, t_dev, steps, flush, NAT, torch):
    """Per-kernel device time of eager steps (the library's own stream-ordered cudaEvents, recorded around each launch
    after all host-side preparation)."""
    v_e, t_e = v_dev.detach().requires_grad_(), t_dev.detach().requires_grad_()
    crit(v_e, t_e).backward()
    torch.cuda.synchronize()
    NAT.timing_read()
    NAT.timing_enable(True)
    for _ in range(steps):
        flush.zero_()
        v_e, t_e = v_dev.detach().requires_grad_(), t_dev.detach().requires_grad_()
        crit(v_e, t_e).backward()
    torch.cuda.synchronize()
    NAT.timing_enable(False)
    return NAT.timing_read()


def isolated_kernel_ms(M, NAT, torch, crit, v_dev, t_dev, group, flush, reps):
    """Median device time of the dominant kernel ALONE: the C ABI's `crossclr_bwd_accumulate` (control-word memset + the
    backward similarity / gradient kernel, nothing else) and `crossclr_fwd` (statistics memset + forward kernel) of this
    rank's real problem, each captured as its own CUDA graph and replayed between two CUDA events on the launching stream, L2
    flushed before every replay.  No host work falls inside the bracket (eager brackets carry ~10 us of launch latency)."""
    import ctypes
    from crossmodal_contrastive_learning_b200 import loss as L
    ops = L._ops()
    lib = M.load_native()
    with torch.no_grad():
        _, prob, code, (feat, rnorm, coef, scal) = L._forward_impl(ops, v_dev, t_dev, TAU, W, "auto", group)
    stats = torch.empty((prob.nseg * prob.bseg, 2), dtype=torch.float32, device=v_dev.device)
    ws_bytes = int(lib.crossclr_workspace_bytes(ctypes.byref(prob), code))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=v_dev.device)

    def bwd():
        NAT.check(lib.crossclr_bwd_accumulate(ctypes.byref(prob), code, L._ptr(feat), L._ptr(coef), L._ptr(scal), L._ptr(ws),
                                              ws_bytes, L._stream()), "crossclr_bwd_accumulate")

    def fwd():
        NAT.check(lib.crossclr_fwd(ctypes.byref(prob), code, L._ptr(feat), L._ptr(stats), None, 0, L._stream()), "crossclr_fwd")

    out = {}
    for name, fn in (("bwd", bwd), ("fwd", fwd)):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        ts = []
        for _ in range(reps):
            flush.zero_()
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record()
            g.replay()
            b_.record()
            torch.cuda.synchronize()
            ts.append(a_.elapsed_time(b_))
        ts.sort()
        out[name] = {"median_ms": ts[len(ts) // 2], "min_ms": ts[0], "mean_ms": sum(ts) / len(ts), "launches": len(ts)}
    return out


def run_extra_workload(name, M, NAT, torch, dist, world, rank, dev, group, flush, steps=3, warmup=2):
    """One of the other BASELINE configs, eager module calls (host overhead is noise at these sizes): ms/step (device,
    max over ranks), pairs/s, per-kernel ms, step fraction of the measured sustained bf16 peak."""
    wl = WORKLOADS[name]
    Bg, D = wl["B"], wl["D"]
    Bl = Bg // world
    need = 26 * Bl * D + 2 * Bg * D * 2 + 40 * Bg + (160 << 20)        # inputs, grads, stacked rows, fp32 accumulator, pool
    free, _ = torch.cuda.mem_get_info()
    if free < need:
        return {"skipped": f"needs ~{need >> 20} MiB of device memory, {free >> 20} MiB free"}
    v_host, t_host = synthetic_shard(Bg, D, rank * Bl, (rank + 1) * Bl, seed=1)
    v_dev, t_dev = v_host.to(dev), t_host.to(dev)
    del v_host, t_host
    crit = M.CrossCLR_onlyIntraModality(TAU, W, process_group=group, exchange=EXCHANGE[0]).to(dev)

    def step():
        v, t = v_dev.detach().requires_grad_(), t_dev.detach().requires_grad_()
        loss = crit(v, t)
        loss.backward()
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    loss = None
    for a_, b_ in evs:
        flush.zero_()
        a_.record()
        loss = step()
        b_.record()
    torch.cuda.synchronize()
    per_step = sorted(a_.elapsed_time(b_) for a_, b_ in evs)
    tt = torch.tensor([per_step[len(per_step) // 2]], dtype=torch.float64, device=dev)      # median step, max over ranks
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = float(tt.item())
    kt = kernel_times_eager(crit, v_dev, t_dev, 2, flush, NAT, torch)
    iso = isolated_kernel_ms(M, NAT, torch, crit, v_dev, t_dev, group, flush, 3)
    _, sustained, _ = measured_peaks()
    alg = 14.0 * Bg * Bg * D / world
    name_bwd = bwd_kernel_from_library(M, NAT, world, rank, Bl, D)
    out = {"workload": wl["desc"], "B_global": Bg, "D": D, "ms_per_step": ms, "pairs_per_s": Bg / (ms * 1e-3), "steps": steps,
           "ms_per_step_is": "median of the timed eager steps (device events), max over ranks", "ms_steps_this_rank": per_step,
           "loss": float(loss.item()), "bwd_kernel": name_bwd,
           "kernel_ms_per_step": {k: v[0] / 2 for k, v in kt.items()},
           "step_tflops_alg_per_gpu": alg / (ms * 1e-3) / 1e12, "step_frac_of_sustained": alg / (ms * 1e-3) / 1e12 / sustained,
           "bwd_frac_of_burst": None}
    burst, _, _ = measured_peaks()
    bt = torch.tensor([iso["bwd"]["mean_ms"], iso["fwd"]["mean_ms"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(bt, op=dist.ReduceOp.MAX)
    out["bwd_kernel_ms"], out["fwd_kernel_ms"] = float(bt[0]), float(bt[1])
    out["bwd_frac_of_burst"] = 8.0 * Bg * Bg * D / world / (float(bt[0]) * 1e-3) / 1e12 / burst
    del v_dev, t_dev
    torch.cuda.empty_cache()
    return out


def run_shape_variant(M, NAT, torch, dev, flush, B, D, dtype, what, steps=20):
    """A variant of the headline shape on ONE GPU -- ragged B / D, or fp32 features -- as the same one-graph-launch step
    (`HostFedCrossCLR`, device-resident inputs): ms/step, the path the library chose, and parity of the eager module against
    the CPU oracle on 64 sampled rows (fp32 gradients; checker use of oracle/ only)."""
    import ctypes
    import numpy as np
    from oracle import crossclr_oracle as O
    g = torch.Generator().manual_seed(B + D)
    v = torch.randn(B, D, generator=g).to(dtype)
    t = (v.float() + 2.0 * torch.randn(B, D, generator=g)).to(dtype)
    crit = M.CrossCLR_onlyIntraModality(TAU, W).to(dev)
    pipe = M.HostFedCrossCLR(crit, B, D, dtype=dtype, device=dev, feed="device")
    for s_ in range(2):
        pipe.video[s_].detach().copy_(v.to(dev))
        pipe.text[s_].detach().copy_(t.to(dev))
    for _ in range(3):
        pipe.step()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a_, b_ in evs:
        flush.zero_()
        a_.record()
        pipe.step()
        b_.record()
    torch.cuda.synchronize()
    ts = sorted(a_.elapsed_time(b_) for a_, b_ in evs)
    prob = NAT.Problem(2, B, D, 0, 2 * B, TAU, W)
    lib = M.load_native()
    code = lib.crossclr_choose_path(ctypes.byref(prob), {torch.float32: NAT.F32, torch.float16: NAT.F16, torch.bfloat16: NAT.BF16}[dtype], 0)
    rows = np.arange(0, B, max(1, B // 64))[:64]
    rl, rdv, rdt = O.loss_and_grads(v.float().numpy(), t.float().numpy(), TAU, W, rows=rows, row_block=2048)
    vd = v.float().to(dev).requires_grad_()       # fp32 tensors of the same values: fp32 gradients out
    td = t.float().to(dev).requires_grad_()
    loss = crit(vd, td)
    loss.backward()
    dv, dt = vd.grad.double().cpu().numpy()[rows], td.grad.double().cpu().numpy()[rows]
    del pipe
    return {"what": what, "B": B, "D": D, "in_dtype": str(dtype).replace("torch.", ""), "ms_per_step": ts[len(ts) // 2],
            "ms_per_step_is": f"median of {steps} one-graph-launch steps (pack + forward + backward), device events, L2 flushed",
            "path": {NAT.PATH_SIMT: "simt", NAT.PATH_TC: "tc", NAT.PATH_TC_SPLIT: "tc_split"}.get(code, str(code)),
            "bwd_kernel": lib.crossclr_bwd_kernel_name(ctypes.byref(prob), code).decode() if code > 0 else "",
            "parity": {"loss_rel": abs(float(loss.item()) - rl) / abs(rl), "dv_rel": float(np.linalg.norm(dv - rdv) / np.linalg.norm(rdv)),
                       "dt_rel": float(np.linalg.norm(dt - rdt) / np.linalg.norm(rdt)), "rows_checked": int(len(rows)),
                       "inputs": "the timed values as fp32 tensors through the eager module (fp32 inputs take the fp32 rule of `auto`)"}}


def run_next_rows(M, NAT, torch, dev, flush, reps=10):
    """SURVEY.md section 8 rows f1 / f4 beside the headline, on ONE GPU: `MaxMargin_coot` (trainer/loss.py:17-41) forward +
    backward and the retrieval ranks over the same score tiles, bf16 embeddings, device events, L2 flushed before each call.
    Algorithmic FLOPs: forward one B x B x D product (2 B^2 D), backward two (4 B^2 D).  Parity against the CPU oracle at
    B = 4096 (checker use of oracle/ only) and the reference's own CPU forward + autograd backward timed beside it."""
    import types
    import numpy as np
    from oracle.maxmargin_oracle import maxmargin_loss_and_grads
    from oracle.retrieval_oracle import retrieval_ranks as oracle_ranks
    burst, _, _ = measured_peaks()
    lib = M.load_native()
    out = {}
    for B, D in ((4096, 512), (16384, 512)):
        g = torch.Generator().manual_seed(B)
        im = (torch.randn(B, D, generator=g) / D ** 0.5).to(torch.bfloat16)
        s = (0.15 * im.float() + torch.randn(B, D, generator=g) / D ** 0.5).to(torch.bfloat16)
        a = im.to(dev).requires_grad_()
        b = s.to(dev).requires_grad_()
        crit = M.MaxMargin_coot(True, 0.1)
        go = torch.ones((), dtype=torch.bfloat16, device=dev)

        def timed(fn):
            for _ in range(3):
                fn()
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
            for e0, e1 in evs:
                flush.zero_()
                e0.record()
                fn()
                e1.record()
            torch.cuda.synchronize()
            ts = sorted(e0.elapsed_time(e1) for e0, e1 in evs)
            return ts[len(ts) // 2]

        state = {}

        def fwd():
            state["loss"] = crit(a, b)

        def bwd():
            a.grad = b.grad = None
            state["loss"].backward(go, retain_graph=True)

        def step():
            a.grad = b.grad = None
            crit(a, b).backward()

        t_fwd = timed(fwd)
        t_bwd = timed(bwd)
        t_step = timed(step)
        t_rank = timed(lambda: M.retrieval_ranks(a.detach(), b.detach()))
        code = NAT.BF16
        rec = {"B": B, "D": D, "in_dtype": "bfloat16", "margin": 0.1,
               "kernel": lib.crossclr_maxmargin_kernel_name(a.data_ptr(), b.data_ptr(), code, a.stride(0), b.stride(0), B, D).decode(),
               "fwd_ms": t_fwd, "bwd_ms": t_bwd, "step_ms": t_step, "pairs_per_s": B / (t_step * 1e-3),
               "fwd_tflops_alg": 2.0 * B * B * D / (t_fwd * 1e-3) / 1e12, "bwd_tflops_alg": 4.0 * B * B * D / (t_bwd * 1e-3) / 1e12,
               "step_frac_of_peak": 6.0 * B * B * D / (t_step * 1e-3) / 1e12 / burst,
               "retrieval_ranks_ms": t_rank, "retrieval_tflops_alg": 2.0 * B * B * D / (t_rank * 1e-3) / 1e12,
               "timing": f"median of {reps} eager module calls (launches + autograd), device events, L2 flushed"}
        if B == 4096:
            rl, rda, rdb = maxmargin_loss_and_grads(im.float().numpy(), s.float().numpy(), 0.1)
            af = im.float().to(dev).bfloat16()
            ws_n = int(lib.crossclr_maxmargin_workspace_bytes(B, D, code))
            ws = torch.empty(ws_n, dtype=torch.uint8, device=dev)
            loss64 = torch.empty((), dtype=torch.float64, device=dev)
            da = torch.empty(B, D, dtype=torch.float32, device=dev)
            db = torch.empty(B, D, dtype=torch.float32, device=dev)
            bd = b.detach()
            st = torch.cuda.current_stream().cuda_stream
            NAT.check(lib.crossclr_maxmargin_fwd(af.data_ptr(), bd.data_ptr(), code, D, D, B, D, 0.1, ws.data_ptr(), ws_n,
                                                 loss64.data_ptr(), st), "crossclr_maxmargin_fwd")
            NAT.check(lib.crossclr_maxmargin_bwd(af.data_ptr(), bd.data_ptr(), code, D, D, B, D, 0.1, ws.data_ptr(), ws_n, None,
                                                 da.data_ptr(), D, db.data_ptr(), D, NAT.F32, st), "crossclr_maxmargin_bwd")
            torch.cuda.synchronize()
            # the same values as fp32 tensors: fp16 hi + lo operands, K = 3 D
            a32 = im.float().to(dev).requires_grad_()
            b32 = s.float().to(dev).requires_grad_()

            def step32():
                a32.grad = b32.grad = None
                crit(a32, b32).backward()

            rec["fp32_inputs"] = {"step_ms": timed(step32), "parity_dim_rel": float(np.linalg.norm(a32.grad.double().cpu().numpy() - rda)
                                                                                     / np.linalg.norm(rda)),
                                  "kernel": lib.crossclr_maxmargin_kernel_name(a32.data_ptr(), b32.data_ptr(), NAT.F32, D, D, B, D).decode(),
                                  "what": "fp32 embeddings staged as fp16 hi + lo pairs (per-tensor power-of-two scale)"}
            ra, rb, _ = oracle_ranks(im.float().numpy(), s.float().numpy())
            ga, gb = M.retrieval_ranks(a.detach(), bd)
            rec["parity"] = {"loss_rel": abs(loss64.item() - rl) / abs(rl),
                             "dim_rel": float(np.linalg.norm(da.double().cpu().numpy() - rda) / np.linalg.norm(rda)),
                             "ds_rel": float(np.linalg.norm(db.double().cpu().numpy() - rdb) / np.linalg.norm(rdb)),
                             "rank_mismatches": int((ga.cpu().numpy() != ra).sum() + (gb.cpu().numpy() != rb).sum()),
                             "recall_at_10": float((ga < 10).double().mean()), "oracle_recall_at_10": float((ra < 10).mean()),
                             "against": "oracle/maxmargin_oracle.py, oracle/retrieval_oracle.py (float64), fp32 gradients out of the C ABI"}
            ref_path = os.path.join(ROOT, "baseline", "_ref", "trainer", "loss.py")
            if os.path.exists(ref_path):
                import importlib.util
                spec = importlib.util.spec_from_file_location("_crossclr_reference_loss_mm", ref_path)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                me = types.SimpleNamespace(margin=0.1, sim=mod.cosine_sim, use_cuda=False)   # the ctor is a NameError (:24)
                x = im.float().requires_grad_()
                y = s.float().requires_grad_()
                best = None
                for _ in range(3):
                    x.grad = y.grad = None
                    t0 = time.perf_counter()
                    mod.MaxMargin_coot.forward(me, x, y).backward()
                    dt_ = (time.perf_counter() - t0) * 1e3
                    best = dt_ if best is None else min(best, dt_)
                rec["cpu_reference"] = {"step_ms": best, "cores": torch.get_num_threads(), "dtype": "fp32",
                                        "what": "unbound MaxMargin_coot.forward of baseline/_ref + autograd backward, best of 3"}
        out[f"maxmargin_b{B}"] = rec
        del a, b, crit
    return out


def bwd_kernel_from_library(M, NAT, world, rank, Bl, D):
    import ctypes
    lib = M.load_native()
    prob = NAT.Problem(2 * world, Bl, D, 2 * rank * Bl, 2 * Bl, TAU, W)
    code = lib.crossclr_choose_path(ctypes.byref(prob), NAT.BF16, 0)
    return lib.crossclr_bwd_kernel_name(ctypes.byref(prob), code).decode() if code > 0 else ""


def parity_vs_oracle(crit, v_dev, t_dev, v_all, t_all, lo, torch, dist, world, dev, max_rows=256):
    """This rank's loss and gradients (eager module, fp32 gradients out of bf16-representable fp32 inputs) against the CPU
    oracle on the concatenated batch, on a sample of the rank's own rows; MAX over ranks.  Checker use of oracle/ only."""
    import numpy as np
    from oracle import crossclr_oracle as O
    v = v_dev.float().detach().requires_grad_()
    t = t_dev.float().detach().requires_grad_()
    loss = crit(v, t)
    loss.backward()
    torch.cuda.synchronize()
    Bl = v.shape[0]
    sel = np.unique(np.linspace(0, Bl - 1, min(max_rows, Bl)).astype(np.int64))
    rl, rdv, rdt = O.loss_and_grads(v_all, t_all, TAU, W, rows=lo + sel, row_block=1024)
    dv = v.grad.double().cpu().numpy()[sel]
    dt = t.grad.double().cpu().numpy()[sel]
    rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
    vals = torch.tensor([abs(loss.item() - rl) / abs(rl), rel(dv, rdv), rel(dt, rdt)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    return {"loss_rel": float(vals[0]), "dv_rel": float(vals[1]), "dt_rel": float(vals[2]), "rows_checked_per_rank": int(len(sel)),
            "checker": "oracle/crossclr_oracle.py (fp64 restatement pinned by reference-generated goldens) on the concatenated batch",
            "tolerance": 1e-3, "inputs": "bf16-representable fp32 features, fp32 gradients"}


def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"          # NCCL prints its version banner on stdout: the bench prints ONE line
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD

    import crossmodal_contrastive_learning_b200 as M
    from crossmodal_contrastive_learning_b200 import _native as NAT
    M.load_native()

    wl = WORKLOADS[args.workload]
    Bg, D = wl["B"], wl["D"]
    assert Bg % world == 0
    Bl = Bg // world
    lo, hi = rank * Bl, (rank + 1) * Bl
    v_host, t_host = synthetic_shard(Bg, D, lo, hi)

    # how the ranks exchange their row shards: peer stores over NVLink (csrc/peer.cu) where CUDA IPC works on this box,
    # else the two NCCL all-gathers; every rank takes the same decision
    exchange, exchange_note = "nccl", None
    if world > 1 and args.exchange in ("auto", "peer"):
        ok = 1
        try:
            from crossmodal_contrastive_learning_b200 import peer as PEER
            PEER.plan_for(group, dev, 4096, 4096)
        except Exception as exc:
            ok, exchange_note = 0, f"peer exchange unavailable ({type(exc).__name__}: {str(exc)[:120]})"
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()):
            exchange = "peer"
        elif args.exchange == "peer":
            raise SystemExit(exchange_note or "peer exchange unavailable on another rank")
    EXCHANGE[0] = exchange
    crit = M.CrossCLR_onlyIntraModality(TAU, W, process_group=group, exchange=exchange).to(dev)
    v_dev = v_host.to(dev)
    t_dev = t_host.to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)      # > 126 MB L2
    use_graph = not args.no_graph
    graph_note = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_cap0 = M.launch_count()
    resident, pipe, kernels_per_step = None, None, None
    if use_graph:
        # The whole step is ONE graph launch (HostFedCrossCLR, the package's public capture API): `resident` = forward +
        # backward on device-resident inputs; `pipe` = the same plus, on a second captured stream, the H2D copy of the next
        # step's inputs from pinned memory and the D2H copy of the loss.
        try:
            resident = M.HostFedCrossCLR(crit, Bl, D, dtype=torch.bfloat16, device=dev, feed="device")
            n_cap1 = M.launch_count()
            kernels_per_step = (n_cap1 - n_cap0) // (2 * 4)      # 3 warm-up steps + 1 capture, for each of two buffer sets
            pipe = M.HostFedCrossCLR(crit, Bl, D, dtype=torch.bfloat16, device=dev, feed="host")
            for s in range(2):
                resident.video[s].detach().copy_(v_dev)
                resident.text[s].detach().copy_(t_dev)
                pipe.host_video[s].copy_(v_host)
                pipe.host_text[s].copy_(t_host)
            ok = 1
        except Exception as exc:          # e.g. a collective that cannot be captured on this stack: time the eager module
            graph_note = f"capture failed ({type(exc).__name__}: {str(exc)[:120]})"
            ok = 0
        if world > 1:                     # all ranks must agree on the launch mode
            flag = torch.tensor([ok], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = int(flag.item())
        if not ok:
            use_graph, resident, pipe = False, None, None

    def step_resident():
        if use_graph:
            resident.step()
            return None
        v = v_dev.detach().requires_grad_()
        t = t_dev.detach().requires_grad_()
        loss = crit(v, t)
        loss.backward()
        return loss

    def timed_resident(steps):
        """Sum of per-step device times (CUDA events on the launching stream); L2 flushed between steps."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for a_, b_ in evs:
            flush.zero_()
            a_.record()
            step_resident()
            b_.record()
        barrier()
        total = sum(a_.elapsed_time(b_) for a_, b_ in evs)
        tt = torch.tensor([total], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    loss_host_eager = torch.empty((), dtype=torch.float64).pin_memory()

    def timed_e2e(steps):
        """Device time of `steps` end-to-end steps through the public API: every step's inputs cross PCIe from pinned host
        memory and its loss is read back, all inside the timed region.  Graph mode: one launch per step, the upload of step
        i+1 captured beside the kernels of step i; eager mode: serial copy, step, read-back."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        if use_graph:
            pipe.prime()
            for _ in range(steps):
                pipe.step()
            pipe.drain()
        else:
            for _ in range(steps):
                v_dev.copy_(v_host, non_blocking=True)
                t_dev.copy_(t_host, non_blocking=True)
                loss = step_resident()
                loss_host_eager.copy_(loss.detach(), non_blocking=True)
        e1.record()
        barrier()
        total = e0.elapsed_time(e1)
        tt = torch.tensor([total], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    warm = max(args.warmup, 3)
    clk = ClockSampler(local_rank) if rank == 0 else None   # started early: nvidia-smi takes a moment to produce its first sample
    for _ in range(warm):
        step_resident()
    timed_e2e(warm)
    torch.cuda.synchronize()

    n0 = M.launch_count()
    if clk is not None:
        clk.begin()                       # the sampling window spans both timed regions (value, then e2e)
    total_ms = timed_resident(args.steps)
    launches = (kernels_per_step * args.steps) if use_graph else (M.launch_count() - n0)
    e2e_ms = timed_e2e(args.steps)
    # nvidia-smi delivers a sample every ~15 ms whatever -lms asks for, and the two timed regions together last a few tens
    # of ms at this step time: keep the same step running (untimed) until the window is long enough to hold samples
    t_probe = time.time()
    while time.time() - t_probe < 0.4:
        for _ in range(50):
            step_resident()
        torch.cuda.synchronize()
    if clk is not None:
        clk.end()
    loss_val = float(pipe.loss_host[(pipe._k - 1) & 1]) if use_graph else float(loss_host_eager)

    # per-kernel device times: the dominant kernel alone (graph replay), and every kernel family of eager steps
    iso = isolated_kernel_ms(M, NAT, torch, crit, v_dev, t_dev, group, flush, args.steps)
    ktimes = kernel_times_eager(crit, v_dev, t_dev, args.steps, flush, NAT, torch)
    iso_t = torch.tensor([iso["bwd"]["mean_ms"], iso["fwd"]["mean_ms"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(iso_t, op=dist.ReduceOp.MAX)
    barrier()

    parity = None
    if not args.no_parity:
        # every rank regenerates the global batch (same seed) for the checker
        va, ta = synthetic_shard(Bg, D, 0, Bg)
        parity = parity_vs_oracle(crit, v_dev, t_dev, va.float().numpy(), ta.float().numpy(), lo, torch, dist, world, dev)
        del va, ta

    # (before the large workloads: the seconds after c5's 0.3 s steps run at reduced clocks under the power cap)
    shapes = {}
    if not args.no_extras and world == 1 and args.workload in (None, "c2"):
        # f3 coverage next to the headline: the same step on ragged B / D and on fp32 features (reference: any [B, D], fp32 GEMMs)
        for key, (B_, D_, dt_, what) in {
                "c2_ragged": (4000, 500, torch.bfloat16, "ragged neighbour of c2 on the tensor-core path (zero-padded layout)"),
                "c2_fp32": (4096, 512, torch.float32, "c2 with fp32 features: fp16 hi + lo operand pairs, K = 3 D"),
                "c2_fp32_ragged": (4000, 500, torch.float32, "both")}.items():
            try:
                shapes[key] = run_shape_variant(M, NAT, torch, dev, flush, B_, D_, dt_, what)
            except Exception as exc:
                shapes[key] = {"error": f"{type(exc).__name__}: {str(exc)[:200]}"}

    next_rows = None
    if not args.no_extras and world == 1 and args.workload in (None, "c2"):
        try:
            next_rows = run_next_rows(M, NAT, torch, dev, flush)
        except Exception as exc:
            next_rows = {"error": f"{type(exc).__name__}: {str(exc)[:300]}"}

    extras = {}
    if not args.no_extras:
        for name in (["c3"] if world == 1 else []) + ["c4", "c5"]:
            if name == args.workload:
                continue
            try:
                extras[name] = run_extra_workload(name, M, NAT, torch, dist, world, rank, dev, group, flush)
            except Exception as exc:
                extras[name] = {"error": f"{type(exc).__name__}: {str(exc)[:200]}"}
            barrier()

    cb = None
    if rank == 0 and not args.no_cpu_baseline:
        cb, _, _ = time_reference(Bg, D, steps=3, warmup=1, max_seconds=30.0)

    if rank == 0:
        burst, sustained, src = measured_peaks()
        ms_step = total_ms / args.steps
        # dominant kernel = backward similarity/gradient kernel: 4 products = 8 B^2 D algorithmic FLOPs per step,
        # this rank computes 1/world of them (DESIGN.md "Roofline accounting")
        bwd_ms, bwd_n = float(iso_t[0]), iso["bwd"]["launches"]
        fwd_ms = float(iso_t[1])
        alg_bwd = 8.0 * Bg * Bg * D / world
        alg_fwd = 6.0 * Bg * Bg * D / world
        achieved = alg_bwd / (bwd_ms * 1e-3) / 1e12 if bwd_ms > 0 else None
        kname = bwd_kernel_from_library(M, NAT, world, rank, Bl, D)
        traffic = ncu_traffic(args.workload, kname)
        line = {
            "metric": METRIC, "value": Bg / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f16 (tensor-core operands; f32 accumulate; bf16 features in, bf16 grads out)",
            "data": "synthetic",
            "config": {"workload": f"{args.workload}: {wl['desc']}", "B_global": Bg, "B_per_gpu": Bl, "D": D,
                       "temperature": TAU, "negative_weight": W,
                       "l2": "value: flushed between steps (256 MiB write); e2e: each step's inputs arrive by H2D copy",
                       "launch": ("one CUDA graph launch per step (HostFedCrossCLR: pack + forward + backward"
                                  + ((", peer-store exchanges captured" if exchange == "peer" else ", NCCL all-gathers captured")
                                     if world > 1 else "") + ")" if use_graph
                                  else "eager module calls" + (f"; {graph_note}" if graph_note else "")),
                       "parallelism": (f"row-sharded x{world}, " + ("row shards and row statistics stored into every rank's buffers over "
                                       "NVLink peer memory (crossclr_peer_exchange: stores + flag barrier in one kernel, twice per step)"
                                       if exchange == "peer" else "NCCL all-gather of features + row stats")
                                       + (f"; {exchange_note}" if exchange_note else "")) if world > 1 else "single GPU",
                       "exchange": exchange if world > 1 else None,
                       "nvlink_bytes_stored_per_rank_per_step": ((world - 1) * (2 * Bl * (D + 64) * 2 + 2 * Bl * 8)
                                                                 if (world > 1 and exchange == "peer") else None),
                       "loss": loss_val},
            "e2e": {"value": Bg / (e2e_ms / args.steps * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": 2 * Bl * D * 2, "d2h_bytes_per_step": 8,
                    "pipeline": ("one graph launch per step (pack + forward + backward); the H2D copy of step i+1 and the D2H copy of "
                                 "step i-1's loss run on a copy stream beside step i's kernels" if use_graph else "serial")},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": kname,
                         "achieved": achieved, "peak": burst,
                         "unit": "TFLOP/s", "frac": (achieved / burst if achieved else None), "traffic": traffic,
                         "peak_source": f"{src} bf16_tflops (burst; kernel timed alone with CUDA events)",
                         "algorithmic_flops_per_launch": alg_bwd, "avg_launch_ms": bwd_ms, "launches_timed": bwd_n,
                         "timing": "mean over launches_timed replays of a CUDA graph holding ONLY crossclr_bwd_accumulate (control-word "
                                   "memset + the kernel), CUDA events on the launching stream around each replay, L2 flushed before "
                                   "each, max over ranks",
                         "median_launch_ms": iso["bwd"]["median_ms"], "min_launch_ms": iso["bwd"]["min_ms"],
                         "fwd_kernel": {"avg_launch_ms": fwd_ms, "algorithmic_flops_per_launch": alg_fwd,
                                        "achieved": (alg_fwd / (fwd_ms * 1e-3) / 1e12 if fwd_ms > 0 else None)},
                         "step": {"algorithmic_flops": 14.0 * Bg * Bg * D / world,
                                  "achieved": 14.0 * Bg * Bg * D / world / (ms_step * 1e-3) / 1e12,
                                  "frac_of_sustained": 14.0 * Bg * Bg * D / world / (ms_step * 1e-3) / 1e12 / sustained},
                         "kernel_ms_per_step": {k: ms / args.steps for k, (ms, n) in ktimes.items()},
                         "kernel_ms_per_step_is": "library cudaEvents around each launch of EAGER steps (every bracket carries ~10 us "
                                                  "of launch latency on a starved stream: shares, not absolutes)"},
            "clocks": clk.summary(),
        }
        if parity is not None:
            line["parity"] = parity
        if extras:
            line["workloads"] = extras
        if shapes:
            line["shapes"] = shapes
        if next_rows:
            line["next_rows"] = next_rows
        if cb is not None:
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    if world > 1:
        # Captured graphs hold NCCL kernels; tearing the communicator down under them hung at exit on this stack.  The
        # result is printed: drop the graphs, make sure every rank is done, and leave without the teardown.
        resident = pipe = None
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer", "nccl"],
                    help="N > 1: how ranks exchange row shards (auto: NVLink peer stores where CUDA IPC works, else NCCL)")
    ap.add_argument("--no-graph", action="store_true", help="time eager module calls instead of CUDA-graph replays")
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE configs (c3 / c4 / c5 records)")
    ap.add_argument("--no-parity", action="store_true", help="skip the per-rank oracle check of the headline workload")
    ap.add_argument("--shape", default=None, help="experiment override B,D (not a BASELINE config)")
    args = ap.parse_args()
    if args.workload is None:
        args.workload = "c2"
    if args.shape:
        B, D = (int(x) for x in args.shape.split(","))
        WORKLOADS["x"] = dict(B=B, D=D, desc=f"experiment B={B} D={D} bf16")
        args.workload = "x"
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
