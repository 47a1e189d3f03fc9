"""Graph-replayed times of the single-rank forward and backward entry points for shapes and input dtypes (ragged shapes and
the fp32 hi + lo path beside their aligned / bf16 neighbours):

    python scripts/gpu_step_time.py [B D dtype]...      dtype in bf16 | f16 | f32 | f32tc (fp32 inputs, path forced to tc)
"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import crossmodal_contrastive_learning_b200 as M  # noqa: E402
from crossmodal_contrastive_learning_b200 import _native as N, loss as L  # noqa: E402

DT = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32, "f32tc": torch.float32}


def timed(fn, reps=15):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        fn()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); gr.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def one(B, D, dname):
    ops, lib = L._ops(), M.load_native()
    dt_in = DT[dname]
    g = torch.Generator().manual_seed(0)
    v = torch.randn(B, D, generator=g).to(dt_in).cuda()
    t = (v.float().cpu() + 2.0 * torch.randn(B, D, generator=g)).to(dt_in).cuda()
    prob = N.Problem(2, B, D, 0, 2 * B, 0.03, 0.8)
    code, fdt, pitch = ops.plan(prob, dt_in, False, N.PATH_TC if dname == "f32tc" else None)
    S = ops.seg_rows(code, B)
    feat = torch.empty((2, S, pitch), dtype=fdt, device="cuda")
    rn = torch.empty(2 * B, dtype=torch.float32, device="cuda")
    stats = torch.empty((2 * S, 2), dtype=torch.float32, device="cuda"); coef = torch.empty_like(stats)
    scal = torch.empty(4, dtype=torch.float32, device="cuda"); loss = torch.empty((), dtype=torch.float64, device="cuda")
    go = torch.ones((), dtype=torch.float64, device="cuda")
    dv = torch.empty((B, D), dtype=dt_in, device="cuda"); dtt = torch.empty_like(dv)
    ws_bytes = int(lib.crossclr_workspace_bytes(ctypes.byref(prob), code))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")

    def fwd():
        ops.forward_single(prob, code, v, t, feat, rn, stats, coef, scal, loss)

    def bwd():
        N.check(lib.crossclr_bwd(ctypes.byref(prob), code, L._ptr(feat), L._ptr(rn), L._ptr(coef), L._ptr(scal), L._ptr(go), 1.0,
                                 L._ptr(dv), dv.stride(0), L._ptr(dtt), dtt.stride(0), L._DTYPE_CODE[dv.dtype], L._ptr(ws), ws_bytes,
                                 L._stream()), "bwd")
    tf = timed(fwd)
    tb = timed(bwd)
    name = lib.crossclr_bwd_kernel_name(ctypes.byref(prob), code).decode()
    print(f"B={B} D={D} {dname}: path {code} ({name}) forward(pack+fwd+finalize) {tf:.1f} us  backward(+grad_finish) {tb:.1f} us  "
          f"step {tf + tb:.1f} us  loss {loss.item():.6f}", flush=True)


def main():
    args = sys.argv[1:]
    cases = [(int(args[i]), int(args[i + 1]), args[i + 2]) for i in range(0, len(args), 3)] or \
        [(4096, 512, "bf16"), (4000, 500, "bf16"), (4096, 512, "f32"), (4096, 512, "f32tc"), (4000, 500, "f32"),
         (8192, 512, "bf16"), (8000, 500, "bf16"), (2048, 1024, "bf16"), (2000, 1000, "bf16"), (1024, 256, "bf16"), (1000, 250, "bf16")]
    for B, D, dn in cases:
        one(B, D, dn)


if __name__ == "__main__":
    main()
