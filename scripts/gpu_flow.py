"""GPU-box check of the dataflow backward (csrc/flow_kernels.cu): building-block self-test, parity against the oracle for
the symmetric and the row-band schedules (ranks emulated on one GPU through the C ABI), and kernel times.

    python scripts/gpu_flow.py [quick|full]
"""
import ctypes
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import crossmodal_contrastive_learning_b200 as M  # noqa: E402
from crossmodal_contrastive_learning_b200 import _native as N  # noqa: E402
from crossmodal_contrastive_learning_b200 import loss as L  # noqa: E402
from oracle import crossclr_oracle as O  # noqa: E402


def selftest(variant, n, k):
    lib = M.load_native()
    rng = np.random.default_rng(variant * 100 + n + k)
    m = 256 if variant == 4 else 128
    a = rng.standard_normal((m, k)).astype(np.float16)
    b = (rng.standard_normal((k, n)) if variant == 1 else rng.standard_normal((n, k))).astype(np.float16)
    out = np.zeros((m, n), dtype=np.float32)
    rc = lib.crossclr_selftest(variant, a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p),
                               out.ctypes.data_as(ctypes.c_void_p), n, k)
    ref = a.astype(np.float32) @ (b.astype(np.float32) if variant == 1 else b.astype(np.float32).T)
    return rc, float(np.abs(out - ref).max())


def run_ranks(v, t, world, tau=0.03, w=0.8, path="tc", reps=1):
    """The criterion of `world` ranks emulated on one GPU through the C ABI (what each rank would launch).
    Returns loss, dv, dt (global), and per-family kernel ms of the last repetition."""
    ops = L._ops()
    Bg, D = v.shape
    B = Bg // world
    dev = v.device
    probs = [N.Problem(2 * world, B, D, 2 * r * B, 2 * B, tau, w) for r in range(world)]
    code, fdt, pitch = ops.plan(probs[0], v.dtype, path == "simt")
    feat_all = torch.empty((2 * world, B, pitch), dtype=fdt, device=dev)
    rnorm = torch.empty((world, 2 * B), dtype=torch.float32, device=dev)
    stats = torch.empty((2 * world * B, 2), dtype=torch.float32, device=dev)
    coef = torch.empty_like(stats)
    scal = torch.empty(4, dtype=torch.float32, device=dev)
    loss = torch.empty((), dtype=torch.float64, device=dev)
    go = torch.ones((), dtype=torch.float64, device=dev)
    dv = torch.empty((Bg, D), dtype=torch.float32, device=dev)      # fp32 gradients whatever the input dtype
    dt = torch.empty((Bg, D), dtype=torch.float32, device=dev)
    for rep in range(reps):
        if rep == reps - 1:
            torch.cuda.synchronize()
            N.timing_read()
            N.timing_enable(True)
        for r in range(world):
            ops.pack2(v[r * B:(r + 1) * B], t[r * B:(r + 1) * B], feat_all[2 * r:2 * r + 2], rnorm[r])
        for r in range(world):
            ops.fwd(probs[r], code, feat_all, stats)
        ops.finalize(probs[0], code, stats, coef, loss, scal)
        for r in range(world):
            ops.bwd(probs[r], code, feat_all, rnorm[r], coef, scal, go, 1.0, dv[r * B:(r + 1) * B], dt[r * B:(r + 1) * B])
    torch.cuda.synchronize()
    N.timing_enable(False)
    kt = {k: (ms / max(n, 1), n) for k, (ms, n) in N.timing_read().items()}
    name = M.load_native().crossclr_bwd_kernel_name(ctypes.byref(probs[0]), code).decode()
    return loss.item(), dv, dt, kt, name


def check(B, D, world, aligned=2.0, sample=None, reps=3):
    g = torch.Generator().manual_seed(B + D + world)
    v = torch.randn(B, D, generator=g).to(torch.bfloat16).float()
    t = (v + aligned * torch.randn(B, D, generator=g)).to(torch.bfloat16).float() if aligned else \
        torch.randn(B, D, generator=g).to(torch.bfloat16).float()
    t0 = time.time()
    loss, dv, dt, kt, name = run_ranks(v.to("cuda", IN_DTYPE), t.to("cuda", IN_DTYPE), world, reps=reps)
    rows = None if sample is None else np.arange(0, B, B // sample) + 1
    rl, rdv, rdt = O.loss_and_grads(v.numpy(), t.numpy(), 0.03, 0.8, rows=rows, row_block=2048)
    dvn, dtn = dv.double().cpu().numpy(), dt.double().cpu().numpy()
    if rows is not None:
        dvn, dtn = dvn[rows], dtn[rows]
    ev = np.linalg.norm(dvn - rdv) / np.linalg.norm(rdv)
    et = np.linalg.norm(dtn - rdt) / np.linalg.norm(rdt)
    mv = np.abs(dvn - rdv).max() / np.abs(rdv).max()
    ok = abs(loss - rl) <= 1e-3 * abs(rl) and ev <= 1e-3 and et <= 1e-3 and mv <= 1e-3
    alg = 8.0 * B * B * D / world
    print(f"B={B} D={D} world={world} {name}: loss {loss:.6f} (oracle {rl:.6f}) dv_rel {ev:.2e} dt_rel {et:.2e} max {mv:.2e} "
          f"{'OK' if ok else 'FAIL'} | bwd {kt['bwd'][0] * 1e3:.1f} us/launch ({alg / (kt['bwd'][0] * 1e-3) / 1e12:.0f} TF/s alg) "
          f"fwd {kt['fwd'][0] * 1e3:.1f} us [{time.time() - t0:.1f}s]", flush=True)
    return ok


IN_DTYPE = torch.bfloat16


def main():
    global IN_DTYPE
    mode = sys.argv[1] if len(sys.argv) > 1 else "quick"
    if len(sys.argv) > 2:
        IN_DTYPE = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}[sys.argv[2]]
    for variant, n, k in ((5, 128, 128), (5, 64, 64), (5, 256, 128), (3, 128, 128)):
        rc, err = selftest(variant, n, k)
        print(f"selftest variant {variant} n={n} k={k}: rc={rc} max err {err:.3e} {'OK' if rc == 0 and err < 1e-2 else 'FAIL'}", flush=True)
    ok = True
    cases = [(1024, 512, 1, 2.0, None), (2048, 256, 1, 0.0, None), (4096, 512, 1, 0.0, None), (4096, 512, 1, 2.0, None),
             (1536, 384, 1, 2.0, None), (2048, 512, 4, 2.0, None), (4096, 512, 8, 0.0, None), (2560, 128, 1, 2.0, None),
             (1024, 1024, 1, 2.0, None), (2048, 1024, 1, 0.0, None), (2048, 1024, 4, 2.0, None), (2048, 768, 1, 2.0, None),
             (4096, 1024, 2, 2.0, None)]
    if mode == "full":
        cases += [(8192, 512, 1, 2.0, 16), (8192, 512, 2, 0.0, 16), (16384, 512, 1, 2.0, 16)]
    for B, D, world, al, sample in cases:
        try:
            ok &= check(B, D, world, al, sample)
        except Exception as exc:  # keep going: a poisoned context shows up as failures below
            print(f"B={B} D={D} world={world}: EXCEPTION {type(exc).__name__}: {str(exc)[:300]}", flush=True)
            ok = False
            break
    print("FLOW CHECK", "OK" if ok else "FAILED")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
