import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import crossmodal_contrastive_learning_b200 as M
from crossmodal_contrastive_learning_b200 import _native as N, loss as L
what = sys.argv[1]
if what == "mixed":
    lib = M.load_native()
    rng = np.random.default_rng(1)
    a = rng.standard_normal((128, 128)).astype(np.float16)
    bf = torch.from_numpy(rng.standard_normal((128, 128)).astype(np.float32)).to(torch.bfloat16)
    b = bf.view(torch.int16).numpy().copy()
    out = np.zeros((128, 128), dtype=np.float32)
    rc = lib.crossclr_selftest(6, a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), 128, 128)
    ref = a.astype(np.float32) @ bf.float().numpy().T
    print("mixed f16 x bf16 selftest rc", rc, "max err", np.abs(out - ref).max(), lib.crossclr_last_error())
else:
    B, D, dt = int(sys.argv[2]), int(sys.argv[3]), {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}[sys.argv[4]]
    ops = L._ops()
    g = torch.Generator().manual_seed(0)
    v = torch.randn(B, D, generator=g).to(dt).cuda(); t = torch.randn(B, D, generator=g).to(dt).cuda()
    prob = N.Problem(2, B, D, 0, 2 * B, 0.03, 0.8)
    code, fdt, pitch = ops.plan(prob, dt, False)
    feat = torch.empty((2, B, pitch), dtype=fdt, device="cuda"); rn = torch.empty(2 * B, device="cuda")
    stats = torch.empty((2 * B, 2), device="cuda"); coef = torch.empty_like(stats); scal = torch.empty(4, device="cuda")
    loss = torch.empty((), dtype=torch.float64, device="cuda"); go = torch.ones((), dtype=torch.float64, device="cuda")
    dv = torch.empty((B, D), device="cuda"); dtt = torch.empty((B, D), device="cuda")
    ops.pack2(v, t, feat, rn); torch.cuda.synchronize(); print("pack ok; q sample", feat.view(torch.int16)[0, 0, D:D + 2].view(torch.float32).item() if fdt != torch.float32 else None)
    if what in ("fwd", "all"):
        ops.forward_single(prob, code, v, t, feat, rn, stats, coef, scal, loss); torch.cuda.synchronize(); print("fwd ok loss", loss.item())
    if what in ("all",):
        print("bwd kernel", M.load_native().crossclr_bwd_kernel_name(ctypes.byref(prob), code))
        ops.bwd(prob, code, feat, rn, coef, scal, go, 1.0, dv, dtt); torch.cuda.synchronize(); print("bwd ok", dv.norm().item())
