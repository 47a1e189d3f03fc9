"""GPU parity tests: the CUDA path (through the C ABI, via the module mirror) against
  * the reference-generated golden vectors in tests/golden/ (outputs of the unmodified reference),
  * the CPU oracle (oracle/crossclr_oracle.py) on the same seeded inputs,
  * closed-form known-answer cases and size-independent properties at BASELINE.json sizes.

Tolerances (BASELINE.json north_star: 1e-3 relative fp32):
    |L - L_ref| <= 1e-3 |L_ref| + 1e-7,   ||g - g_ref||_F <= 1e-3 ||g_ref||_F,   max|g - g_ref| <= 1e-3 max|g_ref|
The exact-fp32 SIMT path is held to a 20x tighter bar.
"""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

TOL = 1e-3
TOL_SIMT = 5e-5
# below tau ~ 0.0073 the fp32 logits reach log2e max(1,|w|)/tau = 300..600 in magnitude: one ulp there is 3e-5..6e-5 of a probability
TOL_SMALL_TAU = 3e-4


def _mod():
    import crossmodal_contrastive_learning_b200 as M
    return M


def run_gpu(v, t, tau, w, path="auto", dtype=torch.float32, grad_out=None):
    M = _mod()
    vd = torch.as_tensor(v).to("cuda", dtype).requires_grad_()
    td = torch.as_tensor(t).to("cuda", dtype).requires_grad_()
    crit = M.CrossCLR_onlyIntraModality(tau, w, path=path)
    loss = crit(vd, td)
    assert loss.dtype == torch.float64 and loss.dim() == 0 and loss.is_cuda
    if grad_out is None:
        loss.backward()
    else:
        (loss * grad_out).backward()
    assert vd.grad.dtype == dtype and td.grad.dtype == dtype
    return loss.item(), vd.grad.double().cpu().numpy(), td.grad.double().cpu().numpy()


def check(loss, dv, dt, rloss, rdv, rdt, tol):
    assert abs(loss - rloss) <= tol * abs(rloss) + 1e-7, (loss, rloss)
    for g, r, name in ((dv, rdv, "dv"), (dt, rdt, "dt")):
        rel = np.linalg.norm(g - r) / max(np.linalg.norm(r), 1e-300)
        mx = np.abs(g - r).max() / max(np.abs(r).max(), 1e-300)
        assert rel <= tol, (name, "rel-norm", rel)
        assert mx <= tol, (name, "max", mx)


FULL = sorted(p for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
              if not os.path.basename(p).startswith(("kat_", "c1_")))
TC_SHAPED = [p for p in FULL if (lambda g: g["v"].shape[0] % 128 == 0 and g["v"].shape[1] % 64 == 0)(np.load(p))]
# the tensor-core kernels use one shift for all rows: log2e max(1,|w|)/tau <= 200 (include/crossclr_b200.h, crossclr_choose_path)
_in_tc_range = lambda g: 1.4426950408889634 * max(1.0, abs(float(g["w"]))) / float(g["tau"]) <= 200.0
TC_OK = [p for p in FULL if _in_tc_range(np.load(p))]       # any shape: the tensor-core layout zero-pads B to 128 and D to 64
SMALL_TAU = [p for p in FULL if not _in_tc_range(np.load(p))]


# ---------------------------------------------------------------------------------------------
# tcgen05 / TMA building blocks
@pytest.mark.parametrize("variant,n,k", [(0, 128, 64), (0, 128, 256), (0, 64, 128), (0, 256, 128), (1, 64, 128),
                                         (1, 64, 64), (1, 128, 128), (1, 256, 128), (2, 128, 128), (3, 128, 128), (3, 64, 64),
                                         (4, 256, 128), (4, 128, 64), (4, 64, 256)])
def test_tc_selftest(variant, n, k):
    import ctypes
    lib = _mod().load_native()
    rng = np.random.default_rng(variant * 100 + n + k)
    m = 256 if variant == 4 else 128          # variant 4: one cta_group::2 MMA stream over a CTA pair
    a = rng.standard_normal((m, k)).astype(np.float16)
    b = (rng.standard_normal((k, n)) if variant == 1 else rng.standard_normal((n, k))).astype(np.float16)
    out = np.zeros((m, n), dtype=np.float32)
    rc = lib.crossclr_selftest(variant, a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p),
                               out.ctypes.data_as(ctypes.c_void_p), n, k)
    assert rc == 0, lib.crossclr_last_error()
    ref = a.astype(np.float32) @ (b.astype(np.float32) if variant == 1 else b.astype(np.float32).T)
    err = np.abs(out - ref).max()
    assert err < 1e-2, (variant, n, k, err)


# ---------------------------------------------------------------------------------------------
# golden vectors produced by the unmodified reference
@pytest.mark.parametrize("path", FULL, ids=[os.path.basename(p)[:-4] for p in FULL])
def test_simt_matches_reference_goldens(path):
    g = np.load(path)
    loss, dv, dt = run_gpu(g["v"], g["t"], float(g["tau"]), float(g["w"]), path="simt")
    if "zero_row" in path:
        # the zero row's gradient (~1e13, through the eps clamp) dominates every norm: compare rows separately
        z = 3
        keep = np.arange(len(dv)) != z
        check(loss, dv[keep], dt, float(g["loss"]), g["dv"][keep], g["dt"], TOL_SIMT)
        assert np.linalg.norm(dv[z] - g["dv"][z]) <= 1e-3 * np.linalg.norm(g["dv"][z])
    else:
        check(loss, dv, dt, float(g["loss"]), g["dv"].astype(np.float64), g["dt"].astype(np.float64),
              TOL_SMALL_TAU if path in SMALL_TAU else TOL_SIMT)


@pytest.mark.parametrize("path", TC_OK, ids=[os.path.basename(p)[:-4] for p in TC_OK])
def test_tc_matches_reference_goldens(path):
    """Every reference golden inside the temperature range on the tensor-core path, ragged shapes included (B = 1, 8, 160;
    D = 4, 16, 192): the zero-padded layout must give the reference's numbers for the caller's rows."""
    g = np.load(path)
    loss, dv, dt = run_gpu(g["v"], g["t"], float(g["tau"]), float(g["w"]), path="tc")
    if "zero_row" in path:
        z = 3                     # its gradient (~1e13, through the eps clamp) dominates every norm: compare rows separately
        keep = np.arange(len(dv)) != z
        check(loss, dv[keep], dt, float(g["loss"]), g["dv"][keep].astype(np.float64), g["dt"].astype(np.float64), TOL)
        assert np.linalg.norm(dv[z] - g["dv"][z]) <= 1e-3 * np.linalg.norm(g["dv"][z])
    else:
        check(loss, dv, dt, float(g["loss"]), g["dv"].astype(np.float64), g["dt"].astype(np.float64), TOL)


@pytest.mark.parametrize("B,D,dtype,path,tau,s0", [
    (512, 256, torch.bfloat16, "tc", 0.03, 1.3), (512, 256, torch.bfloat16, "tc", 0.03, 0.4), (100, 40, torch.float32, "simt", 0.05, 1.3),
    (2048, 256, torch.float32, "auto", 0.03, 1.3), (2048, 256, torch.float32, "auto", 0.03, 0.5), (333, 77, torch.bfloat16, "auto", 0.03, 0.6),
    (128, 64, torch.float32, "auto", 0.006, 1.3)])
def test_learnable_temperature_gradient(B, D, dtype, path, tau, s0):
    """Opt-in extension: with learnable_temperature=True every logit is multiplied by the dormant
    `logit_scale` Parameter (trainer/loss.py:52).  Loss and feature gradients equal the oracle's at temperature tau / s, and
    d loss / d s equals the central difference of the float64 oracle loss in s."""
    from oracle import crossclr_oracle as O
    M = _mod()
    h = 1e-4
    v, t = _seeded(B, D, 5 + B + D, aligned=2.0, dtype=torch.bfloat16)
    rloss, rdv, rdt = O.loss_and_grads(v, t, tau / s0, 0.8)
    lp = O.loss_and_grads(v, t, tau / (s0 + h), 0.8)[0]
    lm = O.loss_and_grads(v, t, tau / (s0 - h), 0.8)[0]
    rds = (lp - lm) / (2 * h)
    crit = M.CrossCLR_onlyIntraModality(tau, 0.8, path=path, learnable_temperature=True).cuda()
    with torch.no_grad():
        crit.logit_scale.fill_(s0)
    vd = torch.from_numpy(v).to("cuda", dtype).requires_grad_()
    td = torch.from_numpy(t).to("cuda", dtype).requires_grad_()
    loss = crit(vd, td)
    (2.0 * loss).backward()
    tol = TOL if dtype == torch.float32 else 5e-3          # 16-bit gradients are rounded to the input dtype on the way out
    check(loss.item(), vd.grad.double().cpu().numpy() / 2, td.grad.double().cpu().numpy() / 2, rloss, rdv, rdt, tol)
    ds = crit.logit_scale.grad.item() / 2
    # d loss / d s = (1/s) [ E_softmax(logit) - positive logit ], averaged over rows: the difference of two sums of size
    # ~1/tau_eff that nearly cancel around the loss-optimal scale (s0 = 1.3 here: they are ~300 x their difference).  The bar is
    # 1e-3 of the value plus 5e-6 of that size -- the tensor-core paths' ~1e-5 accuracy on each sum
    assert abs(ds - rds) <= 1e-3 * abs(rds) + 5e-6 * s0 / tau, (ds, rds)
    # the default stays the reference's: the Parameter is dormant and gets no gradient
    ref = M.CrossCLR_onlyIntraModality(tau / s0, 0.8, path=path).cuda()
    l2 = ref(vd.detach().requires_grad_(), td.detach().requires_grad_())
    l2.backward()
    assert abs(l2.item() - loss.item()) <= 1e-6 * abs(loss.item()) and ref.logit_scale.grad is None   # s0 is an fp32 Parameter


def test_learnable_temperature_gradient_sums_over_ranks():
    """Every rank's crossclr_bwd_scale_grad holds its own rows' share: the shares of a 4-rank job sum to the single-rank value."""
    import ctypes
    from crossmodal_contrastive_learning_b200 import _native as N, loss as L
    ops = L._ops()
    Bg, D, tau, s0 = 2048, 256, 0.03, 0.9
    v, t = _seeded(Bg, D, 3, aligned=2.0)
    vd, td = torch.from_numpy(v).to("cuda", torch.bfloat16), torch.from_numpy(t).to("cuda", torch.bfloat16)
    totals = []
    for world in (1, 4):
        B = Bg // world
        probs = [N.Problem(2 * world, B, D, 2 * r * B, 2 * B, tau / s0, 0.8) for r in range(world)]
        code, fdt, pitch = ops.plan(probs[0], vd.dtype, False)
        S = ops.seg_rows(code, B)
        feat = torch.empty((2 * world, S, pitch), dtype=fdt, device="cuda")
        rnorm = torch.empty((world, 2 * B), dtype=torch.float32, device="cuda")
        stats = torch.empty((2 * world * S, 2), dtype=torch.float32, device="cuda")
        coef = torch.empty_like(stats)
        scal = torch.empty(4, dtype=torch.float32, device="cuda")
        loss = torch.empty((), dtype=torch.float64, device="cuda")
        go = torch.ones((), dtype=torch.float64, device="cuda")
        dv, dt = torch.empty((Bg, D), dtype=torch.float32, device="cuda"), torch.empty((Bg, D), dtype=torch.float32, device="cuda")
        ds = torch.zeros(world, dtype=torch.float64, device="cuda")
        for r in range(world):
            ops.pack2(vd[r * B:(r + 1) * B], td[r * B:(r + 1) * B], feat[2 * r:2 * r + 2], rnorm[r], code)
        for r in range(world):
            ops.fwd(probs[r], code, feat, stats)
        ops.finalize(probs[0], code, stats, coef, loss, scal)
        for r in range(world):
            ops.bwd(probs[r], code, feat, rnorm[r], coef, scal, go, 1.0, dv[r * B:(r + 1) * B], dt[r * B:(r + 1) * B], ds[r], s0)
        torch.cuda.synchronize()
        totals.append(ds.sum().item())
    assert abs(totals[0] - totals[1]) <= 1e-4 * abs(totals[0]), totals


@pytest.mark.parametrize("B,D,world,tau,noise", [(1024, 512, 1, 0.03, 2.0), (4096, 512, 1, 0.03, 2.0), (4096, 512, 1, 0.01, 4.0),
                                                 (2048, 256, 2, 0.03, 2.0), (1500, 500, 1, 0.02, 2.0), (2048, 1024, 1, 0.03, 2.0),
                                                 (4096, 640, 4, 0.0075, 8.0)])
def test_fp32_inputs_take_the_split_path(B, D, world, tau, noise):
    """The reference multiplies fp32 features in fp32 (trainer/loss.py:83-88).  `auto` keeps fp32 inputs on the tensor cores as
    fp16 hi + lo pairs -- S = hi.hi + lo.hi + hi.lo over K = 3 D -- so that nothing the caller passed is rounded to fp16:
    full-precision random inputs, several temperatures, single rank and row bands, against the float64 oracle at the
    tolerance the exact-operand path holds for 16-bit inputs (what is left is the fp16 probability tile)."""
    from oracle import crossclr_oracle as O
    from crossmodal_contrastive_learning_b200 import _native as N, loss as L
    Bl = B // world
    prob = N.Problem(2 * world, Bl, D, 0, 2 * Bl, tau, 0.8)
    assert L._ops().plan(prob, torch.float32, False)[0] == N.PATH_TC_SPLIT
    # `noise` keeps the problem un-converged at the small temperatures: at a loss of ~1e-12 the float64 reference itself is
    # rounding noise (its softmax - onehot cancels to the last bits), and so is the oracle that restates it
    g = torch.Generator().manual_seed(B + D + world)
    v = torch.randn(B, D, generator=g)
    t = v + noise * torch.randn(B, D, generator=g)
    rows = None if B <= 2048 else np.arange(0, B, B // 24) + 1
    rloss, rdv, rdt = O.loss_and_grads(v.numpy(), t.numpy(), tau, 0.8, rows=rows, row_block=2048)
    assert rloss > 1e-6
    loss, dv, dt = _run_ranks_on_one_gpu(v.cuda(), t.cuda(), world, tau, 0.8, "auto")
    if rows is not None:
        dv, dt = dv[rows], dt[rows]
    # below tau ~ 0.02 the softmax of this data is peaky: a few dominant probabilities carry their own fp16 rounding (2^-11)
    # and the tensor cores' fp32 accumulation error (~1e-6 of a logit's cosine, times log2e / tau) shows; still inside 1e-3
    tol = TOL_RAW if tau >= 0.02 else 5e-4
    check(loss, dv, dt, rloss, rdv, rdt, tol)
    # the module takes the same path and returns the same numbers
    if world == 1:
        lm, dvm, dtm = run_gpu(v.numpy(), t.numpy(), tau, 0.8, path="auto")
        if rows is not None:
            dvm, dtm = dvm[rows], dtm[rows]
        check(lm, dvm, dtm, rloss, rdv, rdt, tol)


@pytest.mark.parametrize("B,D,world", [(4000, 500, 1), (333, 77, 1), (1000, 200, 1), (130, 70, 1), (2100, 520, 1), (600, 72, 4),
                                       (3000, 500, 3), (4000, 1000, 2)])
def test_ragged_shapes_on_the_tensor_core_path(B, D, world):
    """trainer/loss.py:83-88 takes any [B, D].  `auto` keeps ragged shapes on the tensor-core kernels (segments zero-padded to
    128 rows, rows to 64 columns; the forward masks the padded columns out of the row sums, zero rows drop out of the
    backward); results equal the oracle's on the caller's rows at the exact-operand tolerance, single rank (symmetric
    schedules, fused finalize) and row-band ranks."""
    from oracle import crossclr_oracle as O
    from crossmodal_contrastive_learning_b200 import _native as N, loss as L
    Bl = B // world
    prob = N.Problem(2 * world, Bl, D, 0, 2 * Bl, 0.03, 0.8)
    assert L._ops().plan(prob, torch.bfloat16, False)[0] == N.PATH_TC
    v, t = _seeded(B, D, 77 + B + D, aligned=2.0)
    rows = None if B <= 2048 else np.arange(0, B, B // 24) + 1
    rloss, rdv, rdt = O.loss_and_grads(v, t, 0.03, 0.8, rows=rows, row_block=2048)
    loss, dv, dt = _run_ranks_on_one_gpu(torch.from_numpy(v).to("cuda", torch.bfloat16), torch.from_numpy(t).to("cuda", torch.bfloat16),
                                         world, 0.03, 0.8, "auto")
    if rows is not None:
        dv, dt = dv[rows], dt[rows]
    # the fp16 rounding of a probability tile (2^-11 per element) averages over D products: a little less at D < 128
    check(loss, dv, dt, rloss, rdv, rdt, TOL_RAW if D >= 128 else 3e-4)


@pytest.mark.parametrize("path", SMALL_TAU, ids=[os.path.basename(p)[:-4] for p in SMALL_TAU])
def test_small_temperature_goldens_route_to_row_shift(path):
    """tau = 0.005 (the low end of the reference's range): `auto` takes the exact path's online-maximum mode and matches the
    reference's float64 max-subtracted softmax; forcing the tensor-core path is refused, not silently flushed."""
    g = np.load(path)
    assert len(SMALL_TAU) >= 3
    loss, dv, dt = run_gpu(g["v"], g["t"], float(g["tau"]), float(g["w"]), path="auto")
    assert np.isfinite(loss) and np.isfinite(dv).all() and np.isfinite(dt).all()
    check(loss, dv, dt, float(g["loss"]), g["dv"].astype(np.float64), g["dt"].astype(np.float64), TOL_SMALL_TAU)
    if path in TC_SHAPED:
        with pytest.raises(RuntimeError, match="temperature"):
            run_gpu(g["v"], g["t"], float(g["tau"]), float(g["w"]), path="tc")


@pytest.mark.parametrize("path", ["simt", "tc"])
def test_c1_config_golden(path):
    """BASELINE.json configs[0]: B=256 D=512 fp32 seed 0."""
    g = np.load(os.path.join(GOLDEN, "c1_b256_d512_seed0.npz"))
    loss, dv, dt = run_gpu(g["v"], g["t"], 0.03, 0.8, path=path)
    tol = TOL_SIMT if path == "simt" else TOL
    assert abs(loss - 7.00737577904705) <= tol * 7.0074
    assert abs(np.linalg.norm(dv) - float(g["dv_norm"])) <= tol * float(g["dv_norm"])
    assert abs(np.linalg.norm(dt) - float(g["dt_norm"])) <= tol * float(g["dt_norm"])
    rows = g["rows"]
    check(loss, dv[rows], dt[rows], float(g["loss"]), g["dv_rows"].astype(np.float64),
          g["dt_rows"].astype(np.float64), tol)


def test_known_answers():
    from oracle import crossclr_oracle as O
    ref = np.load(os.path.join(GOLDEN, "kat_reference_values.npz"))
    cases = [(O.kat_identity(2, 1.0, 0.8), "identity_n2_tau1_w0.8"), (O.kat_identity(4, 0.5, 0.3), "identity_n4_tau0.5_w0.3"),
             (O.kat_collinear(3, 1.0, 0.8), "collinear_n3_tau1_w0.8"), (O.kat_collinear(5, 0.5, 0.25), "collinear_n5_tau0.5_w0.25"),
             (O.kat_antipodal(4, 0.5, 0.8), "antipodal_n4_tau0.5_w0.8")]
    for (v, t, tau, w, expect), key in cases:
        loss, dv, dt = run_gpu(v.astype(np.float32), t.astype(np.float32), tau, w, path="simt")
        assert abs(loss - expect) <= 1e-5 * abs(expect) + 1e-7, (key, loss, expect)
        assert abs(loss - float(ref[key])) <= 1e-5 * abs(expect) + 1e-7
    # KAT-4 gradient: v = t = I_2, tau = 1, w = .8 -> off-diagonal 0.9 / (e + 3), diagonal 0
    _, dv, dt = run_gpu(np.eye(2, dtype=np.float32), np.eye(2, dtype=np.float32), 1.0, 0.8, path="simt")
    g = 0.9 / (np.e + 3)
    assert np.allclose(dv, [[0, g], [g, 0]], atol=1e-6) and np.allclose(dt, [[0, g], [g, 0]], atol=1e-6)
    # converged regime: loss ~ 5e-14 must not collapse to 0 / nan (log1p form)
    v, t, tau, w, expect = O.kat_identity(8, 0.03, 0.8)
    loss, _, _ = run_gpu(v.astype(np.float32), t.astype(np.float32), tau, w, path="simt")
    assert abs(loss - expect) <= 1e-3 * expect


# ---------------------------------------------------------------------------------------------
# oracle on seeded inputs at sizes it finishes in seconds
def _seeded(B, D, seed, aligned=0.0, dtype=torch.bfloat16):
    g = torch.Generator().manual_seed(seed)
    v = torch.randn(B, D, generator=g)
    t = torch.randn(B, D, generator=g)
    if aligned:
        t = v + aligned * t
    return v.to(dtype).float().numpy(), t.to(dtype).float().numpy()     # bf16-representable fp32


@pytest.mark.parametrize("B,D,aligned,path,dtype", [
    (1024, 512, 0.0, "tc", torch.bfloat16), (1024, 512, 3.0, "tc", torch.bfloat16), (1024, 512, 0.5, "tc", torch.bfloat16),
    (512, 1024, 0.0, "tc", torch.bfloat16), (384, 320, 3.0, "tc", torch.float16), (640, 768, 0.0, "tc", torch.float32),
    (1000, 200, 0.0, "simt", torch.float32), (333, 77, 2.0, "auto", torch.bfloat16), (2048, 256, 0.0, "auto", torch.bfloat16),
])
def test_against_oracle(B, D, aligned, path, dtype):
    from oracle import crossclr_oracle as O
    v, t = _seeded(B, D, 100 + B + D, aligned, dtype if dtype != torch.float32 else torch.bfloat16)
    rloss, rdv, rdt = O.loss_and_grads(v, t, 0.03, 0.8)
    loss, dv, dt = run_gpu(v, t, 0.03, 0.8, path=path, dtype=dtype)
    # low-precision outputs: the gradient is rounded to the input dtype on the way out, as in the reference
    tol = TOL if dtype == torch.float32 else 5e-3 if dtype == torch.bfloat16 else 1.5e-3
    check(loss, dv, dt, rloss, rdv, rdt, tol)


def test_c2_config_against_oracle():
    """BASELINE.json configs[1]: B=4096 D=512 bf16 features, full-matrix check (fp32 grads out for a tight bar)."""
    from oracle import crossclr_oracle as O
    v, t = _seeded(4096, 512, 0)
    rloss, rdv, rdt = O.loss_and_grads(v, t, 0.03, 0.8)
    loss, dv, dt = run_gpu(v, t, 0.03, 0.8, path="tc", dtype=torch.float32)
    check(loss, dv, dt, rloss, rdv, rdt, TOL)
    lb, dvb, dtb = run_gpu(v, t, 0.03, 0.8, path="tc", dtype=torch.bfloat16)
    assert abs(lb - rloss) <= TOL * abs(rloss)
    assert np.linalg.norm(dvb - rdv) <= 5e-3 * np.linalg.norm(rdv)     # bf16 output rounding (2^-9 per element)


def test_c3_config_sampled_rows_and_properties():
    """BASELINE.json configs[2]: B=16384 D=1024 bf16.  Loss vs the row-blocked oracle, gradients on sampled
    rows, plus size-independent properties (symmetry, scale invariance, dv orthogonal to v)."""
    from oracle import crossclr_oracle as O
    B, D = 16384, 1024
    v, t = _seeded(B, D, 7)
    rows = np.arange(0, B, 1024) + 3
    rloss, rdv, rdt = O.loss_and_grads(v, t, 0.03, 0.8, rows=rows, row_block=2048)
    loss, dv, dt = run_gpu(v, t, 0.03, 0.8, path="tc", dtype=torch.float32)
    check(loss, dv[rows], dt[rows], rloss, rdv, rdt, TOL)
    # L(v, t) == L(t, v) and the gradients swap
    loss2, dv2, dt2 = run_gpu(t, v, 0.03, 0.8, path="tc", dtype=torch.float32)
    assert abs(loss - loss2) <= 1e-6 * abs(loss)
    assert np.linalg.norm(dv2 - dt) <= 1e-4 * np.linalg.norm(dt)
    # scale invariance: L(4v, t) == L(v, t), dv scales by 1/4; dv is orthogonal to v
    loss3, dv3, _ = run_gpu(4.0 * v, t, 0.03, 0.8, path="tc", dtype=torch.float32)
    assert abs(loss - loss3) <= 1e-6 * abs(loss)
    assert np.linalg.norm(4.0 * dv3 - dv) <= 1e-4 * np.linalg.norm(dv)
    cosang = np.abs((dv * v).sum(1)) / (np.linalg.norm(dv, axis=1) * np.linalg.norm(v, axis=1))
    assert cosang.max() < 1e-3


def _gpu_step(v, t, path="tc"):
    M = _mod()
    vd, td = v.detach().clone().requires_grad_(), t.detach().clone().requires_grad_()
    loss = M.CrossCLR_onlyIntraModality(0.03, 0.8, path=path).cuda()(vd, td)
    loss.backward()
    torch.cuda.synchronize()
    return loss.item(), vd.grad, td.grad


def _oracle_rows_given_stats(v, t, rows, logzv, logzt, tau, w):
    """dL/dv, dL/dt of the sample rows `rows` from the oracle's formulas (oracle.loss_and_grads), with
    the global log-normalisers given -- O(len(rows) * B * D), feasible at B = 131072."""
    from oracle import crossclr_oracle as O
    vh, nv = O.normalize_rows(v)
    th, nt = O.normalize_rows(t)
    B = vh.shape[0]
    c = 1.0 / (2.0 * B)
    r = np.arange(len(rows))
    a = vh[rows] @ th.T / tau
    ga = np.exp(a - logzv[rows, None]) + np.exp(a - logzt[None, :])
    ga[r, rows] -= 2.0
    cv = w * (vh[rows] @ vh.T) / tau
    gv = w * (np.exp(cv - logzv[rows, None]) + np.exp(cv - logzv[None, :]))
    gv[r, rows] = 0.0
    dvh = (c / tau) * (ga @ th + gv @ vh)
    at = th[rows] @ vh.T / tau
    gat = np.exp(at - logzt[rows, None]) + np.exp(at - logzv[None, :])
    gat[r, rows] -= 2.0
    ct = w * (th[rows] @ th.T) / tau
    gt = w * (np.exp(ct - logzt[rows, None]) + np.exp(ct - logzt[None, :]))
    gt[r, rows] = 0.0
    dth = (c / tau) * (gat @ vh + gt @ th)
    v64, t64 = np.asarray(v, np.float64), np.asarray(t, np.float64)
    return (O._normalize_backward(dvh, v64[rows], vh[rows], nv[rows]), O._normalize_backward(dth, t64[rows], th[rows], nt[rows]))


def _abi_step(v, t, tau=0.03, w=0.8, want_stats=False):
    """Single-rank forward + backward through the C ABI: inputs in their own dtype (bf16 here), fp32 gradients out, so that
    the 16-bit output rounding of the module path (gradient dtype = input dtype) does not mask kernel errors."""
    import ctypes
    from crossmodal_contrastive_learning_b200 import _native as N, loss as L
    ops = L._ops()
    B, D = v.shape
    prob = N.Problem(2, B, D, 0, 2 * B, tau, w)
    code, fdt, pitch = ops.plan(prob, v.dtype, False)
    S = ops.seg_rows(code, B)
    feat = torch.empty((2, S, pitch), dtype=fdt, device="cuda")
    rnorm = torch.empty(2 * B, dtype=torch.float32, device="cuda")
    stats = torch.empty((2 * S, 2), dtype=torch.float32, device="cuda")
    coef = torch.empty_like(stats)
    scal = torch.empty(4, dtype=torch.float32, device="cuda")
    loss = torch.empty((), dtype=torch.float64, device="cuda")
    go = torch.ones((), dtype=torch.float64, device="cuda")
    dv = torch.empty((B, D), dtype=torch.float32, device="cuda")
    dt = torch.empty((B, D), dtype=torch.float32, device="cuda")
    ops.forward_single(prob, code, v, t, feat, rnorm, stats, coef, scal, loss)
    ops.bwd(prob, code, feat, rnorm, coef, scal, go, 1.0, dv, dt)
    torch.cuda.synchronize()
    if want_stats:
        return loss.item(), dv, dt, stats, float(M_lib().crossclr_shift(ctypes.byref(prob))), code
    return loss.item(), dv, dt


@pytest.mark.parametrize("B,D", [(65536, 512), (131072, 1024)], ids=["c4_single_gpu", "c5_single_gpu"])
def test_c4_c5_full_size_against_oracle(B, D):
    """BASELINE.json configs[3] / configs[4] (bf16 features) at their full global batch on ONE GPU (the reference cannot run
    them at all: it would need 0.9 / 3.7 TB) against the CPU oracle on sampled rows:
      * the row statistics (log Z, positive logit, loss term) of 32 + 32 sampled stacked rows against oracle._row_stats
        -- O(B D) per row;
      * the gradients of 32 sampled sample rows against the oracle's gradient formulas, with the global log-normalisers
        taken from the GPU statistics that the first check spot-checks;
      * the loss against the mean of the GPU per-row values, which the first check pins row by row;
    plus the size-independent properties (swap symmetry, scale invariance, dv orthogonal to v)."""
    from oracle import crossclr_oracle as O
    from crossmodal_contrastive_learning_b200 import _native as N
    free, _ = torch.cuda.mem_get_info()
    if free < 30 * B * D:
        pytest.skip("not enough free device memory")
    tau, w = 0.03, 0.8
    g = torch.Generator().manual_seed(B + D)
    v_cpu = torch.randn(B, D, generator=g).to(torch.bfloat16)
    t_cpu = (v_cpu.float() + 2.0 * torch.randn(B, D, generator=g)).to(torch.bfloat16)
    v, t = v_cpu.cuda(), t_cpu.cuda()
    loss, dv, dt, stats, shift, code = _abi_step(v, t, tau, w, want_stats=True)
    assert code == N.PATH_TC
    st = stats.double().cpu().numpy()
    ln2 = np.log(2.0)
    # log Z_g (natural log, unshifted) of every stacked row from the GPU statistics
    X, xp = st[:, 0], st[:, 1]
    logz_gpu = (np.log2(X + np.exp2(xp)) + shift) * ln2
    # (1) sampled rows against the oracle's row statistics: the operands are the caller's own bf16 values, so the bars are
    # fp32-epilogue tight
    vn, tn = v_cpu.float().numpy(), t_cpu.float().numpy()
    vh, _ = O.normalize_rows(vn)
    th, _ = O.normalize_rows(tn)
    rows = (np.arange(32) * (B // 32) + 5).astype(np.int64)
    for mod, fh in ((0, vh), (1, th)):
        logz, pos, _, _ = O._row_stats(fh[rows], mod, rows, vh, th, tau, w)
        gidx = mod * B + rows
        assert np.abs(logz_gpu[gidx] - logz).max() <= 2e-4, ("logZ", mod, np.abs(logz_gpu[gidx] - logz).max())
        pos_gpu = (xp[gidx] + shift) * ln2
        assert np.abs(pos_gpu - pos).max() <= 2e-4, ("positive logit", mod, np.abs(pos_gpu - pos).max())
        term_gpu, term = logz_gpu[gidx] - pos_gpu, logz - pos
        assert (np.abs(term_gpu - term) <= TOL * np.abs(term) + 1e-6).all(), ("row loss term", mod, np.abs(term_gpu - term).max())
    # (2) the loss is the mean of the per-row terms
    loss_rows = np.log1p(X * np.exp2(-xp))
    assert abs(loss - loss_rows.mean()) <= 1e-6 * abs(loss_rows.mean())
    # (3) gradients of sampled rows from the oracle's formulas, given the (spot-checked) normalisers
    rdv, rdt = _oracle_rows_given_stats(vn, tn, rows, logz_gpu[:B], logz_gpu[B:], tau, w)
    for got, ref, name in ((dv[rows].double().cpu().numpy(), rdv, "dv"), (dt[rows].double().cpu().numpy(), rdt, "dt")):
        rel = np.linalg.norm(got - ref) / np.linalg.norm(ref)
        mx = np.abs(got - ref).max() / np.abs(ref).max()
        assert rel <= TOL and mx <= TOL, (name, rel, mx)
    # size-independent properties
    nrm = lambda x: float(x.double().norm())
    assert np.isfinite(loss) and 0.0 < loss < np.log(2.0 * B)
    loss2, dv2, dt2 = _abi_step(t, v, tau, w)
    assert abs(loss - loss2) <= 1e-6 * abs(loss)
    assert nrm(dv2 - dt) <= 1e-4 * nrm(dt) and nrm(dt2 - dv) <= 1e-4 * nrm(dv)
    del dv2, dt2
    loss3, dv3, _ = _abi_step((4.0 * v.float()).to(torch.bfloat16), t, tau, w)      # exact: a power of two
    assert abs(loss - loss3) <= 1e-6 * abs(loss)
    assert nrm(4.0 * dv3 - dv) <= 1e-4 * nrm(dv)
    del dv3
    vf = v.float()
    cosang = (dv * vf).sum(1).abs() / (dv.norm(dim=1) * vf.norm(dim=1))
    assert float(cosang.max()) < 1e-3


def M_lib():
    return _mod().load_native()


def _run_ranks_on_one_gpu(v, t, world, tau, w, path):
    # gradients always come back in fp32 (the C ABI's out_dtype is independent of the input dtype)
    """The criterion of `world` ranks through the C ABI on one GPU: each rank's pack / fwd / bwd launches against the shared
    stacked matrix, exactly what the ranks of a process group launch (the all-gathers become plain slices of one buffer)."""
    from crossmodal_contrastive_learning_b200 import _native as N, loss as L
    ops = L._ops()
    Bg, D = v.shape
    B = Bg // world
    probs = [N.Problem(2 * world, B, D, 2 * r * B, 2 * B, tau, w) for r in range(world)]
    code, fdt, pitch = ops.plan(probs[0], v.dtype, path == "simt", L._FORCED.get(path))
    S = ops.seg_rows(code, B)                 # the tensor-core layout pads every segment to a multiple of 128 rows
    feat = torch.empty((2 * world, S, pitch), dtype=fdt, device="cuda")
    rnorm = torch.empty((world, 2 * B), dtype=torch.float32, device="cuda")
    stats = torch.empty((2 * world * S, 2), dtype=torch.float32, device="cuda")
    coef = torch.empty_like(stats)
    scal = torch.empty(4, dtype=torch.float32, device="cuda")
    loss = torch.empty((), dtype=torch.float64, device="cuda")
    go = torch.ones((), dtype=torch.float64, device="cuda")
    dv, dt = torch.empty_like(v, dtype=torch.float32), torch.empty_like(t, dtype=torch.float32)
    for r in range(world):
        ops.pack2(v[r * B:(r + 1) * B], t[r * B:(r + 1) * B], feat[2 * r:2 * r + 2], rnorm[r], code)
    for r in range(world):
        ops.fwd(probs[r], code, feat, stats)
    ops.finalize(probs[0], code, stats, coef, loss, scal)
    for r in range(world):
        ops.bwd(probs[r], code, feat, rnorm[r], coef, scal, go, 1.0, dv[r * B:(r + 1) * B], dt[r * B:(r + 1) * B])
    torch.cuda.synchronize()
    return loss.item(), dv.double().cpu().numpy(), dt.double().cpu().numpy()


TOL_RAW = 2e-4      # tensor-core paths on 16-bit inputs: exact operands, fp32 epilogue; what is left is the fp16 P tile


@pytest.mark.parametrize("B,D,world,path,dtype", [
    (2048, 512, 4, "tc", torch.bfloat16), (4096, 512, 8, "tc", torch.bfloat16), (1024, 256, 2, "tc", torch.bfloat16),
    (2048, 1024, 4, "tc", torch.bfloat16), (1536, 384, 3, "tc", torch.float16), (4096, 512, 1, "tc", torch.bfloat16),
    (1024, 512, 1, "tc", torch.float16), (512, 1024, 1, "tc", torch.bfloat16), (2048, 512, 4, "tc", torch.float32),
    (600, 72, 4, "simt", torch.float32)])
def test_row_band_ranks_emulated_on_one_gpu(B, D, world, path, dtype):
    """Multi-rank gradients without a multi-GPU lease: the launches of every rank of a `world`-rank job (row-band
    schedule: owned rows x all columns, no symmetric half) run on one GPU against the oracle on the concatenated batch.
    16-bit inputs take the exact-operand tensor-core path and are held to TOL_RAW; fp32 inputs (rows rounded to fp16 after
    normalisation) to the north-star 1e-3."""
    from oracle import crossclr_oracle as O
    v, t = _seeded(B, D, 31 + B + world, aligned=2.0, dtype=dtype if dtype != torch.float32 else torch.bfloat16)
    rloss, rdv, rdt = O.loss_and_grads(v, t, 0.03, 0.8)
    loss, dv, dt = _run_ranks_on_one_gpu(torch.from_numpy(v).to("cuda", dtype), torch.from_numpy(t).to("cuda", dtype), world,
                                         0.03, 0.8, path)
    check(loss, dv, dt, rloss, rdv, rdt, TOL_SIMT if path == "simt" else TOL if dtype == torch.float32 else TOL_RAW)


@pytest.mark.parametrize("tau", [0.03, 0.01, 0.0075])
def test_small_temperature_on_exact_operands(tau):
    """With the caller's own 16-bit values as tensor-core operands the logit error no longer grows as 1 / tau."""
    from oracle import crossclr_oracle as O
    v, t = _seeded(512, 256, 5, aligned=2.0)
    rloss, rdv, rdt = O.loss_and_grads(v, t, tau, 0.8)
    loss, dv, dt = _run_ranks_on_one_gpu(torch.from_numpy(v).to("cuda", torch.bfloat16), torch.from_numpy(t).to("cuda", torch.bfloat16),
                                         1, tau, 0.8, "tc")
    check(loss, dv, dt, rloss, rdv, rdt, TOL_RAW)


def test_nan_features_give_nan_loss():
    """A NaN anywhere in the features makes the reference's loss NaN (trainer/loss.py:79-114 propagate it); so here."""
    M = _mod()
    for B, D, path in ((256, 128, "tc"), (2048, 256, "tc"), (40, 24, "simt")):
        v = torch.randn(B, D, device="cuda")
        t = torch.randn(B, D, device="cuda")
        v[3, 5] = float("nan")
        loss = M.CrossCLR_onlyIntraModality(0.03, 0.8, path=path)(v.requires_grad_(), t.requires_grad_())
        assert torch.isnan(loss).item(), (B, D, path, loss.item())


def test_host_fed_single_graph_step_matches_oracle():
    """HostFedCrossCLR: the whole step (H2D of the next inputs, forward, backward, D2H of the loss) as one graph launch;
    results equal the oracle's for every step's own inputs."""
    from oracle import crossclr_oracle as O
    M = _mod()
    B, D = 512, 256
    crit = M.CrossCLR_onlyIntraModality(0.03, 0.8).cuda()
    pipe = M.HostFedCrossCLR(crit, B, D, dtype=torch.float32)     # fp32 I/O: no 16-bit output rounding in the comparison
    data = [_seeded(B, D, 40 + k, aligned=2.0) for k in range(4)]
    pipe.host_video[0].copy_(torch.from_numpy(data[0][0]))
    pipe.host_text[0].copy_(torch.from_numpy(data[0][1]))
    pipe.prime()
    for k in range(4):
        if k + 1 < 4:
            pipe.host_video[(k + 1) & 1].copy_(torch.from_numpy(data[k + 1][0]))
            pipe.host_text[(k + 1) & 1].copy_(torch.from_numpy(data[k + 1][1]))
        s = pipe.step()
        pipe.drain()
        pipe.done[s].synchronize()
        rloss, rdv, rdt = O.loss_and_grads(data[k][0], data[k][1], 0.03, 0.8)
        check(float(pipe.loss_host[s]), pipe.grad_video[s].double().cpu().numpy(), pipe.grad_text[s].double().cpu().numpy(),
              rloss, rdv, rdt, TOL)
        torch.cuda.synchronize()          # the next step's upload overwrites the other set: keep the host buffers stable
    res = M.HostFedCrossCLR(crit, B, D, dtype=torch.float32, feed="device")
    res.video[0].detach().copy_(torch.from_numpy(data[2][0]))
    res.text[0].detach().copy_(torch.from_numpy(data[2][1]))
    s = res.step()
    torch.cuda.synchronize()
    rloss, rdv, rdt = O.loss_and_grads(data[2][0], data[2][1], 0.03, 0.8)
    check(res.loss[s].item(), res.grad_video[s].double().cpu().numpy(), res.grad_text[s].double().cpu().numpy(), rloss, rdv, rdt, TOL)


# ---------------------------------------------------------------------------------------------
# module surface / error behaviour
def test_module_surface_and_errors():
    M = _mod()
    from trainer.loss import CrossCLR_onlyIntraModality as FromTrainer
    assert FromTrainer is M.CrossCLR_onlyIntraModality
    crit = M.CrossCLR_onlyIntraModality(temperature=0.03, negative_weight=0.8).cuda()
    assert list(crit.state_dict().keys()) == ["logit_scale"]
    assert crit.temperature == 0.03 and crit.negative_w == 0.8 and crit.logger is None
    v = torch.randn(64, 32, device="cuda", requires_grad=True)
    t = torch.randn(64, 32, device="cuda", requires_grad=True)
    loss = crit(v, t)
    loss.backward()
    assert crit.logit_scale.grad is None
    with torch.no_grad():
        assert not crit(v, t).requires_grad
    for bad in [(torch.randn(64, 32, device="cuda"), torch.randn(32, 32, device="cuda")),
                (torch.randn(64, 32, device="cuda"), torch.randn(64, 16, device="cuda")),
                (torch.randn(2, 64, 32, device="cuda"), torch.randn(2, 64, 32, device="cuda")),
                (torch.randn(64, 32), torch.randn(64, 32))]:
        with pytest.raises(RuntimeError):
            crit(*bad)
    # attributes are read per call
    l1 = crit(v, t).item()
    crit.temperature = 0.1
    assert abs(crit(v, t).item() - l1) > 1e-3
    # non-contiguous inputs, upstream gradient scaling, float64 inputs
    crit.temperature = 0.03
    big = torch.randn(64, 64, device="cuda")
    vn = big[:, ::2].detach().requires_grad_()
    l2 = crit(vn, t)
    (l2 * 3.0).backward()
    vc = big[:, ::2].contiguous().requires_grad_()
    crit(vc, t).backward()
    assert torch.allclose(vn.grad, 3.0 * vc.grad, rtol=1e-5, atol=1e-9)
    v64 = v.detach().double().requires_grad_()
    l3 = crit(v64, t.detach().double())
    l3.backward()
    assert v64.grad.dtype == torch.float64 and abs(l3.item() - l1) <= 1e-5 * abs(l1)


def test_b1_and_w0_and_small_tau():
    for name in ("b1_d16", "w0_b64_d32", "tau01_b64_d64"):
        g = np.load(os.path.join(GOLDEN, name + ".npz"))
        loss, dv, dt = run_gpu(g["v"], g["t"], float(g["tau"]), float(g["w"]))
        check(loss, dv, dt, float(g["loss"]), g["dv"].astype(np.float64), g["dt"].astype(np.float64), TOL_SIMT)
    # tau = 0.0075 (beyond the fp32 exp range without the constant shift; the smallest temperature the constant
    # shift covers for arbitrary data, DESIGN.md "Numerics"): compare with the float64 oracle
    from oracle import crossclr_oracle as O
    v, t = _seeded(256, 128, 5, aligned=2.0)
    for tau in (0.0075,):
        rloss, rdv, rdt = O.loss_and_grads(v, t, tau, 0.8)
        for path in ("simt", "tc"):
            loss, dv, dt = run_gpu(v, t, tau, 0.8, path=path)
            # fp16 operand rounding of the unit rows perturbs each logit by ~1.7e-5/tau (DESIGN.md "Numerics"):
            # the tensor-core path's gradient error grows as 1/tau and is ~1.3e-3 here
            check(loss, dv, dt, rloss, rdv, rdt, TOL_SIMT if path == "simt" else 4e-3)


def test_cuda_graph_replay_matches_eager_and_oracle():
    """GraphedCrossCLR (forward + backward graphs) replays the same kernels: results equal the eager module's, new
    inputs are picked up on every call, and the gradients follow the upstream gradient."""
    from oracle import crossclr_oracle as O
    M = _mod()
    B, D = 512, 256
    crit = M.CrossCLR_onlyIntraModality(0.03, 0.8).cuda()
    step = M.GraphedCrossCLR(crit, B, D, dtype=torch.float32)
    for seed, scale in ((1, 1.0), (2, 0.5), (3, 3.0)):
        v, t = _seeded(B, D, seed, aligned=2.0)
        rloss, rdv, rdt = O.loss_and_grads(v, t, 0.03, 0.8)
        vd = torch.from_numpy(v).cuda().requires_grad_()
        td = torch.from_numpy(t).cuda().requires_grad_()
        loss = step(vd, td)
        (loss * scale).backward()
        torch.cuda.synchronize()
        check(loss.item(), vd.grad.double().cpu().numpy() / scale, td.grad.double().cpu().numpy() / scale, rloss, rdv, rdt, TOL)
        ve = torch.from_numpy(v).cuda().requires_grad_()
        te = torch.from_numpy(t).cuda().requires_grad_()
        le = crit(ve, te)
        (le * scale).backward()
        assert abs(le.item() - loss.item()) <= 1e-6 * abs(le.item())
        assert torch.allclose(ve.grad, vd.grad, rtol=1e-4, atol=1e-9)
        last_dv = vd.grad.clone()          # gradients are static buffers of the capture: the next replay overwrites them
    # the static inputs can be filled in place (no device-to-device copy on the call)
    step.video.detach().copy_(vd.detach())
    step.text.detach().copy_(td.detach())
    step.video.grad = None
    l2 = step(step.video, step.text)
    l2.backward()
    assert abs(l2.item() - loss.item()) <= 1e-6 * abs(loss.item())
    assert torch.allclose(step.video.grad * scale, last_dv, rtol=1e-4, atol=1e-9)
    with pytest.raises(RuntimeError):
        step(vd[:256], td[:256])


def test_custom_op_registration_and_torch_compile():
    """The single-rank criterion is two torch.library custom ops: opcheck passes (schema, fake kernel, autograd
    registration) and torch.compile traces through it with fullgraph=True, giving the eager numbers."""
    M = _mod()
    v, t = _seeded(256, 128, 5, aligned=2.0)
    vd = torch.from_numpy(v).cuda().requires_grad_()
    td = torch.from_numpy(t).cuda().requires_grad_()
    torch.library.opcheck(torch.ops.crossclr_b200.forward.default, (vd.detach(), td.detach(), 0.03, 0.8, "auto"),
                          test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))
    crit = M.CrossCLR_onlyIntraModality(0.03, 0.8).cuda()
    le = crit(vd, td)
    le.backward()
    ref_dv = vd.grad.clone()

    @torch.compile(fullgraph=True, backend="aot_eager")
    def f(a, b):
        return crit(a * 1.0, b * 1.0) * 2.0

    vc = torch.from_numpy(v).cuda().requires_grad_()
    tc = torch.from_numpy(t).cuda().requires_grad_()
    lc = f(vc, tc)
    lc.backward()
    # two runs differ by the order of fp32 atomics in the row sums (1e-8 relative) and what that does to fp16 roundings
    assert abs(lc.item() - 2.0 * le.item()) <= 1e-6 * abs(le.item())
    assert (vc.grad - 2.0 * ref_dv).norm() <= 1e-4 * ref_dv.norm()


def test_launch_counter_moves():
    M = _mod()
    n0 = M.launch_count()
    v, t = _seeded(128, 64, 1)
    run_gpu(v, t, 0.03, 0.8)
    assert M.launch_count() - n0 >= 4          # pack, forward (+ fused finalize), backward, grad_finish
