"""The CPU oracle (oracle/crossclr_oracle.py) against the reference-generated golden vectors and the
closed-form known-answer cases.  CPU only."""
import glob
import os

import numpy as np
import pytest

from oracle import crossclr_oracle as O

from conftest import GOLDEN

FULL = sorted(p for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
              if not os.path.basename(p).startswith(("kat_", "c1_")))


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("path", FULL, ids=[os.path.basename(p)[:-4] for p in FULL])
def test_oracle_matches_reference_goldens(path):
    g = np.load(path)
    loss, dv, dt = O.loss_and_grads(g["v"], g["t"], float(g["tau"]), float(g["w"]), row_block=48)
    ref = float(g["loss"])
    # the reference computes -log(softmax) in float64 from fp32 matmuls; allow its own noise floor
    assert abs(loss - ref) <= 2e-6 * abs(ref) + 1e-9, (loss, ref)
    if "zero_row" in path:
        # eps-clamp path: the zero row's gradient is ~1e13; compare relatively per row
        assert rel(dv, g["dv"]) < 1e-4 and rel(dt, g["dt"]) < 1e-4
    else:
        assert rel(dv, g["dv"]) < 2e-5, rel(dv, g["dv"])
        assert rel(dt, g["dt"]) < 2e-5, rel(dt, g["dt"])


def test_oracle_c1_config():
    g = np.load(os.path.join(GOLDEN, "c1_b256_d512_seed0.npz"))
    loss, dv, dt = O.loss_and_grads(g["v"], g["t"], 0.03, 0.8)
    assert abs(loss - 7.00737577904705) < 1e-6
    assert abs(loss - float(g["loss"])) < 1e-6
    assert abs(np.linalg.norm(dv) - float(g["dv_norm"])) < 1e-6 * float(g["dv_norm"])
    assert abs(np.linalg.norm(dt) - float(g["dt_norm"])) < 1e-6 * float(g["dt_norm"])
    rows = g["rows"]
    assert rel(dv[rows], g["dv_rows"]) < 2e-5
    assert rel(dt[rows], g["dt_rows"]) < 2e-5
    # row-subset mode returns the same rows
    _, dvs, dts = O.loss_and_grads(g["v"], g["t"], 0.03, 0.8, rows=rows)
    assert np.allclose(dvs, dv[rows], rtol=0, atol=1e-15)
    assert np.allclose(dts, dt[rows], rtol=0, atol=1e-15)


def test_known_answers_closed_form_and_reference():
    ref = np.load(os.path.join(GOLDEN, "kat_reference_values.npz"))
    cases = [
        (O.kat_identity(2, 1.0, 0.8), "identity_n2_tau1_w0.8"),
        (O.kat_identity(4, 0.5, 0.3), "identity_n4_tau0.5_w0.3"),
        (O.kat_collinear(3, 1.0, 0.8), "collinear_n3_tau1_w0.8"),
        (O.kat_collinear(5, 0.5, 0.25), "collinear_n5_tau0.5_w0.25"),
        (O.kat_antipodal(4, 0.5, 0.8), "antipodal_n4_tau0.5_w0.8"),
    ]
    for (v, t, tau, w, expect), key in cases:
        got = O.loss_only(v, t, tau, w)
        assert abs(got - expect) < 1e-12, (key, got, expect)
        assert abs(got - float(ref[key])) < 1e-6, (key, got, float(ref[key]))
    # gradient KAT-4: v = t = I_2, tau = 1, w = .8 -> off-diagonal 0.9/(e+3), diagonal 0
    v, t, tau, w, _ = O.kat_identity(2, 1.0, 0.8)
    _, dv, dt = O.loss_and_grads(v, t, tau, w)
    e = 0.9 / (np.e + 3.0)
    assert np.allclose(dv, [[0, e], [e, 0]], atol=1e-12)
    assert np.allclose(dt, [[0, e], [e, 0]], atol=1e-12)
    assert abs(float(ref["grad_identity_n2_dv01"]) - e) < 1e-7


def test_converged_identity_no_cancellation():
    # v = t = I_8, tau = .03: analytic 5.007e-14 (the reference itself returns 4.996e-14)
    v, t, tau, w, expect = O.kat_identity(8, 0.03, 0.8)
    got = O.loss_only(v, t, tau, w)
    assert abs(got - expect) < 2e-15


def test_symmetry_and_scale_invariance():
    rng = np.random.default_rng(0)
    v = rng.standard_normal((24, 16))
    t = rng.standard_normal((24, 16))
    l1, dv, dt = O.loss_and_grads(v, t)
    l2, dt2, dv2 = O.loss_and_grads(t, v)
    assert abs(l1 - l2) < 1e-12 and np.allclose(dv, dv2) and np.allclose(dt, dt2)
    l3, dv3, _ = O.loss_and_grads(3.0 * v, t)
    assert abs(l1 - l3) < 1e-12 and np.allclose(dv3 * 3.0, dv)
    assert np.abs((dv * v).sum(1)).max() < 1e-12          # dv is orthogonal to v


def test_w0_adds_B_to_every_denominator():
    rng = np.random.default_rng(1)
    B = 12
    v = rng.standard_normal((B, 8)); t = rng.standard_normal((B, 8))
    vh, _ = O.normalize_rows(v); th, _ = O.normalize_rows(t)
    a = vh @ th.T / 0.05
    lv = np.log(np.exp(a).sum(1) + B) - np.diag(a)
    lt = np.log(np.exp(a).sum(0) + B) - np.diag(a)
    assert abs(O.loss_only(v, t, 0.05, 0.0) - (lv.mean() + lt.mean()) / 2) < 1e-12


def test_sharded_emulation_matches_global():
    rng = np.random.default_rng(2)
    v = rng.standard_normal((64, 32)); t = v + rng.standard_normal((64, 32))
    l, dv, dt = O.loss_and_grads(v, t)
    ls, dvs, dts = O.sharded_loss_and_grads(v, t, 8)
    assert abs(l - ls) < 1e-13
    assert np.allclose(np.concatenate(dvs), dv, atol=1e-14)
    assert np.allclose(np.concatenate(dts), dt, atol=1e-14)


def test_gradient_matches_finite_differences():
    rng = np.random.default_rng(3)
    v = rng.standard_normal((6, 5)); t = rng.standard_normal((6, 5))
    _, dv, dt = O.loss_and_grads(v, t, 0.2, 0.6)
    h = 1e-6
    for (arr, g) in ((v, dv), (t, dt)):
        for idx in [(0, 0), (2, 3), (5, 4)]:
            old = arr[idx]
            arr[idx] = old + h; lp = O.loss_only(v, t, 0.2, 0.6)
            arr[idx] = old - h; lm = O.loss_only(v, t, 0.2, 0.6)
            arr[idx] = old
            assert abs((lp - lm) / (2 * h) - g[idx]) < 1e-7
