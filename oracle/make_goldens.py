"""Generate tests/golden/*.npz from the UNMODIFIED reference (test infrastructure only).

Run in the build container, where /root/reference exists:

    python oracle/make_goldens.py

It imports `/root/reference/trainer/loss.py` as-is.  The only shim is `torch.Tensor.cuda -> identity`
(loss.py:66,103,104 hard-code `.cuda()`, and this container has no GPU); every arithmetic op is the
reference's own.  The GPU box never sees /root/reference -- it only sees the committed .npz files.
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("CROSSCLR_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def reference_module(tau, w):
    sys.path.insert(0, REF)
    torch.Tensor.cuda = lambda self, *a, **k: self   # placement shim only
    from trainer.loss import CrossCLR_onlyIntraModality
    return CrossCLR_onlyIntraModality(temperature=tau, negative_weight=w)


def run_ref(v, t, tau, w, dtype=torch.float32):
    v = torch.tensor(v, dtype=dtype, requires_grad=True)
    t = torch.tensor(t, dtype=dtype, requires_grad=True)
    loss = reference_module(tau, w)(v, t)
    loss.backward()
    return loss.item(), v.grad.double().numpy(), t.grad.double().numpy()


def randn_case(seed, B, D, aligned=0.0):
    torch.manual_seed(seed)
    v = torch.randn(B, D)
    t = torch.randn(B, D)
    if aligned:
        t = v + aligned * t
    return v.numpy().copy(), t.numpy().copy()


def bf16_representable(x):
    return torch.tensor(x).to(torch.bfloat16).to(torch.float32).numpy()


def main():
    os.makedirs(OUT, exist_ok=True)
    cases = {}
    # --- RNG goldens + extra regimes.  Inputs are stored, not just seeds,
    #     so the fixtures do not depend on torch's CPU RNG stream.
    spec = [
        # name,            seed, B,   D,   aligned, tau,  w,   bf16-representable inputs
        ("rand_b8_d4",       0,   8,   4,  0.0, 0.03, 0.8, False),
        ("rand_b32_d64",     2,  32,  64,  0.0, 0.03, 0.8, True),
        ("rand_b128_d64",    3, 128,  64,  0.0, 0.03, 0.8, True),
        ("rand_b160_d192",   4, 160, 192,  0.0, 0.07, 0.5, True),     # ragged: B not /128, D not /256
        ("align_b128_d128",  5, 128, 128,  3.0, 0.03, 0.8, True),     # non-uniform softmax
        ("conv_b128_d128",   6, 128, 128,  0.3, 0.03, 0.8, True),     # near-converged
        ("w0_b64_d32",       7,  64,  32,  0.0, 0.05, 0.0, True),     # negative_weight = 0
        ("tau01_b64_d64",    8,  64,  64,  0.0, 0.01, 0.8, True),     # small temperature
        ("rand_b256_d256",   9, 256, 256,  1.0, 0.03, 0.8, True),
        ("b1_d16",          10,   1,  16,  0.0, 0.03, 0.8, True),     # B = 1
        # temperatures below the constant-shift range (reference float64 softmax subtracts the row max, loss.py:59-60)
        ("tau0075_b128_d128", 12, 128, 128, 0.0, 0.0075, 0.8, True),  # weakly aligned rows: every cosine small
        ("tau005_b128_d128",  13, 128, 128, 0.0, 0.005, 0.8, True),
        ("tau005_align_b128_d64", 14, 128, 64, 2.0, 0.005, 0.8, True),   # mixed: some rows converged, some not
        ("tau005_w2_b96_d48", 15,  96,  48, 2.0, 0.005, 2.0, True),   # |w| > 1, ragged
    ]
    for name, seed, B, D, al, tau, w, rep in spec:
        v, t = randn_case(seed, B, D, al)
        if rep:
            v, t = bf16_representable(v), bf16_representable(t)
        loss, dv, dt = run_ref(v, t, tau, w)
        cases[name] = dict(v=v.astype(np.float32), t=t.astype(np.float32), tau=tau, w=w, loss=loss,
                           dv=dv.astype(np.float32), dt=dt.astype(np.float32))
        print(f"{name:18s} loss={loss!r}")
    for name, c in cases.items():
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **c)

    # --- config c1 (BASELINE.json configs[0]): B=256 D=512 fp32 seed 0.  Inputs regenerate from the
    #     seed; store loss, gradient norms and a strided sample of gradient rows (keeps the file small).
    v, t = randn_case(0, 256, 512)
    loss, dv, dt = run_ref(v, t, 0.03, 0.8)
    rows = np.arange(0, 256, 16)
    np.savez_compressed(os.path.join(OUT, "c1_b256_d512_seed0.npz"), v=v, t=t, tau=0.03, w=0.8, loss=loss,
                        dv_norm=np.linalg.norm(dv), dt_norm=np.linalg.norm(dt), rows=rows,
                        dv_rows=dv[rows].astype(np.float32), dt_rows=dt[rows].astype(np.float32),
                        dv00=dv[0, 0])
    print(f"c1 loss={loss!r} |dv|={np.linalg.norm(dv)!r} |dt|={np.linalg.norm(dt)!r} dv00={dv[0,0]!r}")

    # --- closed-form known-answer values as evaluated BY THE REFERENCE (App. C.1)
    kat = {}
    eye = lambda n: np.eye(n, dtype=np.float32)
    kat["identity_n2_tau1_w0.8"] = run_ref(eye(2), eye(2), 1.0, 0.8)[0]
    kat["identity_n4_tau0.5_w0.3"] = run_ref(eye(4), eye(4), 0.5, 0.3)[0]
    col = lambda n: np.tile(np.array([[3.0, 4.0]], dtype=np.float32), (n, 1))
    kat["collinear_n3_tau1_w0.8"] = run_ref(col(3), 2 * col(3), 1.0, 0.8)[0]
    kat["collinear_n5_tau0.5_w0.25"] = run_ref(col(5), 2 * col(5), 0.5, 0.25)[0]
    kat["antipodal_n4_tau0.5_w0.8"] = run_ref(eye(4), -eye(4), 0.5, 0.8)[0]
    l, dv, dt = run_ref(eye(2), eye(2), 1.0, 0.8)
    kat["grad_identity_n2_dv01"] = dv[0, 1]
    kat["grad_identity_n2_dv00"] = dv[0, 0]
    # zero-norm row (eps clamp path, loss.py:79): finite loss, huge finite grad
    vz, tz = randn_case(11, 8, 8)
    vz[3] = 0.0
    lz, dvz, dtz = run_ref(vz, tz, 0.03, 0.8)
    np.savez_compressed(os.path.join(OUT, "zero_row_b8_d8.npz"), v=vz, t=tz, tau=0.03, w=0.8, loss=lz,
                        dv=dvz, dt=dtz)
    np.savez(os.path.join(OUT, "kat_reference_values.npz"), **{k: np.float64(x) for k, x in kat.items()})
    for k, x in kat.items():
        print(f"{k:30s} {x!r}")


if __name__ == "__main__":
    main()
