"""The 3xFP16 splitting the float32 tensor-core path uses (csrc/matvec_tf32.cu, `prepare_f16_kernel`), restated in
NumPy: per row a power-of-two scale puts the row maximum in [2^14, 2^15), H = fp16(x 2^s), R = fp16(x 2^s - H), and
x.z 2^(sx+sz) = H.H + R.H + H.R.  Checks on the CPU that the scheme delivers float32-level dot products whatever the
magnitudes of the rows (the GPU tests compare the kernel itself with the oracle)."""
import numpy as np


def split_rows(X):
    X = np.asarray(X, np.float32)
    mx = np.abs(X).max(axis=1)
    e = np.frexp(np.where(mx > 0, mx, 1.0))[1]          # mx = f 2^e, f in [0.5, 1)
    s = np.where(mx > 0, np.clip(15 - e, -100, 100), 0)
    xs = X * np.exp2(s).astype(np.float32)[:, None]     # exact (power of two)
    H = xs.astype(np.float16)
    R = (xs - H.astype(np.float32)).astype(np.float16)  # the difference is exact in float32
    return H, R, np.exp2(-s.astype(np.float64))


def dot3(Hp, Rp, rp, Hq, Rq, rq):
    f = np.float64  # products of FP16 numbers are exact in the tensor core; accumulate wide here
    acc = Hp.astype(f) @ Hq.astype(f).T + Rp.astype(f) @ Hq.astype(f).T + Hp.astype(f) @ Rq.astype(f).T
    return acc * rp[:, None] * rq[None, :]


def test_split_is_nearly_exact_and_in_range():
    rng = np.random.default_rng(0)
    X = rng.standard_normal((64, 90)).astype(np.float32)
    X[::5] *= np.float32(1e-6)
    X[1::7] *= np.float32(3e4)
    X[3] = 0
    H, R, r = split_rows(X)
    assert np.isfinite(H.astype(np.float32)).all() and np.abs(H.astype(np.float32)).max() < 2.0 ** 15
    back = (H.astype(np.float64) + R.astype(np.float64)) * r[:, None]
    row = np.abs(X).max(axis=1, keepdims=True).astype(np.float64)
    assert (np.abs(back - X) <= 2.0 ** -21 * row + 1e-300).all()  # 22 bits relative to the row maximum


def test_dot_products_reach_float32_accuracy_for_any_row_magnitudes():
    rng = np.random.default_rng(1)
    P = rng.standard_normal((40, 90)).astype(np.float32)
    Q = rng.standard_normal((50, 90)).astype(np.float32)
    P[::3] *= np.float32(1e-4)
    Q[::4] *= np.float32(2e3)
    P[7] = 0
    exact = P.astype(np.float64) @ Q.astype(np.float64).T
    got = dot3(*split_rows(P), *split_rows(Q))
    scale = np.linalg.norm(P.astype(np.float64), axis=1)[:, None] * np.linalg.norm(Q.astype(np.float64), axis=1)[None, :]
    err = np.abs(got - exact) / np.maximum(scale, 1e-300)
    assert err.max() < 2.0 ** -20, err.max()
    # a float32 dot product of the same data (what the FFMA kernels compute) is no better
    f32 = (P[:, None, :] * Q[None, :, :]).sum(-1, dtype=np.float32).astype(np.float64)
    err32 = np.abs(f32 - exact) / np.maximum(scale, 1e-300)
    assert err.max() < 8 * max(err32.max(), 2.0 ** -24)
