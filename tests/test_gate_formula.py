"""The arithmetic of the gate epilogues (csrc/rowepi.cuh: gate_pair) restated in float32 numpy: two gates
tanh(a) * sigmoid(b) share ONE reciprocal, exponents clamped at 30.  Checks the design claims the kernel comment makes --
no overflow / NaN anywhere in the fp32 range, ~1e-7 absolute error -- on the host; the kernel itself is covered by the
-m gpu parity tests (model.py:208-212)."""
import numpy as np

f32 = np.float32
K_TANH, K_SIGM = f32(2.8853900817779268), f32(-1.4426950408889634)    # 2 log2(e), -log2(e)


def gate_pair(a, b):
    """a, b: [n, 2] float32 pre-activations (bias already added) -> forward product, tanh, sigmoid per gate."""
    with np.errstate(over="ignore"):     # v * k may overflow to inf before the clamp, as in the kernel's FMA
        ea = np.exp2(np.minimum(a * K_TANH, f32(30))).astype(f32)
        eb = np.exp2(np.minimum(b * K_SIGM, f32(30))).astype(f32)
    num, q1, p1 = ea - f32(1), ea + f32(1), eb + f32(1)
    den = q1 * p1
    r = (f32(1) / (den[:, 0] * den[:, 1])).astype(f32)
    rd = np.stack([r * den[:, 1], r * den[:, 0]], axis=1)           # 1 / den of each gate
    return num * rd, num * (rd * p1), rd * q1


def test_paired_reciprocal_gate_matches_exact_functions():
    rng = np.random.default_rng(0)
    a = (rng.standard_normal((200000, 2)) * 4).astype(f32)
    b = (rng.standard_normal((200000, 2)) * 4).astype(f32)
    out, f, s = gate_pair(a, b)
    t = np.tanh(a.astype(np.float64))
    g = 1.0 / (1.0 + np.exp(-b.astype(np.float64)))
    assert np.abs(out - t * g).max() < 5e-7
    assert np.abs(f - t).max() < 5e-7 and np.abs(s - g).max() < 5e-7


def test_paired_reciprocal_gate_saturates_without_overflow():
    big = np.finfo(np.float32).max
    vals = np.array([0.0, 1e-30, -1e-30, 10.0, -10.0, 20.0, -20.0, 1e3, -1e3, 1e30, -1e30, big, -big], dtype=f32)
    a = np.array([[x, y] for x in vals for y in vals], dtype=f32)
    for shift in range(len(vals)):
        b = np.roll(a, shift, axis=0)[:, ::-1].copy()
        out, f, s = gate_pair(a, b)
        assert np.isfinite(out).all() and np.isfinite(f).all() and np.isfinite(s).all()
        t = np.tanh(a.astype(np.float64))
        g = 1.0 / (1.0 + np.exp(-np.clip(b.astype(np.float64), -700, 700)))
        assert np.abs(out - t * g).max() < 5e-7
        assert np.abs(f - t).max() < 5e-7 and np.abs(s - g).max() < 5e-7
