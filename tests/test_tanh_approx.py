"""Error bound of the rational tanh used by the tensor-core GEMM epilogues (csrc/tc_common.cuh tanh_rational), restated in
numpy float32 with fused multiply-adds emulated in float64.  The CUDA function itself is exercised by the GPU tests of the
tensor-core forward (test_tc_linear_forward_vs_fp64, learn_Smid / learn_S fixtures)."""
import numpy as np

ALPHA = [4.89352455891786e-03, 6.37261928875436e-04, 1.48572235717979e-05, 5.12229709037114e-08, -8.60467152213735e-11,
         2.00018790482477e-13, -2.76076847742355e-16]
BETA = [4.89352518554385e-03, 2.26843463243900e-03, 1.18534705686654e-04, 1.19825839466702e-06]
f32 = np.float32


def fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + np.float64(c)).astype(f32)


def tanh_rational(x):
    x = np.clip(x, f32(-7.90531110763549805), f32(7.90531110763549805)).astype(f32)
    x2 = (x * x).astype(f32)
    p = np.full_like(x, f32(ALPHA[6]))
    for c in ALPHA[5::-1]:
        p = fma(x2, p, f32(c))
    q = np.full_like(x, f32(BETA[3]))
    for c in BETA[2::-1]:
        q = fma(x2, q, f32(c))
    return ((x * p).astype(f32) / q).astype(f32)


def test_rational_tanh_error_bound():
    rng = np.random.default_rng(0)
    xs = np.concatenate([np.linspace(-10, 10, 1_000_001), rng.standard_normal(500_000) * 0.5, np.linspace(-1e-3, 1e-3, 20_001),
                         [0.0, -0.0, 50.0, -50.0]]).astype(f32)
    ref = np.tanh(xs.astype(np.float64))
    got = tanh_rational(xs).astype(np.float64)
    assert np.abs(got - ref).max() <= 3.0e-7
    assert np.all(np.abs(got) <= 1.0)
    assert np.array_equal(np.sign(got), np.sign(ref))
    # odd function, exact zero at zero
    assert np.array_equal(tanh_rational(-xs), -tanh_rational(xs))
