"""The 3xTF32 split of the default tier (csrc/tcpos.cuh / tcred.cuh / nconv_tc*.cuh) restated in numpy:
kind::tf32 truncates its fp32 operands to 10 mantissa bits, the kernels compute  D = A.B + A.B_lo + A_lo.B  with
x_lo = x - trunc(x)  (exact in fp32; x_lo is truncated again by the MMA).  Checks the design claims on the host: the
split product is fp32-grade (the 1e-4 parity tier has two orders of magnitude of margin), the single-pass product is
not, and the N-concatenated form  A.[B | B_lo] + A_lo.B  (tcred_kernel<true, true>) is the same sum of the same terms."""
import numpy as np

f32 = np.float32


def trunc_tf32(x):
    return (x.astype(f32).view(np.uint32) & np.uint32(0xFFFFE000)).view(f32)


def mma(a, b):
    """What the tensor core computes from fp32 operand bits: products of the truncated operands, wide accumulation."""
    return trunc_tf32(a).astype(np.float64) @ trunc_tf32(b).astype(np.float64)


def test_three_pass_split_is_fp32_grade_and_single_pass_is_not():
    rng = np.random.default_rng(1)
    A = rng.standard_normal((256, 224)).astype(f32)          # positions x (7 segments x 32 channels): the gcn mlp
    B = (rng.standard_normal((224, 32)) * 0.1).astype(f32)
    exact = A.astype(np.float64) @ B.astype(np.float64)
    A_lo, B_lo = A - trunc_tf32(A), B - trunc_tf32(B)
    assert np.array_equal(trunc_tf32(A).astype(np.float64) + A_lo.astype(np.float64), A.astype(np.float64))   # exact split
    x3 = mma(A, B) + mma(A, B_lo) + mma(A_lo, B)
    rel = lambda d: np.linalg.norm(d - exact) / np.linalg.norm(exact)
    assert rel(x3) < 2e-6            # remaining error: the dropped A_lo.B_lo term and the re-truncated remainders
    assert rel(mma(A, B)) > 1e-4     # single-pass TF32 cannot meet the 1e-4 tier


def test_n_concatenated_remainder_product_is_the_same_sum():
    rng = np.random.default_rng(2)
    A = rng.standard_normal((64, 96)).astype(f32)
    B = rng.standard_normal((96, 32)).astype(f32)
    A_lo, B_lo = A - trunc_tf32(A), B - trunc_tf32(B)
    cat = mma(A, np.concatenate([B, B_lo], axis=1))          # one instruction of twice the width: columns [0,N) | [N,2N)
    folded = cat[:, :32] + mma(A_lo, B) + cat[:, 32:]        # A_lo.B accumulates onto [0,N); the drain adds [N,2N)
    three = mma(A, B) + mma(A, B_lo) + mma(A_lo, B)
    assert np.allclose(folded, three, rtol=1e-13, atol=1e-13)
