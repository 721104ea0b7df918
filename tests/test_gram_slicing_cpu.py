"""CPU checks of the specification behind the Gram leaf of the R-only TSQR (cuda-qr_b200/csrc/gram_umma.cu): the slicing of
fp32 values into three bf16 slices (tools/gram_slicing_spec.py restates the kernel's arithmetic in numpy) must be bf16-exact,
integer-valued in the group's quantum and small enough that 128-row sums stay below 2^24 -- the premise of "exact in the fp32
accumulator whatever its rounding mode" -- and the Gram matrix assembled from the slices must match fp64 to the stated level.
The kernel itself is checked on the GPU (tests/test_gpu_qr.py::test_tsqr_gram_leaf*)."""
import importlib.util
import os

import numpy as np
import pytest

_spec = importlib.util.spec_from_file_location(
    "gram_slicing_spec", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "gram_slicing_spec.py"))
spec = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(spec)
F = np.float32


@pytest.mark.parametrize("kind", ["uniform", "normal", "graded", "integers", "one_large", "tiny_and_huge_scale"])
def test_slices_are_bf16_exact_integer_and_bounded(kind):
    rng = np.random.default_rng(7)
    for t in range(60):
        if kind == "uniform":
            v = rng.random(128)
        elif kind == "normal":
            v = rng.standard_normal(128)
        elif kind == "graded":
            v = rng.standard_normal(128) * 10.0 ** rng.uniform(-7, 0, 128)
        elif kind == "integers":
            v = rng.integers(-2000, 2001, 128).astype(np.float64)
        elif kind == "one_large":
            v = rng.standard_normal(128) * 1e-5
            v[rng.integers(0, 128)] = 3.999999
        else:
            v = rng.standard_normal(128) * 2.0 ** rng.integers(-38, 38)
        spec.check_group(v.astype(F))
    # worst case for the accumulation bound: every entry at the top of its binade
    spec.check_group(np.full(128, np.nextafter(F(2.0), F(0.0)), dtype=F))
    spec.check_group(-np.full(128, np.nextafter(F(2.0), F(0.0)), dtype=F))
    spec.check_group(np.zeros(128, dtype=F))


def test_top_binade_values_are_represented_exactly():
    rng = np.random.default_rng(3)
    v = (1.0 + rng.random(128)).astype(F)            # all in [1, 2): the group's top binade (2^E = 2)
    s1, s2, s3, E = spec.slice_group(v)
    assert E == 1
    assert np.array_equal(s1.astype(np.float64) + s2 + s3, v.astype(np.float64))


@pytest.mark.parametrize("kind,tol", [("uniform", 2e-9), ("normal", 2e-8), ("eight_bit", 0.0)])
def test_gram_matrix_from_slices(kind, tol):
    rng = np.random.default_rng(11)
    m, n = 1024 + 77, 24
    if kind == "uniform":
        A = rng.random((m, n))
    elif kind == "normal":
        A = rng.standard_normal((m, n))
    else:
        A = rng.integers(-255, 256, (m, n)) / 256.0
    A = A.astype(F)
    G = spec.gram_of_slices(A)
    Gx = A.astype(np.float64).T @ A.astype(np.float64)
    if tol == 0.0:
        assert np.array_equal(G, Gx)                  # data that fits the first slice: bit-exact
    else:
        assert np.linalg.norm(G - Gx) / np.linalg.norm(Gx) <= tol
    # ... and R = chol(G) is the R factor to fp32-rounding level (positive diagonal)
    if kind != "eight_bit":
        R = np.linalg.cholesky(G).T
        R64 = np.linalg.qr(A.astype(np.float64), mode="r")
        R64 = R64 * np.sign(np.diag(R64))[:, None]
        assert np.linalg.norm(R - R64) / np.linalg.norm(R64) <= 1e-7
