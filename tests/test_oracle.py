"""CPU tests that PIN the oracle (oracle/mmqr_oracle.c) before anything trusts it:
against the reference's one known-answer run, against golden vectors produced by the
unmodified reference (tests/golden/make_golden.py), and -- when oracle/_ref is present --
bit for bit against the compiled reference itself on fresh seeded inputs."""
import ctypes
import json
import os

import numpy as np
import pytest

import oracle
from oracle import metrics

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = [(4, 2, 6, 4), (4, 2, 64, 32), (64, 4, 64, 64), (64, 4, 124, 64), (64, 4, 244, 124), (64, 8, 512, 512)]


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "ref_cases.npz"))


def test_glibc_rand_restatement_matches_libc():
    libc = ctypes.CDLL(None)
    libc.srand(12)
    want = [libc.rand() for _ in range(2000)]
    got = oracle.glibc_rand(12, 2000)
    assert want[:3] == [1687063760, 945274514, 247215794]
    assert np.array_equal(np.array(want), got)


def test_rand_matrix_matches_c_recipe(port):
    assert np.array_equal(port.rand_matrix(37, 11, 12), oracle.rand_matrix(37, 11, 12))
    assert np.array_equal(port.rand_matrix(8, 8, 5), oracle.rand_matrix(8, 8, 5))


def test_known_answer_6x4_demo(port):
    """qr.c:461-523: the only result the reference pins (values as printed)."""
    d = json.load(open(os.path.join(GOLD, "demo_6x4.json")))
    A = oracle.rand_matrix(6, 4, 12)
    assert np.allclose(A, np.array(d["A_rowmajor_rows"], dtype=np.float32), atol=0)
    rv, tau = port.mmqr(A, 4, 2)
    printed_rv_row0 = [-1.411796, -0.752699, -0.820252, -1.395984]
    printed_tau = [1.105875, 1.437911, 1.556454, 1.382508, 1.629178, 1.088601, 2.0, 2.0]
    assert np.allclose(rv[0], printed_rv_row0, atol=1e-6)
    assert np.allclose(tau, printed_tau, atol=1e-6)
    assert np.array_equal(rv, np.array(d["RV_rows"], dtype=np.float32))
    assert np.array_equal(tau, np.array(d["tau"], dtype=np.float32))
    Q, R = port.explicitQR(rv, tau, 4, 2)
    assert np.array_equal(Q, np.array(d["Q_rows"], dtype=np.float32))
    assert np.array_equal(R, np.array(d["R_rows"], dtype=np.float32))
    resid = np.sqrt(np.sum((port.dgemm(Q, R) - A) ** 2, dtype=np.float32))
    assert abs(float(resid) - 3.78809091e-07) < 1e-12


@pytest.mark.parametrize("PR,PC,m,n", CASES)
def test_port_matches_golden_vectors(port, gold, PR, PC, m, n):
    key = f"pr{PR}_pc{PC}_{m}x{n}"
    A = oracle.rand_matrix(m, n, 12)
    rv, tau = port.mmqr(A, PR, PC)
    assert np.array_equal(tau, gold[key + "_tau"])
    assert np.array_equal(np.triu(rv[:n, :])[np.triu_indices(n)], gold[key + "_Rpacked"])
    if key + "_rv" in gold:
        assert np.array_equal(rv, gold[key + "_rv"])
    if key + "_Q" in gold:
        Q, R = port.explicitQR(rv, tau, PR, PC)
        assert np.array_equal(Q, gold[key + "_Q"])
        assert metrics.backward_error(A, Q, R) <= metrics.TOL_BACKWARD
        assert metrics.orthogonality(Q) <= metrics.TOL_ORTH


@pytest.mark.parametrize("PR,PC,m,n,seed", [(4, 2, 10, 6, 1), (4, 2, 34, 34, 2), (64, 4, 184, 64, 3),
                                            (64, 8, 176, 96, 4), (64, 4, 64, 4, 5)])
def test_port_bit_exact_vs_compiled_reference(port, PR, PC, m, n, seed):
    if not oracle.Ref.available(PR, PC):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    assert oracle.legal_shape(m, n, PR, PC)
    ref = oracle.Ref(PR, PC)
    assert ref.panel_dims(m, n) == oracle.panel_dims(m, n, PR, PC)
    A = oracle.rand_matrix(m, n, seed)
    rv, tau = port.mmqr(A, PR, PC)
    rv2, tau2 = ref.mmqr(A)
    assert np.array_equal(rv, rv2) and np.array_equal(tau, tau2)
    Q, R = port.explicitQR(rv, tau, PR, PC)
    Q2, R2 = ref.explicitQR(rv2, tau2)
    assert np.array_equal(Q, Q2) and np.array_equal(R, R2)
    assert np.array_equal(port.dgemm(Q, R), ref.dgemm(Q2, R2))


def test_reference_R_agrees_with_fp64_householder(port):
    """Sanity of the parity metric itself: sign-normalised R of the oracle vs LAPACK fp64."""
    A = oracle.rand_matrix(244, 124, 12)
    rv, _ = port.mmqr(A, 64, 4)
    R64 = np.linalg.qr(A.astype(np.float64), mode="r")
    assert metrics.r_rel_diff(rv, R64) < 1e-5
    assert metrics.gram_error(A, rv) < 1e-5


def test_illegal_shape_is_detected_by_helper():
    assert not oracle.legal_shape(512, 512, 64, 4)   # silently mis-factored by the reference
    assert oracle.legal_shape(512, 512, 64, 8)
    assert oracle.legal_shape(484, 484, 64, 4)
