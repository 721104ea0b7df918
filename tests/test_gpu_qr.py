"""GPU parity tests (pytest -m gpu): every call goes through the C ABI of libcudaqr_b200.so and is
checked against the oracle (oracle/: the reference's algorithm restated, pinned in test_oracle.py),
against golden vectors produced by the unmodified reference, and -- at sizes the oracle cannot
reach -- through size-independent fp64 properties.

Tolerances are BASELINE.json's (north_star): backward error ||A-QR||/(||A|| n eps) <= 10,
orthogonality ||Q^T Q - I||/(n eps) <= 10, sign-normalised R within 1e-4 (normwise) of the
reference's R; eps = 2^-23.
"""
import os

import numpy as np
import pytest

import oracle
from oracle import metrics
from conftest import load_pkg

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def pkg():
    return load_pkg()


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


@pytest.fixture(scope="module")
def ctx(pkg, torch):
    c = pkg.Context(0)
    c.use_torch_stream()
    yield c
    c.close()


def check_factorisation(A, Q, R, r_ref=None):
    """The three north_star acceptance numbers."""
    be = metrics.backward_error(A, Q, R)
    orth = metrics.orthogonality(Q)
    assert be <= metrics.TOL_BACKWARD, f"backward error {be}"
    assert orth <= metrics.TOL_ORTH, f"orthogonality {orth}"
    if r_ref is not None:
        d = metrics.r_rel_diff(R, r_ref)
        assert d <= metrics.TOL_R, f"R differs from the reference's by {d}"
    return be, orth


# ---------------------------------------------------------------------------------------------
# legacy entry points (host buffers), as the reference's own main() drives them (qr.c:461-523)
# ---------------------------------------------------------------------------------------------
def test_legacy_demo_6x4_matches_reference_known_answer(pkg, port):
    A = oracle.rand_matrix(6, 4, 12)
    RV = A.copy(order="F")
    tau = pkg.mmqr(RV)
    assert tau.size == pkg.tau_size(6, 4)
    Q, R = pkg.explicitQR(RV, tau)
    rv_ref, _ = port.mmqr(A, 4, 2)
    check_factorisation(A, Q, R, rv_ref)
    QR = pkg.dgemm(Q, R)
    resid = float(np.sqrt(np.sum((QR - A) ** 2, dtype=np.float32)))
    assert resid < 2e-6   # the reference prints 3.788e-07 for this input (qr.c:515)
    # R above the diagonal, reflectors below: explicit R is exactly triu of the storage
    assert np.array_equal(np.triu(RV)[:4], R[:4])
    assert np.all(R[4:] == 0)


@pytest.mark.parametrize("PR,PC,m,n", [(4, 2, 64, 32), (64, 4, 64, 64), (64, 4, 124, 64), (64, 4, 244, 124),
                                       (64, 8, 512, 512)])
def test_legacy_mmqr_explicitqr_vs_reference_golden(pkg, PR, PC, m, n):
    """R parity against golden vectors written by the UNMODIFIED reference (tests/golden/make_golden.py)."""
    gold = np.load(os.path.join(GOLD, "ref_cases.npz"))
    key = f"pr{PR}_pc{PC}_{m}x{n}"
    A = oracle.rand_matrix(m, n, 12)
    RV = A.copy(order="F")
    tau = pkg.mmqr(RV)
    Q, R = pkg.explicitQR(RV, tau)
    R_ref = np.zeros((n, n), dtype=np.float32)
    R_ref[np.triu_indices(n)] = gold[key + "_Rpacked"]
    check_factorisation(A, Q, R, R_ref)
    assert Q.shape == (m, m) and R.shape == (m, n)


@pytest.mark.parametrize("PR,PC,m,n,seed", [(64, 4, 184, 64, 3), (64, 8, 176, 96, 4), (4, 2, 34, 34, 2),
                                            (64, 4, 484, 484, 7), (64, 4, 1024 - 60 * 0 + 0, 256, 9)])
def test_legacy_vs_oracle_on_seeded_inputs(pkg, port, PR, PC, m, n, seed):
    if not oracle.legal_shape(m, n, PR, PC):
        m = PR + ((m - PR) // (PR - PC)) * (PR - PC)   # snap to the reference's window grid (qr.cu:722-734)
    A = oracle.rand_matrix(m, n, seed)
    rv_ref, _ = port.mmqr(A, PR, PC)
    RV = A.copy(order="F")
    tau = pkg.mmqr(RV)
    Q, R = pkg.explicitQR(RV, tau)
    check_factorisation(A, Q, R, rv_ref)


@pytest.mark.parametrize("m,n", [(1, 1), (5, 1), (7, 7), (100, 37), (300, 70), (257, 129), (1000, 200), (65, 64)])
def test_legacy_ragged_shapes_any_m_ge_n(pkg, m, n):
    """Shapes off the reference's window grid (it silently mis-factors them); here any
    m >= n must work.  Checked against fp64 Householder (numpy/LAPACK)."""
    rng = np.random.default_rng(m * 1000 + n)
    A = np.asfortranarray(rng.random((m, n), dtype=np.float32))
    RV = A.copy(order="F")
    tau = pkg.mmqr(RV)
    Q, R = pkg.explicitQR(RV, tau)
    R64 = np.linalg.qr(A.astype(np.float64), mode="r")
    check_factorisation(A, Q, R, R64)


def test_legacy_edge_cases(pkg):
    # zero matrix: no NaN (the reference divides by a zero norm)
    A = np.zeros((70, 9), dtype=np.float32, order="F")
    RV = A.copy(order="F")
    tau = pkg.mmqr(RV)
    Q, R = pkg.explicitQR(RV, tau)
    assert np.all(np.isfinite(Q)) and np.all(np.isfinite(R)) and np.all(np.isfinite(tau))
    assert metrics.orthogonality(Q) <= metrics.TOL_ORTH
    assert np.abs(Q @ R).max() < 1e-6
    # rank-deficient (duplicated columns) and badly scaled
    rng = np.random.default_rng(3)
    B = rng.standard_normal((200, 20)).astype(np.float32)
    A = np.asfortranarray(np.hstack([B, B, 1e-4 * B[:, :8]]).astype(np.float32))
    RV = A.copy(order="F")
    tau = pkg.mmqr(RV)
    Q, R = pkg.explicitQR(RV, tau)
    assert metrics.backward_error(A, Q, R) <= metrics.TOL_BACKWARD
    assert metrics.orthogonality(Q) <= metrics.TOL_ORTH
    # already upper-triangular input
    A = np.asfortranarray(np.triu(rng.standard_normal((64, 64))).astype(np.float32))
    RV = A.copy(order="F")
    tau = pkg.mmqr(RV)
    Q, R = pkg.explicitQR(RV, tau)
    check_factorisation(A, Q, R, A)


def test_legacy_mmqr_alloc_and_printmat(pkg, port, capfd):
    """The qr.c:55 overload (callee mallocs tau, caller frees) and printMat's text (qr.c:21-33) through the C ABI."""
    import ctypes
    A = oracle.rand_matrix(124, 64, 12)
    RV = A.copy(order="F")
    tau_p = ctypes.POINTER(ctypes.c_float)()
    pkg.lib.mmqr_alloc(RV.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), ctypes.byref(tau_p), 124, 64)
    n_tau = pkg.tau_size(124, 64)
    tau = np.ctypeslib.as_array(tau_p, shape=(n_tau,)).copy()
    ctypes.CDLL(None).free(tau_p)
    RV2 = A.copy(order="F")
    tau2 = pkg.mmqr(RV2)
    assert np.array_equal(RV, RV2) and np.array_equal(tau, tau2)
    Q, R = pkg.explicitQR(RV, tau)
    assert metrics.backward_error(A, Q, R) <= metrics.TOL_BACKWARD
    small = np.asfortranarray(np.array([[1.0, 2.5], [-3.0, 4.0], [0.125, 6.0]], dtype=np.float32))
    pkg.printMat(small)
    ctypes.CDLL(None).fflush(None)
    text = capfd.readouterr().out
    assert text == "Matrix 3 x 2, row by row:\n 1.000000  2.500000 \n-3.000000  4.000000 \n 0.125000  6.000000 \n\n"


def test_legacy_mmqr_chunked_upload_matches_device_path(pkg, torch):
    """Legacy mmqr on PINNED host memory uploads in column chunks that join the factorisation late (driver.cu: catch-up
    streams); the result must agree with the device-resident factorisation of the same matrix and meet the acceptance
    bounds (R against the Gram matrix, Q^T A = R through apply_q)."""
    m = n = 8192
    g = torch.Generator(device="cuda").manual_seed(77)
    A0 = pkg.colmajor(m, n); A0.copy_(torch.rand((m, n), device="cuda", generator=g))
    host = torch.empty((n, m), dtype=torch.float32, pin_memory=True)
    host.copy_(A0.t())
    hnp = host.numpy().T
    tau_h = pkg.mmqr(hnp)
    ctx = pkg.Context(0); ctx.use_torch_stream()
    dA = A0.clone(); tau = torch.zeros(n, device="cuda")
    ctx.geqrf(dA, tau); ctx.synchronize()
    H = torch.from_numpy(np.ascontiguousarray(hnp.T)).cuda().t()          # factored storage from the host path
    d = float((H - dA).norm() / dA.norm())
    assert d < 1e-4, f"chunked-upload factorisation differs from the device path by {d}"   # same algorithm, other GEMM grouping (catch-up slices)
    assert float((torch.from_numpy(tau_h[:n]).cuda() - tau).norm() / tau.norm()) < 1e-4
    Rh = torch.triu(H[:n]).double()
    G = A0.t().double() @ A0.double()
    assert float((Rh.t() @ Rh - G).norm() / G.norm()) < 5e-5
    ctx.close()


@pytest.mark.parametrize("m,n", [(64, 64), (124, 64), (244, 124), (484, 484), (1024 - 60, 256)])
def test_reference_format_export_feeds_the_reference(pkg, port, m, n):
    """SURVEY 8f-2, the strongest parity check: mmqr_reference_format runs the reference's window sweep (PR = 64, PC = 4)
    on the GPU; its storage and tau grid must match what the reference's own mmqr leaves (same algorithm, fp32 rounding
    apart), and the reference's explicitQR (unmodified qr.c when oracle/_ref is built, else the restatement) must turn
    the GPU output back into A = Q R within the acceptance bounds."""
    A = oracle.rand_matrix(m, n, 12)
    RV = A.copy(order="F")
    tau = pkg.mmqr_reference_format(RV)
    ref = oracle.Ref(64, 4) if oracle.Ref.available(64, 4) else None
    rv_ref, tau_ref = ref.mmqr(A) if ref else port.mmqr(A, 64, 4)
    assert tau.size == tau_ref.size
    assert np.linalg.norm(RV - rv_ref) / np.linalg.norm(rv_ref) < 2e-5
    assert np.linalg.norm(tau - tau_ref) / np.linalg.norm(tau_ref) < 2e-5
    assert np.array_equal(tau == 0, tau_ref == 0)                       # same slots used, the rest zero (qr.c:62)
    if m <= 244:                                                           # explicitQR is O(m^3) per reflector (qr.c:415-429)
        Q, R = ref.explicitQR(RV, tau) if ref else port.explicitQR(RV, tau, 64, 4)
        check_factorisation(A, Q, R, r_ref=rv_ref)


def test_reference_format_rejects_shapes_off_the_grid(pkg, torch, ctx):
    A = pkg.colmajor(512, 512); A.fill_(1.0)
    tau = torch.zeros(pkg.tau_size(512, 512), device="cuda")
    assert pkg.lib.cqr_mmqr_reference_format(ctx.h, A.data_ptr(), 512, 512, 512, tau.data_ptr()) == -4   # CQR_EUNSUPPORTED
    with pytest.raises(ValueError):
        pkg.mmqr_reference_format(np.zeros((512, 512), dtype=np.float32, order="F"))


# ---------------------------------------------------------------------------------------------
# double precision (the reference's contemplated `Scalar double`, qr.c:9)
# ---------------------------------------------------------------------------------------------
EPS64 = 2.0 ** -52


@pytest.mark.parametrize("m,n", [(1, 1), (5, 3), (64, 64), (257, 129), (300, 200), (1000, 1000), (5000, 70), (4096, 2048), (20000, 33)])
def test_f64_geqrf_form_q_apply_q(pkg, torch, ctx, m, n):
    """cqr_dgeqrf / cqr_dform_q / cqr_dapply_q / cqr_dextract_r against numpy's fp64 QR: the same three acceptance numbers
    with eps = 2^-52 (backward error and orthogonality <= 10 n eps, R within 1e-12)."""
    rng = np.random.default_rng(123)
    A = np.asfortranarray(rng.random((m, n)))
    dA = pkg.colmajor(m, n, dtype=torch.float64); dA.copy_(torch.from_numpy(A))
    tau = torch.zeros(n, device="cuda", dtype=torch.float64)
    ctx.dgeqrf(dA, tau)
    R = pkg.colmajor(n, n, dtype=torch.float64); ctx.dextract_r(dA, R)
    Q = pkg.colmajor(m, n, dtype=torch.float64); ctx.dform_q(dA, tau, Q)
    C = pkg.colmajor(m, n, dtype=torch.float64); C.copy_(torch.from_numpy(A))
    ctx.dapply_q(dA, tau, C, True)                                  # Q^T A = [R; 0]
    ctx.synchronize()
    Qh, Rh, Ch = Q.cpu().numpy(), R.cpu().numpy(), C.cpu().numpy()
    nrm = np.linalg.norm(A)
    assert np.linalg.norm(A - Qh @ Rh) / (nrm * n * EPS64) <= 10
    assert np.linalg.norm(Qh.T @ Qh - np.eye(n)) / (n * EPS64) <= 10
    assert metrics.r_rel_diff(Rh, np.linalg.qr(A, mode="r")) <= 1e-12
    assert np.linalg.norm(np.triu(Ch[:n]) - Rh) / (nrm * n * EPS64) <= 10
    assert np.sqrt(np.linalg.norm(np.tril(Ch[:n], -1)) ** 2 + np.linalg.norm(Ch[n:]) ** 2) / (nrm * n * EPS64) <= 10


def test_f64_legacy_pair_and_edge_cases(pkg):
    """mmqr_f64 / explicitQR_f64 (host buffers) on the reference's srand(12) input, a zero column, a graded matrix."""
    A32 = oracle.rand_matrix(124, 64, 12)
    for A in (A32.astype(np.float64), np.asfortranarray(np.random.default_rng(5).standard_normal((200, 90)) * np.logspace(0, -12, 90))):
        m, n = A.shape
        if n == 90:
            A[:, 17] = 0.0
            A[:, 40] = A[:, 3]
        RV = A.copy(order="F")
        tau = pkg.mmqr_f64(RV)
        Q, R = pkg.explicitQR_f64(RV, tau)
        assert np.all(np.tril(R, -1) == 0)
        assert np.linalg.norm(A - Q @ R) / (np.linalg.norm(A) * n * EPS64) <= 10
        assert np.linalg.norm(Q.T @ Q - np.eye(m)) / (m * EPS64) <= 10
    assert tau[17] == 0.0                                            # zero column: H = I


def test_comparator_slot_cusolver_agrees(pkg):
    """cqr_compare_cusolver_sgeqrf (the MAGMA slot of qr.cu:555-565 filled with cuSOLVER): when the library is on the box its
    R must agree with ours -- the one place the two are compared; the reference itself never compares."""
    A = oracle.rand_matrix(484, 484, 12)
    CV = A.copy(order="F"); ctau = np.zeros(484, dtype=np.float32)
    import ctypes
    fp = ctypes.POINTER(ctypes.c_float)
    rc = pkg.lib.cqr_compare_cusolver_sgeqrf(CV.ctypes.data_as(fp), ctau.ctypes.data_as(fp), 484, 484)
    if rc != 0:
        assert rc == -4, f"unexpected comparator status {rc}"       # CQR_EUNSUPPORTED: libcusolver not loadable
        pytest.skip("libcusolver is not on this box")
    RV = A.copy(order="F"); pkg.mmqr(RV)
    assert metrics.r_rel_diff(RV, CV) <= metrics.TOL_R


def test_legacy_dgemm_and_identity_vs_oracle(pkg, port):
    rng = np.random.default_rng(0)
    for k, m, n in [(6, 6, 4), (33, 70, 129), (200, 64, 200), (1, 5, 1)]:
        A = np.asfortranarray(rng.standard_normal((k, m)).astype(np.float32))
        B = np.asfortranarray(rng.standard_normal((m, n)).astype(np.float32))
        C = pkg.dgemm(A, B)
        C_ref = port.dgemm(A, B)          # qr.c:443-459 restated
        C64 = A.astype(np.float64) @ B.astype(np.float64)
        scale = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64)
        # both are fp32 sums in different orders: within m*eps of the exact product, like the oracle
        assert np.all(np.abs(C - C64) <= 2 * m * metrics.EPS32 * scale + 1e-30)
        assert np.all(np.abs(C_ref - C64) <= 2 * m * metrics.EPS32 * scale + 1e-30)
    I = pkg.identity(37)
    assert np.array_equal(I, np.eye(37, dtype=np.float32))


# ---------------------------------------------------------------------------------------------
# device-resident API
# ---------------------------------------------------------------------------------------------
def dev(pkg, torch, A_np):
    return pkg.to_colmajor(torch.from_numpy(np.ascontiguousarray(A_np)).cuda())


def host(t):
    return np.asfortranarray(t.cpu().numpy())


@pytest.mark.parametrize("gemm_mode", [0, 1])
@pytest.mark.parametrize("m,n", [(512, 512), (1024, 768), (2048, 2048), (3000, 1000)])
def test_geqrf_device_both_gemm_paths(pkg, torch, ctx, port, gemm_mode, m, n):
    ctx.set_option(pkg.OPT_GEMM, gemm_mode)
    A = oracle.rand_matrix(m, n, 12)
    dA = dev(pkg, torch, A)
    tau = torch.zeros(n, device="cuda")
    ctx.geqrf(dA, tau)
    Q = pkg.colmajor(m, n)
    ctx.form_q(dA, tau, Q)           # thin Q
    R = pkg.colmajor(n, n)
    ctx.extract_r(dA, R)
    ctx.synchronize()
    r_ref = None
    if (m, n) == (512, 512):
        r_ref, _ = port.mmqr(A, 64, 8)
    else:
        r_ref = np.linalg.qr(A.astype(np.float64), mode="r")
    check_factorisation(A, host(Q), host(R), r_ref)
    ctx.set_option(pkg.OPT_GEMM, 1)


@pytest.mark.parametrize("outer", [64, 128, 256, 512])
def test_geqrf_outer_block_widths_agree(pkg, torch, ctx, outer):
    ctx.set_option(pkg.OPT_OUTER_BLOCK, outer)
    m, n = 1536, 1100
    A = oracle.rand_matrix(m, n, 5)
    dA = dev(pkg, torch, A)
    tau = torch.zeros(n, device="cuda")
    ctx.geqrf(dA, tau)
    Q = pkg.colmajor(m, n)
    ctx.form_q(dA, tau, Q)
    R = pkg.colmajor(n, n)
    ctx.extract_r(dA, R)
    ctx.synchronize()
    check_factorisation(A, host(Q), host(R), np.linalg.qr(A.astype(np.float64), mode="r"))
    ctx.set_option(pkg.OPT_OUTER_BLOCK, 256)


@pytest.mark.parametrize("panel_mode", [0, 1])
@pytest.mark.parametrize("m,n", [(1536, 1100), (20000, 64), (5000, 200), (3000, 37), (70000, 64), (40000, 128)])
def test_geqrf_panel_modes(pkg, torch, ctx, panel_mode, m, n):
    """CQR_OPT_PANEL 1 = one-launch multi-CTA Householder panel (1, <=32, >32 and >128 slabs -> fallback),
    0 = TSQR tree + Householder reconstruction; both must meet the north_star tolerances."""
    ctx.set_option(pkg.OPT_PANEL, panel_mode)
    A = oracle.rand_matrix(m, n, 7)
    dA = dev(pkg, torch, A)
    tau = torch.zeros(n, device="cuda")
    ctx.geqrf(dA, tau)
    Q = pkg.colmajor(m, n)
    ctx.form_q(dA, tau, Q)
    R = pkg.colmajor(n, n)
    ctx.extract_r(dA, R)
    ctx.synchronize()
    check_factorisation(A, host(Q), host(R), np.linalg.qr(A.astype(np.float64), mode="r"))
    ctx.set_option(pkg.OPT_PANEL, 1)


@pytest.mark.parametrize("m", [300, 2500, 6000, 9000, 16384, 20000])
def test_geqrf_panel_hh_matches_lapack_storage(pkg, torch, ctx, m):
    """The multi-CTA panel writes LAPACK geqrf storage directly: v below the diagonal and tau must reproduce
    the fp64 Householder vectors of the same sign convention (beta = -sign(alpha) norm, qr.c:149-152), and the
    run must be bitwise reproducible (fixed-order cross-CTA reduction).  Sizes cover one cluster (DSMEM exchange,
    8 .. 64 rows per thread), two clusters (global-flag exchange between them, m > 8192) and the flag-only kernel."""
    n = 64
    rng = np.random.default_rng(3)
    A = np.asfortranarray(rng.standard_normal((m, n)).astype(np.float32))
    outs = []
    for _ in range(2):
        dA = dev(pkg, torch, A)
        tau = torch.zeros(n, device="cuda")
        ctx.geqrf(dA, tau)
        ctx.synchronize()
        outs.append((host(dA), tau.cpu().numpy()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    # fp64 unblocked Householder with the reference's sign convention
    W = A.astype(np.float64)
    taus = np.zeros(n)
    for j in range(n):
        x = W[j:, j].copy()
        nrm = np.linalg.norm(x)
        beta = nrm if x[0] < 0 else -nrm
        u = x[0] - beta
        v = x / u
        v[0] = 1.0
        taus[j] = -u / beta
        W[j:, j:] -= taus[j] * np.outer(v, v @ W[j:, j:])
        W[j + 1:, j] = v[1:]
    got, gtau = outs[0]
    assert np.linalg.norm(gtau - taus) / np.linalg.norm(taus) < 1e-5
    assert np.linalg.norm(np.triu(got[:n]) - np.triu(W[:n])) / np.linalg.norm(np.triu(W[:n])) < 1e-5
    assert np.linalg.norm(np.tril(got, -1) - np.tril(W, -1)) / np.linalg.norm(np.tril(W, -1)) < 1e-4


def test_geqrf_panel_hh_edge_cases(pkg, torch, ctx):
    """Zero column (tau = 0, no NaN: SURVEY App. B5), duplicated columns (rank deficiency) and a length-1 last
    reflector (square matrix: tau = 2 sign flip like the reference)."""
    rng = np.random.default_rng(11)
    A = np.asfortranarray(rng.standard_normal((2500, 96)).astype(np.float32))
    A[:, 5] = 0.0
    A[:, 70] = A[:, 3]
    dA = dev(pkg, torch, A)
    tau = torch.zeros(96, device="cuda")
    ctx.geqrf(dA, tau)
    Q = pkg.colmajor(2500, 96)
    ctx.form_q(dA, tau, Q)
    R = pkg.colmajor(96, 96)
    ctx.extract_r(dA, R)
    ctx.synchronize()
    assert np.isfinite(host(dA)).all() and np.isfinite(tau.cpu().numpy()).all()
    assert float(tau[5]) == 0.0
    check_factorisation(A, host(Q), host(R))
    S = np.asfortranarray(rng.standard_normal((192, 192)).astype(np.float32))
    dS = dev(pkg, torch, S)
    tau = torch.zeros(192, device="cuda")
    ctx.geqrf(dS, tau)
    ctx.synchronize()
    assert float(tau[191]) == 2.0
    Q = pkg.colmajor(192, 192)
    ctx.form_q(dS, tau, Q)
    R = pkg.colmajor(192, 192)
    ctx.extract_r(dS, R)
    ctx.synchronize()
    check_factorisation(S, host(Q), host(R), np.linalg.qr(S.astype(np.float64), mode="r"))


def test_apply_q_and_qt_roundtrip(pkg, torch, ctx):
    m, n, nc = 1500, 300, 77
    rng = np.random.default_rng(1)
    A = np.asfortranarray(rng.standard_normal((m, n)).astype(np.float32))
    C = np.asfortranarray(rng.standard_normal((m, nc)).astype(np.float32))
    dA = dev(pkg, torch, A)
    tau = torch.zeros(n, device="cuda")
    ctx.geqrf(dA, tau)
    dC = dev(pkg, torch, C)
    ctx.apply_q(dA, tau, dC, trans=True)
    QtC = host(dC)
    ctx.apply_q(dA, tau, dC, trans=False)
    ctx.synchronize()
    back = host(dC)
    assert np.linalg.norm(back - C) / np.linalg.norm(C) < 10 * n * metrics.EPS32   # ||Q Q^T - I|| scale
    # Q^T A = R: apply Q^T to A itself
    dA2 = dev(pkg, torch, A)
    ctx.apply_q(dA, tau, dA2, trans=True)
    R = pkg.colmajor(n, n)
    ctx.extract_r(dA, R)
    ctx.synchronize()
    QtA = host(dA2)
    assert np.linalg.norm(QtA[:n] - np.triu(host(R))) / np.linalg.norm(A) < 10 * n * metrics.EPS32
    assert np.linalg.norm(QtA[n:]) / np.linalg.norm(A) < 10 * n * metrics.EPS32
    # least squares through Q^T b and back-substitution agrees with fp64 lstsq
    x_ref = np.linalg.lstsq(A.astype(np.float64), C[:, :3].astype(np.float64), rcond=None)[0]
    x = np.linalg.solve(np.triu(host(R)).astype(np.float64), QtC[:n, :3].astype(np.float64))
    assert np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref) < 1e-3


@pytest.mark.parametrize("m,n,nrhs", [(2000, 700, 5), (4096, 1024, 128), (300, 300, 1), (100, 7, 3)])
def test_solve_least_squares_matches_fp64_lstsq(pkg, torch, ctx, m, n, nrhs):
    """cqr_solve_ls (Q^T b, blocked back substitution) against numpy's fp64 least squares on the same data."""
    rng = np.random.default_rng(21)
    A = np.asfortranarray(rng.random((m, n), dtype=np.float32))
    B = np.asfortranarray(rng.standard_normal((m, nrhs)).astype(np.float32))
    dA, dB = dev(pkg, torch, A), dev(pkg, torch, B)
    tau = torch.zeros(n, device="cuda")
    ctx.geqrf(dA, tau)
    ctx.solve_ls(dA, tau, dB)
    ctx.synchronize()
    X = host(dB)[:n]
    X64 = np.linalg.lstsq(A.astype(np.float64), B.astype(np.float64), rcond=None)[0]
    cond = np.linalg.cond(A.astype(np.float64))
    assert np.linalg.norm(X - X64) / np.linalg.norm(X64) < 50 * cond * metrics.EPS32
    # the normal equations hold to working precision: A^T (A x - b) ~ 0
    res = A.astype(np.float64) @ X.astype(np.float64) - B.astype(np.float64)
    assert np.linalg.norm(A.astype(np.float64).T @ res) / (np.linalg.norm(A) ** 2 * np.linalg.norm(X64)) < 50 * metrics.EPS32
    # exactly singular R is reported, not divided through
    A0 = A.copy(order="F"); A0[:, n - 1] = 0.0
    dA0 = dev(pkg, torch, A0)
    ctx.geqrf(dA0, tau)
    with pytest.raises(pkg.CudaQRError):
        ctx.solve_ls(dA0, tau, dev(pkg, torch, B))


def test_cli_qr_device_matches_reference_interface(pkg):
    """cuda-qr_b200/qr_device (C, legacy entry points only): the reference's size rounding (qr.cu:722-734), result line
    (qr.cu:789) and -- with `check` -- its commented-out residual validation (qr.cu:822-850)."""
    import subprocess
    exe = os.path.join(os.path.dirname(pkg.__file__), "qr_device")
    out = subprocess.run([exe, "512", "512", "check", "compare"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    # the comparator slot (qr.cu:790-806, MAGMA there): cuSOLVER's geqrf timed the same way, or a clear "unavailable"
    assert "cuSOLVER ran QR on 484x484 matrix in" in out.stdout or "cuSOLVER comparator unavailable" in out.stdout
    assert "Exact problem size: 484x484" in out.stdout
    assert "MMQR ran QR on 484x484 matrix in" in out.stdout and "(avg over 3)" in out.stdout
    units = float(out.stdout.split("in units of n*eps:")[1].split(")")[0])
    assert units <= metrics.TOL_BACKWARD
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 1 and "Usage: ./qr_device m n" in out.stdout


# ---------------------------------------------------------------------------------------------
# TSQR (config 3), batched (config 4)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m,n", [(544, 64), (4144, 64), (65536, 64), (100000, 33), (300, 4), (64, 64), (1 << 20, 64),
                                 (70001, 64), (20011, 17)])
def test_tsqr_r_and_thin_q(pkg, torch, ctx, port, m, n):
    if m <= 65536:
        A = oracle.rand_matrix(m, n, 12)
    else:
        A = np.asfortranarray(np.random.default_rng(2).random((m, n), dtype=np.float32))
    dA = dev(pkg, torch, A)
    R1 = pkg.colmajor(n, n)
    ctx.tsqr_r(dA, R1)                   # R-only, A untouched
    ctx.synchronize()
    assert np.array_equal(host(dA), A)
    R2 = pkg.colmajor(n, n)
    ctx.tsqr_factor(dA, R2)              # implicit Q kept
    Q = pkg.colmajor(m, n)
    ctx.tsqr_form_q(Q)
    ctx.synchronize()
    if m < 16384:
        assert np.array_equal(host(R1), host(R2))          # same 256-row tile leaves
    else:
        # both run the warp-resident flat-tree leaf (tsqr_flat.cu); the implicit-Q variant starts every chain from a dense
        # QR of its first block instead of from R = 0, so the two R agree up to row signs and rounding
        assert metrics.r_rel_diff(host(R1), host(R2)) < 2e-5
        # ... and the 256-row tile leaves (CQR_OPT_FLAT_TSQR = 0) give the same R up to row signs and fp32 rounding
        ctx.set_option(pkg.OPT_FLAT_TSQR, 0)
        R3 = pkg.colmajor(n, n)
        dA3 = dev(pkg, torch, A)
        ctx.tsqr_r(dA3, R3)
        R4 = pkg.colmajor(n, n)
        ctx.tsqr_factor(dA3, R4)
        Q4 = pkg.colmajor(m, n)
        ctx.tsqr_form_q(Q4)
        ctx.synchronize()
        ctx.set_option(pkg.OPT_FLAT_TSQR, pkg.TSQR_LEAF_DEFAULT)
        assert np.array_equal(host(R3), host(R4))
        assert metrics.r_rel_diff(host(R1), host(R3)) < 2e-5
        check_factorisation(A, host(Q4), host(R4))
    if oracle.legal_shape(m, n, 64, 4) and m <= 5000:
        r_ref, _ = port.mmqr(A, 64, 4)   # the reference's own flat-tree TSQR on the same input
    else:
        r_ref = np.linalg.qr(A.astype(np.float64), mode="r")
    check_factorisation(A, host(Q), host(R2), r_ref)
    assert metrics.r_rel_diff(host(R1), r_ref) <= metrics.TOL_R
    assert metrics.gram_error(A, host(R1)) < 1e-5


def test_tsqr_flat_leaf_vs_reference_flat_tree(pkg, torch, ctx, port):
    """The warp-resident flat-tree leaf against the reference's own flat tree (60-row windows, qr.c:68-73) on the
    reference's srand(12) input at a shape legal for PR=64/PC=4: same R after sign normalisation."""
    m, n = 64 + 60 * 273, 64                     # 16444 rows: above the flat-leaf threshold
    assert oracle.legal_shape(m, n, 64, 4)
    A = oracle.rand_matrix(m, n, 12)
    dA = dev(pkg, torch, A)
    R = pkg.colmajor(n, n)
    ctx.set_option(pkg.OPT_FLAT_TSQR, pkg.TSQR_LEAF_FLAT)   # this test is about the Householder flat leaf (the default is the Gram leaf)
    l0 = ctx.launch_count()
    ctx.tsqr_r(dA, R)
    ctx.synchronize()
    assert ctx.launch_count() - l0 <= 4          # flat leaf + a short tree, not 7 tile levels
    r_ref, _ = port.mmqr(A, 64, 4)
    assert metrics.r_rel_diff(host(R), r_ref) <= metrics.TOL_R
    assert metrics.r_rel_diff(host(R), np.linalg.qr(A.astype(np.float64), mode="r")) <= 1e-5
    assert metrics.gram_error(A, host(R)) < 1e-5
    # zero columns and a zero block: no NaN, R keeps the zero columns
    A2 = A.copy(order="F")
    A2[:, 5] = 0.0
    A2[4096:8192, :] = 0.0
    ctx.tsqr_r(dev(pkg, torch, A2), R)
    ctx.synchronize()
    Rh = host(R)
    ctx.set_option(pkg.OPT_FLAT_TSQR, pkg.TSQR_LEAF_DEFAULT)
    assert np.isfinite(Rh).all() and np.all(Rh[:, 5][6:] == 0)
    assert metrics.gram_error(A2, Rh) < 1e-5
    # the reference-legal srand(12) input through the default (Gram) leaf: same R as the restated qr.c after sign normalisation
    ctx.tsqr_r(dA, R)
    ctx.synchronize()
    bound, householder = ctx.tsqr_gram_info()
    assert not householder and bound < 32768
    assert metrics.r_rel_diff(host(R), r_ref) <= metrics.TOL_R
    assert metrics.gram_error(A, host(R)) < 1e-6
    # the same input through the default (Gram) leaf: the zero column makes the Cholesky pivot vanish, the device-side
    # gate hands the call to the Householder leaf, and the result is the same
    ctx.tsqr_r(dev(pkg, torch, A2), R)
    ctx.synchronize()
    bound, householder = ctx.tsqr_gram_info()
    assert householder and bound == -1.0
    assert np.array_equal(host(R), Rh)


@pytest.mark.parametrize("leaf", [2, 3])
@pytest.mark.parametrize("m,n,kind", [(16384, 64, "u"), (20011, 17, "u"), (65536, 64, "n"), (100003, 40, "u"), (1 << 20, 64, "u")])
def test_tsqr_r_alternative_leaves(pkg, torch, ctx, m, n, kind, leaf):
    """CQR_OPT_FLAT_TSQR = 2: the flat-tree leaf with its block products on the tensor pipe (tsqr_mma.cu: mma.sync TF32,
    three-product split, every R-sized term on the FMA pipe); = 3: the SIMT leaf with two pivot columns per reduction
    (flat_pair, guard + single-step fallback).  Both against the default SIMT leaf, fp64 and the Gram matrix."""
    g = torch.Generator(device="cuda").manual_seed(9)
    A = pkg.colmajor(m, n)
    A.copy_(torch.rand((m, n), device="cuda", generator=g) if kind == "u" else torch.randn((m, n), device="cuda", generator=g))
    G = A.t().double() @ A.double()
    Rs = {}
    try:
        for mode in (leaf, 1):
            ctx.set_option(pkg.OPT_FLAT_TSQR, mode)
            R = pkg.colmajor(n, n); R.fill_(float("nan"))
            ctx.tsqr_r(A, R); ctx.synchronize()
            assert float(torch.tril(R, -1).abs().max()) == 0.0 if n > 1 else True
            Rd = torch.triu(R.double())
            assert float((Rd.t() @ Rd - G).norm() / G.norm()) < 2e-6
            Rs[mode] = R.cpu().numpy()
    finally:
        ctx.set_option(pkg.OPT_FLAT_TSQR, pkg.TSQR_LEAF_DEFAULT)
    assert metrics.r_rel_diff(Rs[leaf], Rs[1]) <= 1e-5
    if m <= 200000:
        assert metrics.r_rel_diff(Rs[leaf], np.linalg.qr(A.cpu().numpy().astype(np.float64), mode="r")) <= 1e-5



@pytest.mark.parametrize("m,n,kind", [(16384, 64, "u"), (20012, 17, "u"), (65536, 64, "n"), (100004, 40, "u"), (131072 + 128 * 5, 64, "u"),
                                      (262144, 64, "n"), (1 << 20, 64, "u")])
def test_tsqr_gram_leaf(pkg, torch, ctx, m, n, kind):
    """CQR_OPT_FLAT_TSQR = 4 (default): R = chol(A^T A) with the Gram matrix formed error-free on tcgen05 (bf16 slices,
    exact fp32 accumulation, fp64 reduction and Cholesky; gram_umma.cu).  Against the Householder leaf, fp64 and the Gram
    matrix; R has a positive diagonal and zeros below it; two calls give the same bits (fixed-order reductions)."""
    g = torch.Generator(device="cuda").manual_seed(9)
    A = pkg.colmajor(m, n)
    A.copy_(torch.rand((m, n), device="cuda", generator=g) if kind == "u" else torch.randn((m, n), device="cuda", generator=g))
    A0 = A.clone()
    G = A.t().double() @ A.double()
    Rs = {}
    try:
        for mode in (pkg.TSQR_LEAF_GRAM, pkg.TSQR_LEAF_FLAT):
            ctx.set_option(pkg.OPT_FLAT_TSQR, mode)
            R = pkg.colmajor(n, n); R.fill_(float("nan"))
            ctx.tsqr_r(A, R); ctx.synchronize()
            assert float(torch.tril(R, -1).abs().max()) == 0.0 if n > 1 else True
            Rd = torch.triu(R.double())
            assert float((Rd.t() @ Rd - G).norm() / G.norm()) < (2e-7 if mode == pkg.TSQR_LEAF_GRAM else 2e-6)
            Rs[mode] = R.cpu().numpy()
            if mode == pkg.TSQR_LEAF_GRAM:
                bound, householder = ctx.tsqr_gram_info()
                assert not householder and 1.0 <= bound <= 32768.0
                assert bool((torch.diagonal(R) > 0).all())
                R2 = pkg.colmajor(n, n)
                ctx.tsqr_r(A, R2); ctx.synchronize()
                assert torch.equal(R, R2)
        assert torch.equal(A, A0)                                   # R-only: A is never written
    finally:
        ctx.set_option(pkg.OPT_FLAT_TSQR, pkg.TSQR_LEAF_DEFAULT)
    assert metrics.r_rel_diff(Rs[pkg.TSQR_LEAF_GRAM], Rs[pkg.TSQR_LEAF_FLAT]) <= 1e-5
    if m <= 300000:
        r64 = np.linalg.qr(A.cpu().numpy().astype(np.float64), mode="r")
        assert metrics.r_rel_diff(Rs[pkg.TSQR_LEAF_GRAM], r64) <= 2e-7      # fp32 rounding of R: closer to fp64 than the Householder leaf
        assert metrics.r_rel_diff(Rs[pkg.TSQR_LEAF_FLAT], r64) <= 1e-5


def test_tsqr_gram_leaf_is_error_free_on_integer_data(pkg, torch, ctx):
    """Input that fits the first 8-bit slice: the tensor-core Gram matrix must be EXACT, bit for bit (debug export of the
    fp64 matrix the kernel reduced), whatever the rounding mode of the accumulators.  With two or three slices in play
    a 256-row pair's D11/2 + D12 + D13 is rounded to fp32 once before the fp64 sum: 2^-25 of the pair's value, random."""
    import ctypes
    m, n = 65536 + 128, 64
    g = torch.Generator(device="cuda").manual_seed(3)
    for lo, hi, scale in ((0, 256, 1.0 / 256), (-255, 256, 1.0 / 256), (0, 65536, 1.0 / 65536), (-2000, 2001, 1.0)):
        X = torch.randint(lo, hi, (m, n), device="cuda", generator=g).double() * scale
        A = pkg.colmajor(m, n); A.copy_(X.float())
        assert torch.equal(A.double(), X)
        R = pkg.colmajor(n, n)
        ctx.tsqr_r(A, R); ctx.synchronize()
        T = np.zeros((64, 64))
        pkg.lib.cqr_debug_gram_matrix.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        assert pkg.lib.cqr_debug_gram_matrix(ctx.h, T.ctypes.data) == 0
        Gk = T + T.T
        Gx = (X.t() @ X).cpu().numpy()
        if hi == 256:
            assert np.array_equal(Gk, Gx)
        else:
            assert np.abs(Gk - Gx).max() <= 1e-8 * np.abs(Gx).max()
        Rd = torch.triu(R.double())
        assert float((Rd.t() @ Rd - torch.from_numpy(Gx).cuda()).norm() / np.linalg.norm(Gx)) < 1.5e-7


def test_tsqr_gram_leaf_gate(pkg, torch, ctx):
    """The Gram leaf only keeps its R when n ||Rs^-1||_F^2 (>= cond_2 of the unit-diagonal Gram matrix) is at most 32768;
    otherwise -- ill-conditioned, rank-deficient, out-of-scale or non-eligible input -- the Householder leaf enqueued behind
    the device-side gate produces R, with no host synchronisation in between.  Either way R passes the parity bounds."""
    m, n = 131072, 64
    g = torch.Generator(device="cuda").manual_seed(21)
    def graded(cond):
        Q = torch.linalg.qr(torch.randn((m, n), device="cuda", generator=g, dtype=torch.float64)).Q
        V = torch.linalg.qr(torch.randn((n, n), device="cuda", generator=g, dtype=torch.float64)).Q
        sv = torch.logspace(0, -float(np.log10(cond)), n, device="cuda", dtype=torch.float64)
        return (Q * sv) @ V.t()
    def run(X):
        A = pkg.colmajor(X.shape[0], X.shape[1]); A.copy_(X.float())
        R = pkg.colmajor(X.shape[1], X.shape[1])
        ctx.tsqr_r(A, R); ctx.synchronize()
        return A, R
    for cond, expect_householder in ((3.0, False), (10.0, False), (100.0, True), (1e4, True)):
        A, R = run(graded(cond))
        bound, householder = ctx.tsqr_gram_info()
        assert householder == expect_householder and bound >= cond * cond * 0.5
        r64 = np.linalg.qr(A.double().cpu().numpy(), mode="r")
        assert metrics.r_rel_diff(R.cpu().numpy(), r64) <= (1e-6 if not householder else metrics.TOL_R)
        assert metrics.gram_error(A.cpu().numpy(), R.cpu().numpy()) < 1e-6
    # exactly dependent columns, a zero column, values the fp32 accumulators cannot carry: Cholesky / scale guard -> Householder
    X = torch.rand((m, n), device="cuda", generator=g); X[:, 7] = X[:, 3]
    A, R = run(X)
    assert ctx.tsqr_gram_info() [1] and metrics.gram_error(A.cpu().numpy(), R.cpu().numpy()) < 1e-6
    X = torch.rand((m, n), device="cuda", generator=g); X[:, 9] = 0
    A, R = run(X)
    assert ctx.tsqr_gram_info()[1] and metrics.gram_error(A.cpu().numpy(), R.cpu().numpy()) < 1e-6
    X = torch.rand((m, n), device="cuda", generator=g) * 1e-14      # 2^-46: below the 2^-40 the accumulators are allowed; fp32 squares still normal
    A, R = run(X)
    assert ctx.tsqr_gram_info() == (-1.0, True)
    G = A.double().t() @ A.double()
    Rd = torch.triu(R.double())
    assert float((Rd.t() @ Rd - G).norm() / G.norm()) < 1e-6
    # a leading dimension that TMA cannot address (lda % 4 != 0): the Gram leaf is not used at all
    X = torch.rand((m - 1, n), device="cuda", generator=g)
    A, R = run(X)
    with pytest.raises(pkg.CudaQRError):
        ctx.tsqr_gram_info()
    assert metrics.gram_error(A.cpu().numpy(), R.cpu().numpy()) < 1e-6
    # column scales spread over 2^-20 .. 2^20: diagonal scaling does not hurt Cholesky, and the slices are per column
    X = torch.randn((m, n), device="cuda", generator=g) * (2.0 ** torch.linspace(-20, 20, n, device="cuda"))
    A, R = run(X)
    bound, householder = ctx.tsqr_gram_info()
    assert not householder and bound < 8192
    r64 = np.linalg.qr(A.double().cpu().numpy(), mode="r")
    assert metrics.r_rel_diff(R.cpu().numpy(), r64) <= 1e-6


def test_tsqr_pair_leaf_guard_fallback(pkg, torch, ctx):
    """Neighbouring columns that are nearly (or exactly) dependent trip the cancellation guard of the two-column leaf
    (sigma_2 < 0.1 q): the pair must fall back to two single steps and still agree with the single-column leaf."""
    m, n = 40000, 64
    g = torch.Generator(device="cuda").manual_seed(4)
    base = torch.randn((m, n), device="cuda", generator=g)
    A = pkg.colmajor(m, n)
    A.copy_(base)
    for j in range(0, n, 2):                                     # every pair: column j+1 = column j + a small perturbation
        A[:, j + 1] = A[:, j] * (1.0 + 0.01 * j) + 1e-3 * base[:, j + 1]
    A[:, 9] = A[:, 8]                                            # one exact duplicate, one zero column
    A[:, 30] = 0.0
    G = A.t().double() @ A.double()
    Rs = {}
    try:
        for mode in (3, 1):
            ctx.set_option(pkg.OPT_FLAT_TSQR, mode)
            R = pkg.colmajor(n, n)
            ctx.tsqr_r(A, R); ctx.synchronize()
            Rd = torch.triu(R.double())
            assert float((Rd.t() @ Rd - G).norm() / G.norm()) < 2e-6
            Rs[mode] = Rd
    finally:
        ctx.set_option(pkg.OPT_FLAT_TSQR, pkg.TSQR_LEAF_DEFAULT)
    # R itself is ill-determined for dependent columns; the Gram matrices above are the check, plus agreement on the
    # well-determined entries: first column of each pair against the reflectors of first columns (an odd row belongs to
    # the direction of a 1e-3 perturbation, which is only determined to ~1e-3 / eps)
    d = (Rs[3].abs() - Rs[1].abs())[::2, ::2]
    assert float(d.norm() / Rs[1][::2, ::2].norm()) < 1e-3


def test_tsqr_seeded_form_q_is_linear(pkg, torch, ctx):
    """form_q(X) = Q X: the hook the multi-GPU R-tree uses to push its own Q blocks down."""
    m, n = 5000, 48
    rng = np.random.default_rng(4)
    A = np.asfortranarray(rng.standard_normal((m, n)).astype(np.float32))
    X = np.asfortranarray(rng.standard_normal((n, n)).astype(np.float32))
    dA = dev(pkg, torch, A)
    R = pkg.colmajor(n, n)
    ctx.tsqr_factor(dA, R)
    Q = pkg.colmajor(m, n)
    ctx.tsqr_form_q(Q)
    QX = pkg.colmajor(m, n)
    ctx.tsqr_form_q(QX, dev(pkg, torch, X))
    ctx.synchronize()
    want = host(Q).astype(np.float64) @ X.astype(np.float64)
    assert np.linalg.norm(host(QX) - want) / np.linalg.norm(want) < 20 * metrics.EPS32


def test_stack_qr_combines_two_r_factors(pkg, torch, ctx):
    n = 64
    rng = np.random.default_rng(5)
    R1 = np.triu(rng.standard_normal((n, n))).astype(np.float32)
    R2 = np.triu(rng.standard_normal((n, n))).astype(np.float32)
    S = np.asfortranarray(np.vstack([R1, R2]))
    dS = dev(pkg, torch, S)
    tau = torch.zeros(64, device="cuda")
    R = pkg.colmajor(n, n)
    ctx.stack_qr(dS, n, tau, R)
    Qs = pkg.colmajor(2 * n, n)
    ctx.stack_form_q(dS, n, tau, Qs)
    ctx.synchronize()
    check_factorisation(S, host(Qs), host(R), np.linalg.qr(S.astype(np.float64), mode="r"))


@pytest.mark.parametrize("m,n,batch", [(64, 64, 1000), (64, 64, 65536), (100, 40, 300), (256, 64, 50), (8, 8, 10)])
def test_batched_geqrf(pkg, torch, ctx, port, m, n, batch):
    g = torch.Generator(device="cuda").manual_seed(12)
    A3 = torch.rand((batch, n, m), device="cuda", generator=g)      # A3[b].T is matrix b (column-major, lda = m)
    if (m, n) == (64, 64):
        A3[0].copy_(torch.from_numpy(oracle.rand_matrix(64, 64, 12).T.copy()).cuda())
    orig = A3.clone()
    tau = torch.zeros((batch, n), device="cuda")
    ctx.geqrf_batched(A3, tau)
    ctx.synchronize()
    # fp64 check of a sample of matrices (all of them when the batch is small)
    idx = list(range(batch)) if batch <= 300 else [0, 1, batch // 2, batch - 1] + list(range(7, batch, batch // 50))
    for b in idx:
        A = orig[b].t().cpu().numpy()
        V = A3[b].t().cpu().numpy()
        Q = metrics.householder_q(V, tau[b].cpu().numpy(), full=False)
        R = np.triu(V[:n])
        r_ref = np.linalg.qr(A.astype(np.float64), mode="r")
        if b == 0 and (m, n) == (64, 64):
            r_ref, _ = port.mmqr(np.asfortranarray(A), 64, 4)       # the reference's single-window case
        check_factorisation(A, Q, R, r_ref)
    # whole-batch property on device: R^T R == A^T A for every matrix (fp64 on the GPU)
    Rd = torch.triu(A3.transpose(1, 2)[:, :n, :].double())
    G = orig.double() @ orig.double().transpose(1, 2)
    err = (Rd.transpose(1, 2) @ Rd - G).flatten(1).norm(dim=1) / G.flatten(1).norm(dim=1)
    assert float(err.max()) < 1e-5


def test_random_shapes_all_entry_points(pkg, torch, ctx):
    """Seeded sweep over ragged shapes (odd leading dimensions included through row slicing): every device entry point
    against fp64 on the same data -- geqrf + form_q, R-only / implicit-Q TSQR on both leaf kinds, batched."""
    rng = np.random.default_rng(2026)
    for _ in range(6):                                        # blocked Householder, any m >= n
        n = int(rng.integers(1, 700)); m = n + int(rng.integers(0, 900))
        A = np.asfortranarray(rng.standard_normal((m, n)).astype(np.float32))
        dA = dev(pkg, torch, A); tau = torch.zeros(n, device="cuda")
        ctx.geqrf(dA, tau)
        Q = pkg.colmajor(m, n); ctx.form_q(dA, tau, Q)
        R = pkg.colmajor(n, n); ctx.extract_r(dA, R)
        ctx.synchronize()
        check_factorisation(A, host(Q), host(R), np.linalg.qr(A.astype(np.float64), mode="r"))
    for _ in range(6):                                        # tall-skinny: below and above the flat-leaf threshold
        n = int(rng.integers(1, 65)); m = int(rng.integers(max(n, 200), 60000))
        A = np.asfortranarray(rng.random((m, n), dtype=np.float32))
        big = pkg.colmajor(m + 3, n); big.zero_()
        dA = big[1:m + 1]                                     # ld = m + 3 (not a multiple of 4 in general), offset base
        dA.copy_(torch.from_numpy(A).cuda())
        R1 = pkg.colmajor(n, n); ctx.tsqr_r(dA, R1)
        R2 = pkg.colmajor(n, n); ctx.tsqr_factor(dA, R2)
        Q = pkg.colmajor(m, n); ctx.tsqr_form_q(Q)
        ctx.synchronize()
        r_ref = np.linalg.qr(A.astype(np.float64), mode="r")
        check_factorisation(A, host(Q), host(R2), r_ref)
        assert metrics.r_rel_diff(host(R1), r_ref) <= 1e-5
        assert float(big[0].abs().max()) == 0.0 and float(big[m + 1:].abs().max()) == 0.0   # nothing written outside
    for _ in range(4):                                        # batched, m <= 64 (warp kernel) with padding between matrices
        n = int(rng.integers(1, 65)); m = int(rng.integers(n, 65)); batch = int(rng.integers(1, 40)); lda = m + int(rng.integers(0, 5))
        buf = torch.rand((batch, n, lda), device="cuda")
        orig = buf.clone()
        tau = torch.zeros((batch, n), device="cuda")
        pkg._check(pkg.lib.cqr_geqrf_batched(ctx.h, pkg._dptr(buf), lda, n * lda, m, n, batch, pkg._dptr(tau)), "cqr_geqrf_batched")
        ctx.synchronize()
        assert torch.equal(buf[:, :, m:], orig[:, :, m:])      # padding rows untouched
        for b in range(batch):
            A = orig[b, :, :m].t().cpu().numpy()
            V = buf[b, :, :m].t().cpu().numpy()
            Qb = metrics.householder_q(V, tau[b].cpu().numpy(), full=False)
            check_factorisation(A, Qb, np.triu(V[:n]), np.linalg.qr(A.astype(np.float64), mode="r"))


def test_badly_scaled_and_rank_deficient_inputs(pkg, torch, ctx):
    """Householder QR is backward stable whatever the conditioning: graded column scales (1e-6 .. 1e6), duplicated
    columns (exact rank deficiency) and an all-zero matrix through the blocked path, both TSQR leaves and the batched
    kernel -- backward error and orthogonality within the north_star bounds, never a NaN (the reference divides by a zero
    norm there)."""
    rng = np.random.default_rng(77)
    def graded(m, n):
        A = rng.standard_normal((m, n)) * np.logspace(-6, 6, n)[None, :]
        A[:, n // 2] = A[:, n // 3]                                     # exact duplicate column
        return np.asfortranarray(A.astype(np.float32))
    # blocked Householder
    A = graded(1500, 320)
    dA = dev(pkg, torch, A); tau = torch.zeros(320, device="cuda")
    ctx.geqrf(dA, tau)
    Q = pkg.colmajor(1500, 320); ctx.form_q(dA, tau, Q)
    R = pkg.colmajor(320, 320); ctx.extract_r(dA, R)
    ctx.synchronize()
    assert np.isfinite(host(Q)).all() and np.isfinite(host(R)).all()
    check_factorisation(A, host(Q), host(R))
    # duplicates and a zero column inside one 64-column panel, plus a zero row block
    A = np.asfortranarray(rng.standard_normal((2000, 256)).astype(np.float32))
    A[:, 9] = A[:, 5]; A[:, 20] = 0.0; A[:, 70] = 2.0 * A[:, 66]; A[512:1024, :] = 0.0
    dA = dev(pkg, torch, A); tau = torch.zeros(256, device="cuda")
    ctx.geqrf(dA, tau)
    Q = pkg.colmajor(2000, 256); ctx.form_q(dA, tau, Q)
    R = pkg.colmajor(256, 256); ctx.extract_r(dA, R)
    ctx.synchronize()
    assert np.isfinite(host(Q)).all() and np.isfinite(host(R)).all()
    check_factorisation(A, host(Q), host(R))
    # TSQR: tile leaves (m < 16384) and flat leaves
    for m in (3000, 40000):
        A = graded(m, 64)
        A[:, 7] = 0.0; A[:, 12] = A[:, 11]                                # also inside the first 8-column slot group
        dA = dev(pkg, torch, A)
        R1 = pkg.colmajor(64, 64); ctx.tsqr_r(dA, R1)
        R2 = pkg.colmajor(64, 64); ctx.tsqr_factor(dA, R2)
        Qt = pkg.colmajor(m, 64); ctx.tsqr_form_q(Qt)
        ctx.synchronize()
        assert np.isfinite(host(Qt)).all() and np.isfinite(host(R1)).all()
        check_factorisation(A, host(Qt), host(R2))
        # columnwise: |R^T R - A^T A| relative to the column norms (a normwise Gram check would hide the small columns)
        G = A.astype(np.float64).T @ A.astype(np.float64)
        Rd = np.triu(host(R1).astype(np.float64))
        cn = np.sqrt(np.diag(G)); cn[cn == 0] = 1.0                      # the zero column: absolute check
        assert np.max(np.abs(Rd.T @ Rd - G) / np.outer(cn, cn)) < 1e-4
    # batched
    A3 = torch.from_numpy(np.stack([graded(64, 64).T.copy() for _ in range(8)])).cuda()
    A3[3].zero_()                                                        # one all-zero matrix
    orig = A3.clone(); tb = torch.zeros((8, 64), device="cuda")
    ctx.geqrf_batched(A3, tb)
    ctx.synchronize()
    assert bool(torch.isfinite(A3).all()) and bool(torch.isfinite(tb).all())
    assert float(A3[3].abs().max()) == 0.0 and float(tb[3].abs().max()) == 0.0
    for b in (0, 5):
        Ab = orig[b].t().cpu().numpy(); V = A3[b].t().cpu().numpy()
        check_factorisation(Ab, metrics.householder_q(V, tb[b].cpu().numpy(), full=False), np.triu(V))
    # all-zero tall matrix through the flat leaf: R = 0, no NaN
    Z = pkg.colmajor(20000, 64); Z.zero_()
    Rz = pkg.colmajor(64, 64); ctx.tsqr_r(Z, Rz); ctx.synchronize()
    assert float(Rz.abs().max()) == 0.0


# ---------------------------------------------------------------------------------------------
# BASELINE-size property checks (the oracle cannot reach these)
# ---------------------------------------------------------------------------------------------
def test_tsqr_full_size_8m_by_64_properties(pkg, torch, ctx):
    m, n = 8388608, 64
    g = torch.Generator(device="cuda").manual_seed(12)
    A = pkg.colmajor(m, n)
    A.copy_(torch.rand((m, n), device="cuda", generator=g))
    R = pkg.colmajor(n, n)
    ctx.tsqr_r(A, R)
    ctx.synchronize()
    G = (A.t().double() @ A.double())
    Rd = torch.triu(R.double())
    gram = float((Rd.t() @ Rd - G).norm() / G.norm())
    assert gram < 1e-5
    # thin Q: orthogonality and A = Q R, evaluated on device in fp64
    A0 = A.clone()
    ctx.tsqr_factor(A, R)
    Q = pkg.colmajor(m, n)
    ctx.tsqr_form_q(Q)
    ctx.synchronize()
    orth = float(((Q.t().double() @ Q.double()) - torch.eye(n, device="cuda", dtype=torch.float64)).norm()) / (n * metrics.EPS32)
    be = float((A0.double() - Q.double() @ torch.triu(R.double())).norm() / A0.double().norm()) / (n * metrics.EPS32)
    assert orth <= metrics.TOL_ORTH and be <= metrics.TOL_BACKWARD


@pytest.mark.parametrize("size", [8192, 16384])
def test_square_properties_up_to_the_baseline_size(pkg, torch, ctx, size):
    """Config 2's code path at 8192^2 and at the BASELINE size 16384^2 with the full acceptance metrics evaluated on the
    device in fp64, column block by column block (the oracle cannot reach these sizes)."""
    m = n = size
    g = torch.Generator(device="cuda").manual_seed(12)
    A = pkg.colmajor(m, n)
    A.copy_(torch.rand((m, n), device="cuda", generator=g))
    A0 = A.clone()
    tau = torch.zeros(n, device="cuda")
    ctx.geqrf(A, tau)
    R = pkg.colmajor(n, n)
    ctx.extract_r(A, R)
    QRfull = pkg.colmajor(m, n)                  # Q (R) via apply_q: no dense Q needed
    QRfull.copy_(R)
    ctx.apply_q(A, tau, QRfull, trans=False)
    ctx.synchronize()
    num = torch.zeros((), device="cuda", dtype=torch.float64)
    den = torch.zeros((), device="cuda", dtype=torch.float64)
    gnum = torch.zeros((), device="cuda", dtype=torch.float64)
    gden = torch.zeros((), device="cuda", dtype=torch.float64)
    A0d, Rd = A0.double(), R.double()
    for c0 in range(0, n, 2048):
        blk = slice(c0, c0 + 2048)
        d = A0d[:, blk] - QRfull[:, blk].double()
        num += (d * d).sum(); den += (A0d[:, blk] ** 2).sum()
        G = A0d.t() @ A0d[:, blk]                # Gram check R^T R = A^T A, one block of columns at a time
        dG = Rd.t() @ Rd[:, blk] - G
        gnum += (dG * dG).sum(); gden += (G * G).sum()
    be = float((num / den).sqrt()) / (n * metrics.EPS32)
    assert be <= metrics.TOL_BACKWARD
    assert float((gnum / gden).sqrt()) / (n * metrics.EPS32) <= metrics.TOL_BACKWARD
    del A0d, Rd
    Q = pkg.colmajor(m, 256)
    ctx.form_q(A, tau, Q)                        # first 256 columns of Q
    ctx.synchronize()
    orth = float((Q.t().double() @ Q.double() - torch.eye(256, device="cuda", dtype=torch.float64)).norm()) / (256 * metrics.EPS32)
    assert orth <= metrics.TOL_ORTH


@pytest.mark.parametrize("m,n,nf", [(3000, 1000, 256), (16384, 2048, 256), (2048, 4096, 256), (5000, 900, 512), (700, 300, 100),
                                    (512, 1200, 512),
                                    # narrow factor parts (nf < 64): the block scratch must cover n - nf columns
                                    (400, 48, 8), (600, 64, 32), (2000, 700, 8), (4096, 1300, 32), (300, 40, 24)])
def test_geqrf_partial_equals_factor_then_apply(pkg, torch, ctx, m, n, nf):
    """cqr_geqrf_partial (QR of the first nf columns, Q^T applied to all n; n may exceed m) against cqr_geqrf on the
    first nf columns followed by cqr_apply_q on the rest, and against fp64: [Q^T A](:, nf:) and R of A(:, :nf)."""
    rng = np.random.default_rng(31)
    A = np.asfortranarray(rng.standard_normal((m, n)).astype(np.float32))
    d1 = dev(pkg, torch, A); t1 = torch.zeros(nf, device="cuda")
    ctx.geqrf_partial(d1, t1, nf)
    d2 = dev(pkg, torch, A); t2 = torch.zeros(nf, device="cuda")
    ctx.geqrf(d2[:, :nf], t2)
    ctx.apply_q(d2[:, :nf], t2, d2[:, nf:], trans=True)
    ctx.synchronize()
    H1, H2 = host(d1), host(d2)
    # same panels, same reflectors (the K = 64 updates inside the block may pick different split-K counts on the panel partition)
    assert np.linalg.norm(H1[:, :nf] - H2[:, :nf]) / np.linalg.norm(H2[:, :nf]) < 5e-6
    assert np.linalg.norm(host(t1) - host(t2)) / np.linalg.norm(host(t2)) < 5e-6
    dq = np.linalg.norm(H1[:, nf:] - H2[:, nf:]) / np.linalg.norm(H2[:, nf:])
    assert dq < 2e-5, f"Q^T C differs by {dq}"                                              # same Q^T C up to GEMM grouping (K = 64 x 4 vs K = 256)
    Q64, R64 = np.linalg.qr(A[:, :nf].astype(np.float64), mode="complete")
    sgn = np.sign(np.diag(R64[:nf])) * np.sign(np.diag(H1[:nf, :nf]))
    want = (Q64.T @ A[:, nf:].astype(np.float64))
    want[:nf] *= sgn[:, None]                                                                   # row signs of the first nf rows follow R's
    got = H1[:, nf:].astype(np.float64)
    assert np.linalg.norm(got[:nf] - want[:nf]) / np.linalg.norm(want[:nf]) < 1e-4
    # the rows below nf are only determined up to the orthogonal completion: compare their Gram matrix
    assert np.linalg.norm(got[nf:].T @ got[nf:] - want[nf:].T @ want[nf:]) / max(1e-30, np.linalg.norm(want[nf:].T @ want[nf:])) < 1e-4


def test_caqr_single_rank_blocks_match_fp64(pkg, torch, ctx):
    """cuda-qr_b200/dist_caqr.py with world = 1: the per-block local steps (cqr_geqrf / cqr_apply_q on sub-views of
    the local slab) must reproduce the fp64 R; the cross-rank steps are covered by tools/check_dist_caqr.py on 2 GPUs
    and by the gloo schedule tests."""
    import importlib
    dc = importlib.import_module("cuda-qr_b200.dist_caqr")
    m, n = 3000, 700
    A = oracle.rand_matrix(m, n, 21)
    dA = dev(pkg, torch, A)
    cq = dc.DistCAQR(pkg, ctx, m, n, 0, 1, dA.device, kb=256)
    cq.factor(dA)
    R = pkg.colmajor(n, n)
    cq.extract_r(dA, R)
    ctx.synchronize()
    assert metrics.r_rel_diff(host(R), np.linalg.qr(A.astype(np.float64), mode="r")) <= 1e-4


def test_profile_timeline_brackets_are_ordered(pkg, torch, ctx):
    m = n = 1024
    A = oracle.rand_matrix(m, n, 5)
    dA = dev(pkg, torch, A)
    tau = torch.zeros(n, device="cuda")
    ctx.profile_begin()
    ctx.geqrf(dA, tau)
    tl = ctx.profile_timeline()
    prof = ctx.profile_end()
    assert len(tl) > 0 and all(t1 >= t0 >= 0.0 for t0, t1, _ in tl)
    assert {c for _, _, c in tl} <= set(pkg.Context.PROF_CLASSES)
    assert abs(sum(t1 - t0 for t0, t1, _ in tl) - sum(v["ms"] for v in prof.values())) < 1e-3 * max(1.0, len(tl))
    # errors are reported as the (negative) status codes, not as a positive bracket count
    import ctypes
    z = (ctypes.c_double * 1)()
    assert pkg.lib.cqr_profile_timeline(ctx.h, z, z, (ctypes.c_int * 1)(), -1) < 0
    assert pkg.lib.cqr_profile_timeline(ctx.h, None, z, (ctypes.c_int * 1)(), 1) < 0


def test_geqrf_pair_step_cancellation_fallback(pkg, torch, ctx):
    """Panels of 2048 .. 8192 rows take two pivot columns per cluster exchange (panel_wb2.cu).  Nearly dependent
    NEIGHBOURING columns make the algebraic update of the pair cancel; the kernel must notice (sigma_2 guard) and fall
    back to two single steps.  Same acceptance numbers as everywhere else; tools/two_column_step.py shows them at
    1e4 n eps without the guard."""
    rng = np.random.default_rng(21)
    A = np.asfortranarray(rng.standard_normal((3000, 128)).astype(np.float32))
    A[:, 11] = A[:, 10] * np.float32(1.0 + 1e-6)
    A[:, 21] = A[:, 20]
    A[:, 64] = A[:, 65] + np.float32(1e-4) * A[:, 66]
    dA = dev(pkg, torch, A)
    tau = torch.zeros(128, device="cuda")
    ctx.geqrf(dA, tau)
    Q = pkg.colmajor(3000, 128)
    ctx.form_q(dA, tau, Q)
    R = pkg.colmajor(128, 128)
    ctx.extract_r(dA, R)
    ctx.synchronize()
    assert np.isfinite(host(dA)).all() and np.isfinite(tau.cpu().numpy()).all()
    check_factorisation(A, host(Q), host(R))


_FUSED_SNIPPET = r"""
import importlib, sys
import numpy as np, torch
sys.path.insert(0, {root!r})
pkg = importlib.import_module("cuda-qr_b200")
ctx = pkg.Context(0); ctx.use_torch_stream()
m, n = {m}, {n}
A = pkg.colmajor(m, n); A.copy_(torch.rand((m, n), device="cuda", generator=torch.Generator(device="cuda").manual_seed(31)))
tau = torch.zeros(n, device="cuda")
ctx.geqrf(A, tau); ctx.synchronize()
np.save({out!r}, A[:n].cpu().numpy())
print("partition", ctx.get_option(pkg.OPT_PARTITION), "launches", ctx.launch_count())
"""


@pytest.mark.parametrize("m,n", [(3000, 1536), (12000, 2048)])
def test_one_launch_chain_update_agrees_with_three_launch_path(pkg, torch, tmp_path, m, n):
    """chain_update.cu (W = V^T C, X = T^T W, C -= V X in one launch with two grid barriers; fp32 FMA) replaces the three
    tcgen05 launches for the chain's inner updates (<= 4096 rows) and for the panel-wise look-ahead slices.  The same
    factorisation with CQR_CHAIN_FUSED=0 (the switch is read once per process, hence the subprocesses) must give the same
    R to fp32 accuracy, with fewer launches; both against the Gram matrix.  Reference step: trailingUpdateKernel, qr.cu:335-465."""
    import subprocess
    import sys
    from conftest import ROOT
    res = {}
    for mode in ("1", "0"):
        out = str(tmp_path / f"r_{mode}.npy")
        env = dict(os.environ, CQR_CHAIN_FUSED=mode)
        p = subprocess.run([sys.executable, "-c", _FUSED_SNIPPET.format(root=ROOT, m=m, n=n, out=out)], capture_output=True, text=True,
                           timeout=600, env=env)
        assert p.returncode == 0, p.stderr[-2000:]
        words = p.stdout.split()
        res[mode] = (np.triu(np.load(out)).astype(np.float64), int(words[1]), int(words[3]))
    (R1, part, l1), (R0, _, l0) = res["1"], res["0"]
    g = torch.Generator(device="cuda").manual_seed(31)
    A = torch.rand((m, n), device="cuda", generator=g).double()
    G = (A.t() @ A).cpu().numpy()
    for R in (R1, R0):
        assert np.linalg.norm(R.T @ R - G) / np.linalg.norm(G) < 1e-4          # 3xTF32 trailing updates: 4.4e-5 at 12000 x 2048
    assert np.linalg.norm(np.abs(R1) - np.abs(R0)) / np.linalg.norm(R0) < 1e-4
    if part:                                                    # green contexts available: the one-launch form was really used
        assert l1 < l0
