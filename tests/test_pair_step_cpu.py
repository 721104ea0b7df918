"""CPU checks of the two-pivot-columns-per-exchange panel step (panel_wb2.cu): the fp32 numpy spec of the algebra
(tools/two_column_step.py) and the lane-level replay of the kernel's register layout, masks and cluster exchange maps
(tools/emulate_pair_panel.py) against the plain column-by-column Householder sweep (reflector convention of qr.c:144-152)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import emulate_pair_panel as emu  # noqa: E402
import two_column_step as spec  # noqa: E402

F = np.float32


def _metrics(A, V, tau):
    n = A.shape[1]
    eps = 2.0 ** -23
    Q = spec.q_from(V.astype(np.float64), tau)
    R = np.triu(V[:n].astype(np.float64))
    return (np.linalg.norm(A - Q @ R) / (np.linalg.norm(A) * n * eps), np.linalg.norm(Q.T @ Q - np.eye(n)) / (n * eps))


def test_pair_algebra_matches_single_sweep_and_guard_catches_dependent_columns():
    rng = np.random.default_rng(3)
    A = rng.random((256, 64)).astype(F)
    S, _ = spec.sweep_single(A)
    Pp, tp, fb = spec.sweep_pairs(A)
    assert fb == 0
    assert np.linalg.norm(np.triu(S[:64]) - np.triu(Pp[:64])) <= 1e-5 * np.linalg.norm(np.triu(S[:64]))
    be, orth = _metrics(A, Pp, tp)
    assert be <= 10 and orth <= 10                     # BASELINE north_star bounds
    B = rng.standard_normal((256, 64)).astype(F)
    B[:, 11] = B[:, 10] * F(1.0 + 1e-6)
    B[:, 21] = B[:, 20]
    Pg, tg, fb = spec.sweep_pairs(B)
    assert fb >= 2
    be, orth = _metrics(B, Pg, tg)
    assert be <= 10 and orth <= 10
    Pn, tn, _ = spec.sweep_pairs(B, use_guard=False)   # without the guard the step is NOT usable
    assert max(_metrics(B, Pn, tn)) > 100


def test_kernel_replay_cluster_of_4_ragged_rows():
    rng = np.random.default_rng(5)
    assert emu.check("replay 300x64", rng.random((300, 64)).astype(F), 2, 4)


def test_kernel_replay_forced_fallback_and_two_clusters():
    rng = np.random.default_rng(6)
    assert emu.check("replay 256x64 fallback", rng.random((256, 64)).astype(F), 1, 4, mode=2)
    assert emu.check("replay 250x64 two clusters", rng.random((250, 64)).astype(F), 1, 2, NCL=2)


def _every_pair_at(ratio, seed):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((256, 64)).astype(F)
    for j in range(0, 64, 2):            # odd column nearly parallel to its left neighbour: sin^2(angle) = ratio
        A[:, j + 1] = (np.sqrt(1 - ratio) * A[:, j] + np.sqrt(ratio) * A[:, j + 1]).astype(F)
    return A


def test_guard_threshold_keeps_the_acceptance_bounds_in_the_worst_case():
    """Why the kernel's guard is 0.1 and not the first estimate 1e-3 (profiles/r01_pair_guard_study.txt): with EVERY pair
    just above the guard, the pair sweep must still meet orthogonality <= 10 n eps."""
    for ratio in (0.5, 0.12, 0.1, 0.05, 0.01):
        A = _every_pair_at(ratio, 2)
        P, t, _ = spec.sweep_pairs(A)                       # default guard = the kernel's kPairGuard
        be, orth = _metrics(A, P, t)
        assert be <= 10 and orth <= 10, (ratio, be, orth)
    A = _every_pair_at(1.5e-3, 2)
    P, t, fb = spec.sweep_pairs(A, guard=1e-3)
    assert fb == 0 and _metrics(A, P, t)[1] > 10            # the old guard lets these pairs through and fails


def test_k_column_step_spec_meets_the_bounds_for_k_up_to_8():
    """tools/k_column_step.py: k pivot columns per exchange through k x k Gram algebra (study for a 16-exchange panel
    kernel); random, graded and nearly dependent data, and the in-group chain adversary at the guard."""
    import k_column_step as ks
    rng = np.random.default_rng(8)
    cases = [rng.random((300, 64)), rng.standard_normal((300, 64)) * np.logspace(-6, 6, 64)]
    B = rng.standard_normal((300, 64)); B[:, 11] = B[:, 10] * (1.0 + 1e-6); B[:, 21] = B[:, 20]; B[:, 40] = 0
    cases.append(B)
    C = rng.standard_normal((300, 64))
    for j in range(0, 64, 4):                                # chain adversary just above the guard of 0.25
        for r in range(1, 4):
            C[:, j + r] = np.sqrt(1 - 0.26) * C[:, j + r - 1] + np.sqrt(0.26) * C[:, j + r]
    cases.append(C)
    for A in cases:
        A = A.astype(F)
        for k in (2, 4, 8):
            V, tau, nex = ks.sweep_k(A, k, guard=0.25)
            be, orth = ks.metrics(A, V, tau)
            assert be <= 10 and orth <= 10, (k, be, orth)
            assert nex <= 64
