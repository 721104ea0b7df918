"""Numerical spec of the two-pivot-columns-per-exchange Householder step planned for the panel kernel.

One "exchange" delivers, for every column c, the sums p_c = x^T a_c and q_c = y^T a_c over the rows BELOW row j+1
(x = column j, y = column j+1, both as they are before reflector j touches anything) plus rows j and j+1 of the panel.
Everything else -- both reflectors' scalars, every column's two inner products, the update coefficients and the two new
columns of V^T V -- follows algebraically.  This script checks the algebra against the plain column-by-column Householder
sweep in fp32 (same reflector convention as the kernels: beta = -sign(alpha) * norm, u = alpha - beta, tau = -u / beta)
and shows why the cancellation guard on sigma_2 is mandatory.   python tools/two_column_step.py [guard_study]"""
import sys

import numpy as np

F = np.float32


def scalars(alpha, sigma):
    """reflector scalars from the pivot alpha and sigma = sum of squares below it"""
    sj = F(alpha * alpha + sigma)
    if sj < F(1.2e-38):
        return alpha, F(0), F(0)                      # beta, 1/u, tau: H = I
    nrm = F(np.sqrt(sj))
    beta = nrm if alpha < 0 else -nrm
    u = F(alpha - beta)
    return beta, F(1) / u, F(-u / beta)


def sweep_single(A):
    """reference: one column per step; returns the factored panel (R on/above the diagonal, v below) and tau"""
    A = A.astype(F).copy()
    m, n = A.shape
    tau = np.zeros(n, F)
    for j in range(n):
        x = A[j + 1:, j].copy()
        beta, iu, t = scalars(A[j, j], F(x @ x))
        tau[j] = t
        if t != 0:
            for c in range(j + 1, n):
                d = F(A[j, c] + F(x @ A[j + 1:, c]) * iu)
                w = F(t * d)
                A[j, c] -= w
                A[j + 1:, c] -= F(w * iu) * x
            A[j, j] = beta
            A[j + 1:, j] = x * iu
    return A, tau


def sweep_pairs(A, guard=0.1, use_guard=True):
    """two columns per exchange; n even"""
    A = A.astype(F).copy()
    m, n = A.shape
    tau = np.zeros(n, F)
    fallbacks = 0
    for j in range(0, n, 2):
        xh, yh = A[j + 2:, j].copy(), A[j + 2:, j + 1].copy()        # rows below j+1
        # ---- what the exchange delivers
        p = (xh @ A[j + 2:, :]).astype(F)                             # x^T a_c  (all columns)
        q = (yh @ A[j + 2:, :]).astype(F)                             # y^T a_c
        r1, r2 = A[j, :].copy(), A[j + 1, :].copy()                   # rows j, j+1
        xj1, yj1 = r2[j], r2[j + 1]
        # ---- reflector j
        beta1, iu1, t1 = scalars(r1[j], F(p[j] + xj1 * xj1))
        # column j+1 under H_j
        d1n = F(r1[j + 1] + F(p[j + 1] + xj1 * yj1) * iu1)
        t = F(t1 * d1n)
        a1 = F(t * iu1)
        alpha2 = F(yj1 - a1 * xj1)
        sig2 = F(q[j + 1] - F(2) * a1 * p[j + 1] + a1 * a1 * p[j])    # ||y'||^2 below row j+1, by expansion
        if use_guard and sig2 < F(guard) * q[j + 1]:
            # cancellation: y' is tiny against y, so sigma_2 AND every y'^T a'_c from the expansion are noise.  Not
            # fixable by recomputing sigma_2 alone (backward error 1.7e4 n eps): the pair falls back to two single
            # steps, i.e. one more exchange with the true y'.
            fallbacks += 1
            for jj in (j, j + 1):
                x = A[jj + 1:, jj].copy()
                beta, iu, tt = scalars(A[jj, jj], F(x @ x))
                tau[jj] = tt
                if tt != 0:
                    for c in range(jj + 1, n):
                        d = F(A[jj, c] + F(x @ A[jj + 1:, c]) * iu)
                        w = F(tt * d)
                        A[jj, c] -= w
                        A[jj + 1:, c] -= F(w * iu) * x
                    A[jj, jj] = beta
                    A[jj + 1:, jj] = x * iu
            continue
        sig2 = max(sig2, F(0))
        beta2, iu2, t2 = scalars(alpha2, sig2)
        tau[j], tau[j + 1] = t1, t2
        # ---- every column right of the pair
        for c in range(j + 2, n):
            pc = F(p[c] + xj1 * r2[c])                                 # sums including row j+1
            d1 = F(r1[c] + pc * iu1)
            tc = F(t1 * d1)
            ec = F(tc * iu1)
            r2c = F(r2[c] - ec * xj1)                                  # a'(j+1, c)
            inner = F(q[c] - ec * p[j + 1] - a1 * p[c] + a1 * ec * p[j])   # y'^T a'_c below row j+1
            d2 = F(r2c + inner * iu2)
            s2 = F(t2 * d2)
            fc = F(s2 * iu2)
            A[j, c] = r1[c] - tc
            A[j + 1, c] = r2c - s2
            A[j + 2:, c] -= F(ec - fc * a1) * xh + fc * yh             # two FMAs per element, as two single steps
        # ---- the pair's own columns: R entries, reflectors
        A[j, j] = beta1 if t1 != 0 else r1[j]
        A[j, j + 1] = r1[j + 1] - t
        A[j + 1, j + 1] = beta2 if t2 != 0 else alpha2
        A[j + 1, j] = xj1 * iu1
        A[j + 2:, j] = xh * iu1
        A[j + 2:, j + 1] = (yh - a1 * xh) * iu2
    return A, tau, fallbacks


def q_from(V, tau):
    m, n = V.shape
    Q = np.eye(m)
    for j in reversed(range(n)):
        v = np.zeros(m); v[j] = 1.0; v[j + 1:] = V[j + 1:, j]
        Q -= float(tau[j]) * np.outer(v, v @ Q)
    return Q[:, :n]


def report(name, A, **kw):
    S, ts = sweep_single(A)
    Pp, tp, fb = sweep_pairs(A, **kw)
    n = A.shape[1]
    eps = 2.0 ** -23
    out = []
    for tag, (V, tau) in (("single", (S, ts)), ("pairs", (Pp, tp))):
        Q = q_from(V.astype(np.float64), tau)
        R = np.triu(V[:n].astype(np.float64))
        be = np.linalg.norm(A - Q @ R) / (np.linalg.norm(A) * n * eps)
        orth = np.linalg.norm(Q.T @ Q - np.eye(n)) / (n * eps)
        out.append(f"{tag}: backward {be:7.3f} orth {orth:9.3f}")
    dR = np.linalg.norm(np.triu(S[:n]) - np.triu(Pp[:n])) / np.linalg.norm(np.triu(S[:n]))
    print(f"{name:34s} {out[0]}   {out[1]}   |R_pairs - R_single|/|R| {dR:.1e}  guard fallbacks {fb}")


def guard_study(guard, trials=4):
    """Worst case for a given guard: EVERY odd column nearly parallel to its left neighbour, sin^2(angle) = ratio, so that
    sigma_2 / q_{j+1} ~ ratio in all 32 pairs.  Prints the acceptance numbers of the pair sweep and of the single sweep."""
    rng = np.random.default_rng(1)
    eps = 2.0 ** -23
    print(f"guard = {guard}")
    for ratio in (0.5, 0.1, 3e-2, 1e-2, 3e-3, 1.5e-3):
        worst = [0.0, 0.0, 0.0, 0.0]
        fbs = 0
        for _ in range(trials):
            A = rng.standard_normal((512, 64)).astype(F)
            for j in range(0, 64, 2):
                A[:, j + 1] = (np.sqrt(1 - ratio) * A[:, j] + np.sqrt(ratio) * A[:, j + 1]).astype(F)
            Pp, tp, fb = sweep_pairs(A, guard=guard)
            S, ts = sweep_single(A)
            fbs += fb
            for k, (V, tau) in enumerate(((Pp, tp), (S, ts))):
                Q = q_from(V.astype(np.float64), tau)
                R = np.triu(V[:64].astype(np.float64))
                worst[2 * k] = max(worst[2 * k], np.linalg.norm(A - Q @ R) / (np.linalg.norm(A) * 64 * eps))
                worst[2 * k + 1] = max(worst[2 * k + 1], np.linalg.norm(Q.T @ Q - np.eye(64)) / (64 * eps))
        print(f"  sin^2 = {ratio:7.1e}: pairs backward {worst[0]:7.3f} orth {worst[1]:8.3f}   single backward {worst[2]:6.3f} "
              f"orth {worst[3]:6.3f}   fallbacks {fbs}")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "guard_study":
        for g in (1e-3, 0.05, 0.1):
            guard_study(g)
        sys.exit(0)
    rng = np.random.default_rng(3)
    A = rng.random((512, 64)).astype(F)
    report("uniform[0,1) 512x64", A)
    report("N(0,1) graded 1e-6..1e6", (rng.standard_normal((512, 64)) * np.logspace(-6, 6, 64)).astype(F))
    B = rng.standard_normal((512, 64)).astype(F)
    B[:, 11] = B[:, 10] * F(1.0 + 1e-6) ; B[:, 21] = B[:, 20]
    report("neighbouring columns dependent", B)
    report("  ... same, guard OFF", B, use_guard=False)
