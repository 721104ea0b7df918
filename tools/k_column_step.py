"""Numerical spec of a k-pivot-columns-per-exchange Householder step (generalisation of tools/two_column_step.py; study for
the next panel kernel: k = 4 would leave 16 cluster exchanges per 64-column panel instead of 32).

One exchange delivers, for the group J = {j .. j+k-1}:  S = X^T A (k x n) with X = A[j+k:, J] (the group's columns BELOW
the group's pivot rows, before any of the group's reflectors) and the k pivot rows Rw = A[j:j+k, :].  The below part of
every column stays implicit:  a_c(below) = a_c(orig) - X D(:, c),  and every reflector's below part is a combination of
the x's:  w_r = X C(:, r).  With G = X^T X = S[:, J] all inner products follow from k x k algebra:
    w_r^T a_c(below) = C(:, r)^T (S(:, c) - G D(:, c)).
The pivot rows are updated explicitly.  At the end every column gets  a_c(below) -= X D(:, c)  (k FMAs per element, the
work of k single steps) and the group's own columns become  X (e_r - D(:, r)) / u_r.
Cancellation: sigma_r = m^T G m with m = e_r - D(:, r) loses everything when column r is nearly in the span of the
group's earlier columns; a group whose sigma_r < guard * G(r, r) stops after r columns (they are finished from the data
at hand) and the next exchange starts at column j + r with the rest of the group (groups stay aligned to multiples of k).
    python tools/k_column_step.py            # k = 1, 2, 4, 8 on the cases of two_column_step.py + worst-case study"""
import sys

import numpy as np

F = np.float32


def scalars(alpha, sigma):
    sj = F(alpha * alpha + sigma)
    if sj < F(1.2e-38):
        return alpha, F(0), F(0)
    nrm = F(np.sqrt(sj))
    beta = -alpha if sigma == 0 else (nrm if alpha < 0 else -nrm)
    u = F(alpha - beta)
    return beta, F(1) / u, (F(2) if sigma == 0 else F(-u / beta))


def sweep_k(A, k, guard=0.1):
    """k columns per exchange, fp32 throughout; returns the factored panel (LAPACK storage), tau, #exchanges"""
    A = A.astype(F).copy()
    m, n = A.shape
    tau = np.zeros(n, F)
    j = 0
    nex = 0
    while j < n:
        kk = min(k - (j % k), n - j)                         # groups stay aligned to multiples of k (register layout of a kernel)
        nex += 1
        X = A[j + kk:, j:j + kk].copy()                      # below the group's pivot rows
        S = (X.T @ A[j + kk:, :]).astype(F)                  # what the exchange delivers (kk x n)
        Rw = A[j:j + kk, :].copy()                           # pivot rows (kk x n), updated explicitly
        G = S[:, j:j + kk].copy()
        D = np.zeros((kk, n), F)                             # a_c(below) = a_c(orig) - X D(:, c)
        C = np.zeros((kk, kk), F)                            # w_r = X C(:, r)
        done = 0
        for r in range(kk):
            mvec = -D[:, j + r].copy(); mvec[r] += F(1)      # current column j+r below = X mvec
            sig_below = F(mvec @ (G @ mvec).astype(F))
            if r > 0 and sig_below < F(guard) * G[r, r]:
                break                                        # group stops here: r columns done
            sig_below = max(sig_below, F(0))
            sub = Rw[r + 1:, j + r].copy()                   # explicit entries between the pivot and the below part
            beta, iu, t = scalars(Rw[r, j + r], F(sig_below + F(sub @ sub)))
            tau[j + r] = t
            vexp = (sub * iu).astype(F)
            C[:, r] = mvec * iu
            if t != 0:
                Z = (S - (G @ D).astype(F)).astype(F)        # X^T a_c(below, current) for every column
                dots = (Rw[r, :] + (vexp @ Rw[r + 1:, :]).astype(F) + (C[:, r] @ Z).astype(F)).astype(F)
                tc = (t * dots).astype(F)
                cols = np.arange(n) > j + r                  # only the columns right of the pivot column are updated
                Rw[r, cols] -= tc[cols]
                Rw[r + 1:, cols] -= np.outer(vexp, tc[cols]).astype(F)
                D[:, cols] += np.outer(C[:, r], tc[cols]).astype(F)
                Rw[r, j + r] = beta
                Rw[r + 1:, j + r] = vexp
            done = r + 1
        # ---- write back: pivot rows, implicit update of every column right of the finished ones, the reflectors
        A[j:j + kk, :] = Rw
        right = np.arange(n) >= j + done
        A[j + kk:, right] -= (X @ D[:, right]).astype(F)
        for r in range(done):                                # tau = 0 (zero column): C(:, r) = 0 and the column is ~0 anyway
            if tau[j + r] != 0:
                A[j + kk:, j + r] = (X @ C[:, r]).astype(F)
        if done < kk:
            # rows j+done .. j+kk-1 were treated as pivot rows of the group but belong to the below part of the next group:
            # nothing to fix, they were updated explicitly like any other row
            pass
        j += done
    return A, tau, nex


def sweep_single(A):
    return sweep_k(A, 1)[:2]


def q_from(V, tau):
    m, n = V.shape
    Q = np.eye(m)
    for j in reversed(range(n)):
        v = np.zeros(m); v[j] = 1.0; v[j + 1:] = V[j + 1:, j]
        Q -= float(tau[j]) * np.outer(v, v @ Q)
    return Q[:, :n]


def metrics(A, V, tau):
    n = A.shape[1]
    eps = 2.0 ** -23
    Q = q_from(V.astype(np.float64), tau)
    R = np.triu(V[:n].astype(np.float64))
    return np.linalg.norm(A - Q @ R) / (np.linalg.norm(A) * n * eps), np.linalg.norm(Q.T @ Q - np.eye(n)) / (n * eps)


def report(name, A, guard=0.1):
    out = []
    for k in (1, 2, 4, 8):
        V, tau, nex = sweep_k(A, k, guard)
        be, orth = metrics(A, V, tau)
        out.append(f"k={k}: {nex:2d} exch backward {be:6.3f} orth {orth:7.3f}")
    print(f"{name:34s} " + "   ".join(out))


def worst_case(k, guard, trials=4):
    """adversary: inside every group column r is nearly parallel to column r-1, sin^2(angle) = rho, so that after the
    earlier reflectors of the group EVERY column sits at sigma / G(r, r) ~ rho -- just above the guard in the worst rows"""
    rng = np.random.default_rng(4)
    print(f"k = {k}, guard = {guard}")
    for rho in (0.11, 0.15, 0.26, 0.5, 1e-2):
        wb = wo = 0.0
        ex = 0
        for _ in range(trials):
            A = rng.standard_normal((512, 64)).astype(F)
            for j in range(0, 64, k):
                for r in range(1, k):
                    A[:, j + r] = (np.sqrt(1 - rho) * A[:, j + r - 1] + np.sqrt(rho) * A[:, j + r]).astype(F)
            V, tau, nex = sweep_k(A, k, guard)
            be, orth = metrics(A, V, tau)
            wb = max(wb, be); wo = max(wo, orth); ex += nex
        print(f"  chain rho = {rho:5.2f}: backward {wb:7.3f} orth {wo:8.3f}   exchanges per panel {ex / trials:5.1f}")


if __name__ == "__main__":
    rng = np.random.default_rng(3)
    report("uniform[0,1) 512x64", rng.random((512, 64)).astype(F))
    report("uniform[0,1) 4096x64", rng.random((4096, 64)).astype(F))
    report("N(0,1) 512x64", rng.standard_normal((512, 64)).astype(F))
    report("N(0,1) graded 1e-6..1e6", (rng.standard_normal((512, 64)) * np.logspace(-6, 6, 64)).astype(F))
    B = rng.standard_normal((512, 64)).astype(F)
    B[:, 11] = B[:, 10] * F(1.0 + 1e-6); B[:, 21] = B[:, 20]; B[:, 40] = 0
    report("dependent neighbours + zero column", B)
    if len(sys.argv) > 1 and sys.argv[1] == "worst":
        for k in (2, 4, 8):
            for g in (0.1, 0.25):
                worst_case(k, g)
