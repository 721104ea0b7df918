"""fp32 numpy spec of the blocked flat-tree TSQR leaf (cuda-qr_b200/csrc/tsqr_mma.cu).

The running R (64 x 64) absorbs one 64-row block B at a time -- the reference's flat tree (qr.c:68-73,109-141) with the
structured reflector v_j = [e_j; x_j / u_j] -- but in sub-panels of 8 columns: the sub-panel's 8 reflectors are generated
column by column (scalar formulas of qr.c:144-152), their compact-WY T comes from T^-1 = diag(1/tau) + striu(X^T X), and
the columns to the right get  W = R(J, C) + X^T B(:, C),  Y = T^T W,  R(J, C) -= Y,  B(:, C) -= X Y  as 3xTF32
tensor-core products (emulated here: operands split into a TF32 head and an exact fp32 remainder, three products,
fp32 accumulation).   python tools/mma_leaf_spec.py
"""
import numpy as np

f32 = np.float32


def tf32_rna(x):
    """cvt.rna.tf32.f32: round to nearest (ties away) on the 13 dropped bits."""
    u = np.asarray(x, dtype=f32).view(np.uint32).astype(np.uint64)
    u = (u + 0x1000) & 0xFFFFE000
    return u.astype(np.uint32).view(f32)


def tf32_trunc(x):
    return (np.asarray(x, dtype=f32).view(np.uint32) & np.uint32(0xFFFFE000)).view(f32)


def mm3(a, b):
    """3xTF32 product a @ b: (a_lo b_hi + a_hi b_lo) + a_hi b_hi; the hardware truncates the lo operands to TF32."""
    a = np.asarray(a, dtype=f32); b = np.asarray(b, dtype=f32)
    ah, bh = tf32_rna(a), tf32_rna(b)
    al, bl = tf32_trunc(a - ah), tf32_trunc(b - bh)
    acc = (al.astype(np.float64) @ bh.astype(np.float64)).astype(f32)
    acc = (acc.astype(np.float64) + ah.astype(np.float64) @ bl.astype(np.float64)).astype(f32)
    return (acc.astype(np.float64) + ah.astype(np.float64) @ bh.astype(np.float64)).astype(f32)


def block_step(R, B, n=64, w=8, exact_products=False):
    """One 64-row block into the running R (in place, both fp32)."""
    mm = (lambda a, b: (a.astype(np.float64) @ b.astype(np.float64)).astype(f32)) if exact_products else mm3
    for p0 in range(0, n, w):
        J = range(p0, min(p0 + w, n))
        wj = len(J)
        G = np.zeros((wj, wj), dtype=f32)
        tau = np.zeros(wj, dtype=f32)
        for jj, j in enumerate(J):
            x = B[:, j].copy()
            sig = f32(x @ x)
            alpha = R[j, j]
            sj = f32(alpha * alpha + sig)
            if sig == 0 or sj < 1.2e-38:
                inv_u = f32(0); t = f32(0); beta = alpha
            else:
                nrm = f32(np.sqrt(sj))
                beta = nrm if alpha < 0 else -nrm
                u = f32(alpha - beta)
                inv_u = f32(1) / u
                t = f32(-u / beta)
            for i in range(jj):                              # finished columns of the sub-panel hold x~_i
                G[i, jj] = f32(f32(B[:, J[i]] @ x) * inv_u)
            for c in J[jj + 1:]:
                d = f32(x @ B[:, c])
                s = f32(R[j, c] + d * inv_u)
                R[j, c] = f32(R[j, c] - t * s)
                B[:, c] = (B[:, c] - f32(t * s * inv_u) * x).astype(f32)
            R[j, j] = beta
            B[:, j] = (x * inv_u).astype(f32)
            tau[jj] = t
        # T by back substitution on T^-1 = diag(1/tau) + striu(G), column by column
        T = np.zeros((wj, wj), dtype=f32)
        for k in range(wj):
            T[k, k] = tau[k]
            for i in range(k - 1, -1, -1):
                acc = f32(0)
                for l in range(i + 1, k + 1):
                    acc = f32(acc + G[i, l] * T[l, k])
                T[i, k] = f32(-tau[i] * acc)
        C = list(range(p0 + wj, n))
        if not C:
            continue
        X = B[:, list(J)]
        W = (R[np.ix_(list(J), C)] + mm(X.T, B[:, C])).astype(f32)
        Y = mm(T.T, W)
        R[np.ix_(list(J), C)] = (R[np.ix_(list(J), C)] - Y).astype(f32)
        B[:, C] = (B[:, C] - mm(X, Y)).astype(f32)


def flat_tsqr_r(A, **kw):
    m, n = A.shape
    R = np.zeros((64, 64), dtype=f32)
    for r0 in range(0, m, 64):
        B = np.zeros((64, 64), dtype=f32)
        blk = A[r0:r0 + 64]
        B[:blk.shape[0], :n] = blk
        block_step(R, B, n=n, **kw)
    return np.triu(R[:n, :n])


def r_rel_diff(Ra, Rb):
    sa = np.where(np.diag(Ra) < 0, -1.0, 1.0); sb = np.where(np.diag(Rb) < 0, -1.0, 1.0)
    a = Ra.astype(np.float64) * sa[:, None]; b = Rb.astype(np.float64) * sb[:, None]
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


if __name__ == "__main__":
    rng = np.random.default_rng(3)
    for name, A in [("uniform 4096x64", rng.random((4096, 64), dtype=f32)),
                    ("normal 4096x64", rng.standard_normal((4096, 64)).astype(f32)),
                    ("graded 2048x64", (rng.standard_normal((2048, 64)) * np.logspace(0, -6, 64)).astype(f32)),
                    ("ragged 1000x40", rng.random((1000, 40), dtype=f32)),
                    ("dup col 2048x64", None)]:
        if A is None:
            A = rng.standard_normal((2048, 64)).astype(f32); A[:, 17] = A[:, 5]; A[:, 40] = 0
        R64 = np.linalg.qr(A.astype(np.float64), mode="r")
        G = A.astype(np.float64).T @ A.astype(np.float64)
        for label, kw in [("3xTF32", {}), ("exact products", {"exact_products": True})]:
            R = flat_tsqr_r(A, **kw)
            Rd = R.astype(np.float64)
            gram = np.linalg.norm(Rd.T @ Rd - G) / np.linalg.norm(G)
            print(f"{name:18s} {label:15s} gram {gram:.2e}  |R-R64|/|R64| {r_rel_diff(R, R64[:A.shape[1]]):.2e}")
