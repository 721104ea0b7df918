"""fp32 numpy specification of the slicing the Gram leaf (cuda-qr_b200/csrc/gram_umma.cu) applies to every 128-row group of a
column before the tensor-core contraction, and of the properties its exactness argument rests on:
   a = s1 + s2 + s3 + rho,  s1, s2 multiples of 2^(E-8), 2^(E-16) with |s1| <= 256, |s2| <= 128 quanta (2^E > max |a| of the group),
   s3 = bf16(r2) (8 significant bits, |r2| <= 2^(E-17)),  |rho| <= 2^(E-25),
   every slice exactly representable in bf16, and a group's S1^T S1 / S1^T S2 / S2^T S2 sums integers below 2^24 of one quantum,
   i.e. exact in an fp32 accumulator whatever its rounding mode.
Run by tests/test_gram_slicing_cpu.py;  python tools/gram_slicing_spec.py prints a small report."""
import numpy as np

F = np.float32


def bf16_round(x):
    """round-to-nearest-even fp32 -> bf16 -> fp32 (finite inputs)"""
    u = np.asarray(x, dtype=F).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(F)


def is_bf16(x):
    return np.all((np.asarray(x, dtype=F).view(np.uint32) & 0xFFFF) == 0)


def slice_group(v):
    """v: the 128 fp32 values of one column in one group.  Returns (s1, s2, s3, E) exactly as the kernel forms them."""
    v = np.asarray(v, dtype=F)
    mx = np.max(np.abs(v)).astype(F)
    eb = int(mx.view(np.uint32) >> 23)
    if eb == 0:                                   # zero (or denormal) column: the slices are the values themselves
        z = np.zeros_like(v)
        return v.copy(), z, bf16_round(z), None
    sg = lambda e: np.uint32((e << 23) | 0x400000).view(F)       # 1.5 * 2^(e - 127)
    sg1, sg2 = sg(eb + 16), sg(eb + 8)
    a1 = ((v + sg1).astype(F) - sg1).astype(F)
    r1 = (v - a1).astype(F)
    a2 = ((r1 + sg2).astype(F) - sg2).astype(F)
    r2 = (r1 - a2).astype(F)
    return a1, a2, bf16_round(r2), eb + 1 - 127


def check_group(v):
    s1, s2, s3, E = slice_group(v)
    assert is_bf16(s1) and is_bf16(s2) and is_bf16(s3)
    if E is None:
        return
    v64 = np.asarray(v, dtype=np.float64)
    u1 = s1.astype(np.float64) / 2.0 ** (E - 8)
    u2 = s2.astype(np.float64) / 2.0 ** (E - 16)
    assert np.all(u1 == np.round(u1)) and np.abs(u1).max() <= 256
    assert np.all(u2 == np.round(u2)) and np.abs(u2).max() <= 128
    r2 = v64 - s1 - s2                                           # exact in fp64
    assert np.abs(r2).max() <= 2.0 ** (E - 17)
    rho = r2 - s3
    assert np.abs(rho).max() <= 2.0 ** (E - 25)
    # exact accumulation: |sum of 128 products| in quanta stays below 2^24 for the three fixed-point blocks
    assert np.abs(u1).astype(np.int64) @ np.abs(u1).astype(np.int64) < 2 ** 24
    assert np.abs(u1).astype(np.int64) @ np.abs(u2).astype(np.int64) < 2 ** 24
    assert np.abs(u2).astype(np.int64) @ np.abs(u2).astype(np.int64) < 2 ** 24
    return float(np.abs(rho).max() / 2.0 ** E)


def gram_of_slices(A):
    """Gram matrix of an (m x n) fp32 matrix the way the kernel forms it (groups of 128 rows, per-column scales, D33 dropped),
    with every sum carried in fp64 -- legitimate because each fixed-point block is exact in fp32 anyway."""
    A = np.asarray(A, dtype=F)
    m, n = A.shape
    G = np.zeros((n, n))
    for g0 in range(0, m, 128):
        blk = A[g0:g0 + 128]
        if blk.shape[0] < 128:
            blk = np.vstack([blk, np.zeros((128 - blk.shape[0], n), dtype=F)])
        S = [np.zeros((128, n)), np.zeros((128, n)), np.zeros((128, n))]
        for c in range(n):
            s1, s2, s3, _ = slice_group(blk[:, c])
            S[0][:, c], S[1][:, c], S[2][:, c] = s1, s2, s3
        U = 0.5 * S[0].T @ S[0] + S[0].T @ S[1] + S[0].T @ S[2]
        V = 0.5 * S[1].T @ S[1] + S[1].T @ S[2]
        G += U + U.T + V + V.T
    return G


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    worst = 0.0
    for t in range(300):
        scale = 2.0 ** rng.integers(-30, 30)
        v = (rng.random(128) if t % 3 == 0 else rng.standard_normal(128) if t % 3 == 1 else
             rng.standard_normal(128) * 10.0 ** rng.uniform(-6, 0, 128)).astype(F) * F(scale)
        worst = max(worst, check_group(v))
    print(f"300 random groups: every slice bf16-exact, integer bounds hold, max |rho| / 2^E = {worst:.3e} (bound 2^-25 = {2.0 ** -25:.3e})")
    for kind in ("uniform", "normal"):
        A = (rng.random((1024, 16)) if kind == "uniform" else rng.standard_normal((1024, 16))).astype(F)
        G, Gx = gram_of_slices(A), A.astype(np.float64).T @ A.astype(np.float64)
        print(f"{kind:8s} 1024 x 16: |G - Gx|_F / |Gx|_F = {np.linalg.norm(G - Gx) / np.linalg.norm(Gx):.2e}")
