"""Lane-level numpy emulation of panel_wb2.cu (two pivot columns per cluster exchange), written to check the kernel's
register layout, masks, exchange index maps and algebra on the CPU: every warp is a [32 lanes][8 slots][8 pairs][2]
array exactly as in the kernel (lane = 8 h + q owns columns q + 8 i and rows 8 k + 2 h + {0,1} of its 64-row block),
the cluster all-reduce is replayed message by message (owner / slot / peer maps of the kernel), and the result (LAPACK
storage, tau, T) is compared with the plain column-by-column sweep of tools/two_column_step.py.
    python tools/emulate_pair_panel.py            # a few shapes, pairs / forced fallback / dependent columns
    python tools/emulate_pair_panel.py fuzz [seed] # random plans, heights, data kinds and modes"""
import os, sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from two_column_step import sweep_single, q_from  # noqa: E402

F = np.float32


def scalars(alpha, sig):
    sj = F(alpha * alpha + sig)
    ok = sj >= F(1.2e-38)
    sjs = sj if ok else F(1)
    nrm = F(np.sqrt(sjs))
    len1 = sig == 0
    bc = -alpha if len1 else (nrm if alpha < 0 else -nrm)
    u = F(alpha - bc)
    return bc, (F(1) / u if ok else F(0)), ((F(2) if len1 else F(-u / bc)) if ok else F(0)), ok


def run(A, W, CS, mode=1, NCL=1):
    mp, nb = A.shape
    assert nb == 64
    nwt = W * CS * NCL
    assert mp <= 64 * nwt
    Ap = np.zeros((64 * nwt, 64), F); Ap[:mp] = A
    lanes = np.arange(32); qv = lanes & 7; hv = lanes >> 3
    # b[gw][lane][i][k][e]
    b = np.zeros((nwt, 32, 8, 8, 2), F)
    for gw in range(nwt):
        for lane in range(32):
            q, h = qv[lane], hv[lane]
            for i in range(8):
                for k in range(8):
                    r0 = 64 * gw + 8 * k + 2 * h
                    b[gw, lane, i, k, :] = Ap[r0:r0 + 2, q + 8 * i]
    tau = np.zeros(64, F); gs = np.zeros((64, 64), F)
    nex = 0
    fallbacks = 0
    for I0 in range(8):
        jj = 0; second = False
        while jj < 8:
            j = 8 * I0 + jj
            nex += 1
            hj = jj >> 1
            qx = jj + 1 if second else jj; qy = jj + 1
            xs = np.zeros((nwt, 64), F); ys = np.zeros((nwt, 64), F); prl = np.zeros(128, F)
            for gw in range(nwt):
                top = gw == 0
                for lane in range(32):
                    q, h = qv[lane], hv[lane]
                    if q == qx or q == qy:
                        for k in range(8):
                            v = b[gw, lane, I0, k].copy()
                            if top and (k < I0 or (k == I0 and h <= hj)): v[:] = 0
                            if q == qx: xs[gw, 8 * k + 2 * h: 8 * k + 2 * h + 2] = v
                            if q == qy: ys[gw, 8 * k + 2 * h: 8 * k + 2 * h + 2] = v
                    if top and h == hj:
                        for i in range(8):
                            prl[q + 8 * i] = b[gw, lane, i, I0, 0]
                            prl[64 + q + 8 * i] = b[gw, lane, i, I0, 1]
            # per-lane x, y registers and the warp sums (halved first shuffle stage)
            X = np.zeros((nwt, 32, 8, 2), F); Y = np.zeros((nwt, 32, 8, 2), F)
            part = np.zeros((nwt, 128), F)
            for gw in range(nwt):
                for lane in range(32):
                    h = hv[lane]
                    for k in range(8):
                        X[gw, lane, k] = xs[gw, 8 * k + 2 * h: 8 * k + 2 * h + 2]
                        Y[gw, lane, k] = ys[gw, 8 * k + 2 * h: 8 * k + 2 * h + 2]
                dx = np.einsum('lke,like->li', X[gw], b[gw]).astype(F)
                dy = np.einsum('lke,like->li', Y[gw], b[gw]).astype(F)
                keep = np.zeros((32, 8), F)
                for lane in range(32):
                    odd = hv[lane] & 1
                    partner = lane ^ 8
                    give_p = dx[partner] if (hv[partner] & 1) else dy[partner]
                    keep[lane] = (dy[lane] if odd else dx[lane]) + give_p
                keep2 = keep + keep[lanes ^ 16]
                for lane in range(32):
                    q, h = qv[lane], hv[lane]
                    if h < 2:
                        for i in range(8): part[gw, 64 * h + q + 8 * i] = keep2[lane, i]
            # CTA sums and the cluster exchange, message by message (per cluster), then the two-cluster sum
            cl_tot = []; cl_prow = []
            for cl in range(NCL):
                tot_in = np.full((CS, 128), np.nan, F); prow = np.full((CS, 128), np.nan, F)
                sv = np.zeros((CS, 32, 4), F)
                for r in range(CS):
                    for lane in range(32):
                        for ww in range(W): sv[r, lane] += part[(cl * CS + r) * W + ww, 4 * lane: 4 * lane + 4]
                if CS == 1:
                    tot_in[0] = sv[0].reshape(128); prow[0] = prl
                else:
                    wpo = 32 // CS
                    rs_in = np.full((CS, 32, 4), np.nan, F); prs_in = np.full((CS, 32, 4), np.nan, F)
                    for r in range(CS):
                        for lane in range(32):
                            owner = (lane * CS) >> 5; wl = lane - owner * wpo
                            rs_in[owner, r * wpo + wl] = sv[r, lane]
                            if r == 0 and cl == 0: prs_in[owner, wl] = prl[4 * lane: 4 * lane + 4]
                    for r in range(CS):      # owner r
                        t = rs_in[r].copy()
                        assert not np.isnan(t).any()
                        o = 16
                        while o >= wpo:
                            t = t + t[lanes ^ o]; o >>= 1
                        for lane in range(32):
                            slot, peer = lane % wpo, lane // wpo
                            col4 = 4 * (r * wpo + slot)
                            tot_in[peer, col4: col4 + 4] = t[lane]
                            if cl == 0: prow[peer, col4: col4 + 4] = prs_in[r, slot]
                assert not np.isnan(tot_in).any() and (cl > 0 or not np.isnan(prow).any())
                for r in range(1, CS):
                    assert (tot_in[r] == tot_in[0]).all() and (cl > 0 or (prow[r] == prow[0]).all())
                cl_tot.append(tot_in[0]); cl_prow.append(prow[0])
            tot_in = (cl_tot[0] + cl_tot[1])[None] if NCL == 2 else cl_tot[0][None]
            prow = cl_prow[0][None]
            P, Q, R1, R2 = tot_in[0, :64], tot_in[0, 64:], prow[0, :64], prow[0, 64:]

            def colmask(top, k, h):   # registers that hold rows >= j+2
                return (not top) or k > I0 or (k == I0 and h > hj)

            if not second:
                xj1, yj1 = R2[j], R2[j + 1]
                bc1, iu1, t1, ok1 = scalars(R1[j], F(xj1 * xj1 + P[j]))
                d1n = F(F(xj1 * yj1 + P[j + 1]) * iu1 + R1[j + 1])
                tn = F(t1 * d1n); a1 = F(tn * iu1); alpha2 = F(-a1 * xj1 + yj1)
                sig2 = F(F(a1 * a1) * P[j] + F(F(-2 * a1) * P[j + 1] + Q[j + 1]))
                fb = mode == 2 or sig2 < F(0.1) * Q[j + 1]
                tau[j] = t1
                if not fb:
                    bc2, iu2, t2, ok2 = scalars(alpha2, max(sig2, F(0)))
                    tau[j + 1] = t2
                for gw in range(nwt):
                    top = gw == 0
                    for lane in range(32):
                        q, h = qv[lane], hv[lane]
                        for i in range(8):
                            c = q + 8 * i
                            d1 = F(F(xj1 * R2[c] + P[c]) * iu1 + R1[c])
                            if i <= I0 and top and h == 0:
                                if c < j:
                                    gs[c, j] = d1
                                    if not fb: gs[c, j + 1] = F(F(-a1 * P[c] + Q[c]) * iu2 + R2[c])
                                elif c == j and not fb:
                                    gs[j, j + 1] = F(F(iu1 * iu2) * F(-a1 * P[j] + P[j + 1]) + F(xj1 * iu1))
                            if i >= I0:
                                tc = F(t1 * d1); ec = F(tc * iu1)
                                if not fb:
                                    act = i > I0 or q > jj + 1
                                    r2c = F(-ec * xj1 + R2[c])
                                    inner = F(F(a1 * ec) * P[j] + F(-a1 * P[c] + F(-ec * P[j + 1] + Q[c])))
                                    sc2 = F(t2 * F(inner * iu2 + r2c)); fc = F(sc2 * iu2)
                                    cxv = F(fc * a1 - ec) if act else F(0); cyv = -fc if act else F(0)
                                    b[gw, lane, i] = cxv * X[gw, lane] + (cyv * Y[gw, lane] + b[gw, lane, i])
                                    if act and top and h == hj: b[gw, lane, i, I0] = (R1[c] - tc, r2c - sc2)
                                else:
                                    act = i > I0 or q > jj
                                    cxv = -ec if act else F(0)
                                    b[gw, lane, i] = cxv * X[gw, lane] + b[gw, lane, i]
                                    if act and top and h == hj: b[gw, lane, i, I0] = (R1[c] - tc, F(-ec * xj1 + R2[c]))
                        if not fb and q == jj + 1:
                            m2 = iu2 if ok2 else F(1)
                            for k in range(8):
                                nv = (-a1 * X[gw, lane, k] + b[gw, lane, I0, k]) * m2
                                if colmask(top, k, h): b[gw, lane, I0, k] = nv
                            if top and h == hj: b[gw, lane, I0, I0] = (R1[j + 1] - tn, bc2 if ok2 else alpha2)
                        if q == jj and ok1:
                            for k in range(8):
                                if colmask(top, k, h): b[gw, lane, I0, k] *= iu1
                            if top and h == hj: b[gw, lane, I0, I0] = (bc1, xj1 * iu1)
                if fb: second = True; fallbacks += 1
                else: jj += 2
            else:
                bc2, iu2, t2, ok2 = scalars(R2[j + 1], P[j + 1])
                tau[j + 1] = t2
                for gw in range(nwt):
                    top = gw == 0
                    for lane in range(32):
                        q, h = qv[lane], hv[lane]
                        for i in range(8):
                            c = q + 8 * i
                            d2 = F(P[c] * iu2 + R2[c])
                            if i <= I0 and top and h == 0 and c < j + 1: gs[c, j + 1] = d2
                            if i >= I0:
                                act = i > I0 or q > jj + 1
                                sc2 = F(t2 * d2); fc = F(sc2 * iu2)
                                cxv = -fc if act else F(0)
                                b[gw, lane, i] = cxv * X[gw, lane] + b[gw, lane, i]
                                if act and top and h == hj: b[gw, lane, i, I0, 1] = R2[c] - sc2
                        if q == jj + 1 and ok2:
                            for k in range(8):
                                if colmask(top, k, h): b[gw, lane, I0, k] *= iu2
                            if top and h == hj: b[gw, lane, I0, I0, 1] = bc2
                second = False; jj += 2
    out = np.zeros_like(Ap)
    for gw in range(nwt):
        for lane in range(32):
            q, h = qv[lane], hv[lane]
            for i in range(8):
                for k in range(8):
                    r0 = 64 * gw + 8 * k + 2 * h
                    out[r0:r0 + 2, q + 8 * i] = b[gw, lane, i, k]
    # T by the kernel's back substitution
    T = np.zeros((64, 64), F)
    for c in range(64):
        T[c, c] = tau[c]
        for i in range(c - 1, -1, -1):
            T[i, c] = -tau[i] * F(gs[i, i + 1:c + 1] @ T[i + 1:c + 1, c])
    return out[:mp], tau, T, nex, fallbacks


def check(name, A, W, CS, mode=1, NCL=1):
    S, ts = sweep_single(A)
    out, tau, T, nex, fb = run(A, W, CS, mode, NCL)
    n = 64; eps = 2.0 ** -23
    Qm = q_from(out.astype(np.float64), tau)
    R = np.triu(out[:n].astype(np.float64))
    be = np.linalg.norm(A - Qm @ R) / (np.linalg.norm(A) * n * eps)
    orth = np.linalg.norm(Qm.T @ Qm - np.eye(n)) / (n * eps)
    dR = np.linalg.norm(np.triu(S[:n]) - np.triu(out[:n])) / np.linalg.norm(np.triu(S[:n]))
    dV = np.linalg.norm(np.tril(S, -1) - np.tril(out, -1)) / np.linalg.norm(np.tril(S, -1))
    # T against the definition: Q = I - V T V^T
    V = np.tril(out.astype(np.float64), -1); V[np.arange(n), np.arange(n)] = 1.0
    Qfull_thin = (np.eye(A.shape[0]) - V @ T.astype(np.float64) @ V.T)[:, :n]
    dT = np.linalg.norm(Qfull_thin - Qm) / (n * eps)
    ok = be < 10 and orth < 10 and dT < 50
    print(f"{name:44s} W={W} CS={CS:2d} NCL={NCL} mode={mode} exchanges {nex:3d} fallbacks {fb:2d}  backward {be:6.3f} orth {orth:6.3f} "
          f"|dR| {dR:.1e} |dV| {dV:.1e} T-vs-Q {dT:6.2f}  {'ok' if ok else 'FAIL'}")
    return ok


def fuzz(seed, cases=24):
    """random plans (warps per CTA, cluster size, one or two clusters, ragged heights), data kinds (uniform, normal, zero
    columns, nearly dependent neighbours, columns graded over 16 decades, zero rows) and modes"""
    rng = np.random.default_rng(seed)
    bad = 0
    for t in range(cases):
        W = int(rng.choice([1, 2, 4])); CS = int(rng.choice([1, 2, 4, 8])); NCL = 2 if (CS > 1 and rng.random() < 0.25) else 1
        rows_max = 64 * W * CS * NCL
        mp = int(rng.integers(max(64, rows_max - 63 * W), rows_max + 1))
        kind = int(rng.integers(0, 6))
        A = rng.standard_normal((mp, 64)) if kind % 2 else rng.random((mp, 64))
        if kind == 2: A[:, rng.integers(0, 64, 3)] = 0
        if kind == 3:
            for _ in range(3):
                j = int(rng.integers(0, 63)); A[:, j + 1] = A[:, j] * (1 + rng.standard_normal() * 1e-5)
        if kind == 4: A *= np.logspace(-8, 8, 64)[rng.permutation(64)]
        if kind == 5: A[rng.integers(0, mp, mp // 2)] = 0
        bad += not check(f"fuzz {t} kind {kind} {mp}x64", A.astype(F), W, CS, mode=int(rng.choice([1, 1, 1, 2])), NCL=NCL)
    return bad == 0


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "fuzz":
        sys.exit(0 if fuzz(int(sys.argv[2]) if len(sys.argv) > 2 else 0) else 1)
    rng = np.random.default_rng(5)
    good = True
    good &= check("uniform 128x64 (one CTA)", rng.random((128, 64)).astype(F), 2, 1)
    good &= check("uniform 500x64 ragged, 4 CTAs", rng.random((500, 64)).astype(F), 2, 4)
    good &= check("uniform 1024x64, 16 CTAs", rng.random((1024, 64)).astype(F), 1, 16)
    good &= check("uniform 512x64, forced fallback", rng.random((512, 64)).astype(F), 2, 4, mode=2)
    B = rng.standard_normal((512, 64)).astype(F)
    B[:, 11] = B[:, 10] * F(1.0 + 1e-6); B[:, 21] = B[:, 20]; B[:, 40] = 0
    good &= check("dependent neighbours + zero column", B, 4, 2)
    good &= check("uniform 2000x64, two clusters of 16 CTAs", rng.random((2000, 64)).astype(F), 1, 16, NCL=2)
    good &= check("graded N(0,1) 1e-6..1e6", (rng.standard_normal((256, 64)) * np.logspace(-6, 6, 64)).astype(F), 2, 2)
    sys.exit(0 if good else 1)
