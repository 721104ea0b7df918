#!/usr/bin/env python
"""Regenerates tests/golden/* from the UNMODIFIED reference (oracle/_ref, built from
/root/reference/qr.c by oracle/build_ref.sh).  Run in the build container only:

    python tests/golden/make_golden.py

Inputs are the reference's own recipe (srand(12); A[i] = (float)rand()/RAND_MAX,
qr.c:468-474) and are NOT stored: oracle.rand_matrix() regenerates them bit for bit.
Outputs stored: the factored storage RV, tau, and for the small cases Q and R.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# (PR, PC, m, n, store_QR)
CASES = [
    (4, 2, 6, 4, True),        # qr.c main(): the reference's only known-answer run
    (4, 2, 64, 32, True),
    (64, 4, 64, 64, True),     # single window (batched 64x64 config's shape)
    (64, 4, 124, 64, True),
    (64, 4, 244, 124, False),
    (64, 8, 512, 512, False),  # BASELINE config 1 at its legal PR=64 geometry
]


def main():
    out = {}
    for PR, PC, m, n, store_qr in CASES:
        ref = oracle.Ref(PR, PC)
        A = oracle.rand_matrix(m, n, 12)
        rv, tau = ref.mmqr(A)
        key = f"pr{PR}_pc{PC}_{m}x{n}"
        if m * n <= 64 * 64 * 2:
            out[key + "_rv"] = rv
        out[key + "_tau"] = tau
        out[key + "_Rpacked"] = np.triu(rv[:n, :])[np.triu_indices(n)].astype(np.float32)
        if store_qr:
            Q, R = ref.explicitQR(rv, tau)
            out[key + "_Q"] = Q
        print(key, "done")
    np.savez_compressed(os.path.join(HERE, "ref_cases.npz"), **out)
    # The 6x4 demo also as human-readable JSON (values as printed by qr.c main).
    ref = oracle.Ref(4, 2)
    A = oracle.rand_matrix(6, 4, 12)
    rv, tau = ref.mmqr(A)
    Q, R = ref.explicitQR(rv, tau)
    resid = float(np.sqrt(np.sum(((ref.dgemm(Q, R) - A).astype(np.float32)) ** 2, dtype=np.float32)))
    with open(os.path.join(HERE, "demo_6x4.json"), "w") as f:
        json.dump({"source": "qr.c:461-523 (unmodified reference, srand(12), PR=4, PC=2)",
                   "A_rowmajor_rows": A.tolist(), "RV_rows": rv.tolist(), "tau": tau.tolist(),
                   "Q_rows": Q.tolist(), "R_rows": R.tolist(), "residual_fro": resid,
                   "survey_printed_residual": 3.78809091e-07,
                   "first_rand_values": [1687063760, 945274514, 247215794]}, f, indent=1)


if __name__ == "__main__":
    main()
