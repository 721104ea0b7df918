"""oracle/metrics.py -- TEST INFRASTRUCTURE: fp64 acceptance metrics for the QR path.

Tolerances are BASELINE.json's: backward error ||A-QR||_F / (||A||_F n eps) <= 10,
orthogonality ||Q^T Q - I||_F / (n eps) <= 10, sign-normalised R within 1e-4 of the
reference's R (normwise Frobenius), eps = 2^-23 (fp32 spacing at 1).
"""
from __future__ import annotations

import numpy as np

EPS32 = 2.0 ** -23
TOL_BACKWARD = 10.0
TOL_ORTH = 10.0
TOL_R = 1e-4


def sign_normalize(R: np.ndarray, Q: np.ndarray | None = None):
    """Flip row i of R (and column i of Q) wherever R[i, i] < 0."""
    R = np.array(R, dtype=np.float64, copy=True)
    k = min(R.shape)
    s = np.where(np.diagonal(R)[:k] < 0, -1.0, 1.0)
    R[:k, :] *= s[:, None]
    if Q is None:
        return R
    Q = np.array(Q, dtype=np.float64, copy=True)
    Q[:, :k] *= s[None, :]
    return R, Q


def r_rel_diff(R_test: np.ndarray, R_ref: np.ndarray) -> float:
    n = R_test.shape[1]
    a = sign_normalize(np.triu(np.asarray(R_test, dtype=np.float64)[:n, :]))
    b = sign_normalize(np.triu(np.asarray(R_ref, dtype=np.float64)[:n, :]))
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def backward_error(A: np.ndarray, Q: np.ndarray, R: np.ndarray) -> float:
    A64 = np.asarray(A, dtype=np.float64)
    n = A64.shape[1]
    res = A64 - np.asarray(Q, dtype=np.float64) @ np.asarray(R, dtype=np.float64)
    return float(np.linalg.norm(res) / (np.linalg.norm(A64) * n * EPS32))


def orthogonality(Q: np.ndarray) -> float:
    Q64 = np.asarray(Q, dtype=np.float64)
    k = Q64.shape[1]
    return float(np.linalg.norm(Q64.T @ Q64 - np.eye(k)) / (k * EPS32))


def gram_error(A: np.ndarray, R: np.ndarray) -> float:
    """||R^T R - A^T A||_F / ||A^T A||_F: a Q-free check usable at any m."""
    A64 = np.asarray(A, dtype=np.float64)
    n = A64.shape[1]
    R64 = np.triu(np.asarray(R, dtype=np.float64)[:n, :])
    G = A64.T @ A64
    return float(np.linalg.norm(R64.T @ R64 - G) / np.linalg.norm(G))


def householder_q(V: np.ndarray, tau: np.ndarray, full: bool = True) -> np.ndarray:
    """fp64 Q = H_0 H_1 ... from LAPACK-style storage (unit-lower V below the diagonal, tau[j])."""
    m, n = V.shape
    k = m if full else n
    Q = np.eye(m, k)
    V64 = np.asarray(V, dtype=np.float64)
    for j in reversed(range(n)):
        if j >= m:
            continue
        v = np.zeros(m)
        v[j] = 1.0
        v[j + 1:] = V64[j + 1:, j]
        Q -= float(tau[j]) * np.outer(v, v @ Q)
    return Q
