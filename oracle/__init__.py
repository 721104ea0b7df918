"""oracle/ -- TEST INFRASTRUCTURE, never imported by the product package.

Two CPU checkers for the MMQR hot path of brian-kelley/CUDA-QR:

* ``Port``  -- ctypes binding of ``liboracle_mmqr.so`` (``mmqr_oracle.c``), the
  run-time-(PR, PC) restatement of qr.c:55-313 / 330-438.
* ``Ref``   -- ctypes binding of the UNMODIFIED reference ``qr.c`` compiled by
  ``oracle/build_ref.sh`` into ``oracle/_ref/libref_qr_<PR>_<PC>.so``.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_FP = ctypes.POINTER(ctypes.c_float)
_IP = ctypes.POINTER(ctypes.c_int)


def _fptr(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["F_CONTIGUOUS"] or a.ndim == 1
    return a.ctypes.data_as(_FP)


def build(force: bool = False) -> None:
    """Compile the restatement and (when /root/reference is present) oracle/_ref."""
    so = os.path.join(_HERE, "liboracle_mmqr.so")
    src = os.path.join(_HERE, "mmqr_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle_mmqr.so"], stdout=subprocess.DEVNULL)
    ref_dir = os.path.join(_HERE, "_ref")
    if os.path.exists("/root/reference/qr.c") and (force or not os.path.isdir(ref_dir)
                                                   or len(os.listdir(ref_dir)) < 3):
        subprocess.check_call([os.path.join(_HERE, "build_ref.sh")], stdout=subprocess.DEVNULL)


def panel_dims(m: int, n: int, PR: int, PC: int):
    """qr.c:47-53."""
    col = -(-n // PC)
    row = 1 + (-(-(m - PR) // (PR - PC)) if m > PR else 0)
    return row, col


def legal_shape(m: int, n: int, PR: int, PC: int) -> bool:
    """shapes the reference factors correctly (it never checks)."""
    return m >= PR and n <= m and (m - PR) % (PR - PC) == 0 and n % PC == 0 and PR % PC == 0


class Port:
    """The C restatement (mmqr_oracle.c)."""

    kind = "port"

    def __init__(self):
        build()
        self.lib = ctypes.CDLL(os.path.join(_HERE, "liboracle_mmqr.so"))
        L = self.lib
        L.oracle_mmqr.argtypes = [_FP, _FP] + [ctypes.c_int] * 4
        L.oracle_mmqr.restype = ctypes.c_int
        L.oracle_explicitQR.argtypes = [_FP, _FP, _FP, _FP] + [ctypes.c_int] * 4
        L.oracle_explicitQR.restype = ctypes.c_int
        L.oracle_dgemm.argtypes = [_FP, _FP, _FP] + [ctypes.c_int] * 3
        L.oracle_identity.argtypes = [_FP, ctypes.c_int]
        L.oracle_fill_rand.argtypes = [_FP, ctypes.c_long, ctypes.c_uint]
        L.oracle_getPanelDims.argtypes = [ctypes.c_int] * 4 + [_IP, _IP]

    def rand_matrix(self, m: int, n: int, seed: int = 12) -> np.ndarray:
        """qr.c:468-474: srand(seed); A[i] = (float)rand()/RAND_MAX, column-major."""
        a = np.empty((m, n), dtype=np.float32, order="F")
        self.lib.oracle_fill_rand(_fptr(a), m * n, seed)
        return a

    def mmqr(self, A: np.ndarray, PR: int, PC: int):
        """Returns (RV, tau): in-place factored copy of A and the tau array."""
        m, n = A.shape
        rv = np.array(A, dtype=np.float32, order="F", copy=True)
        rp, cp = panel_dims(m, n, PR, PC)
        tau = np.zeros(rp * cp * PC, dtype=np.float32)
        rc = self.lib.oracle_mmqr(_fptr(rv), _fptr(tau), m, n, PR, PC)
        if rc != 0:
            raise RuntimeError("oracle_mmqr failed")
        return rv, tau

    def explicitQR(self, RV: np.ndarray, tau: np.ndarray, PR: int, PC: int):
        m, n = RV.shape
        Q = np.empty((m, m), dtype=np.float32, order="F")
        R = np.empty((m, n), dtype=np.float32, order="F")
        rv = np.asfortranarray(RV, dtype=np.float32)
        rc = self.lib.oracle_explicitQR(_fptr(rv), _fptr(np.ascontiguousarray(tau)), _fptr(Q), _fptr(R),
                                        m, n, PR, PC)
        if rc != 0:
            raise RuntimeError("oracle_explicitQR failed")
        return Q, R

    def dgemm(self, A: np.ndarray, B: np.ndarray) -> np.ndarray:
        k, m = A.shape
        m2, n = B.shape
        assert m == m2
        C = np.empty((k, n), dtype=np.float32, order="F")
        self.lib.oracle_dgemm(_fptr(np.asfortranarray(A, dtype=np.float32)),
                              _fptr(np.asfortranarray(B, dtype=np.float32)), _fptr(C), k, m, n)
        return C


class Ref:
    """The unmodified reference qr.c, compiled for one (PR, PC) (oracle/build_ref.sh)."""

    kind = "reference"

    def __init__(self, PR: int = 4, PC: int = 2):
        build()
        path = os.path.join(_HERE, "_ref", f"libref_qr_{PR}_{PC}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.PR, self.PC = PR, PC
        self.lib = ctypes.CDLL(path)
        L = self.lib
        L.mmqr.argtypes = [_FP, ctypes.POINTER(_FP), ctypes.c_int, ctypes.c_int]
        L.explicitQR.argtypes = [_FP, _FP, _FP, _FP, ctypes.c_int, ctypes.c_int]
        L.dgemm.argtypes = [_FP, _FP, _FP] + [ctypes.c_int] * 3
        L.getPanelDims.argtypes = [ctypes.c_int, ctypes.c_int, _IP, _IP]
        self.libc = ctypes.CDLL(None)
        self.libc.free.argtypes = [ctypes.c_void_p]

    @staticmethod
    def available(PR: int = 4, PC: int = 2) -> bool:
        try:
            build()
        except Exception:
            pass
        return os.path.exists(os.path.join(_HERE, "_ref", f"libref_qr_{PR}_{PC}.so"))

    def panel_dims(self, m: int, n: int):
        rp, cp = ctypes.c_int(), ctypes.c_int()
        self.lib.getPanelDims(m, n, ctypes.byref(rp), ctypes.byref(cp))
        return rp.value, cp.value

    def mmqr(self, A: np.ndarray):
        """qr.c:55 -- callee mallocs *tau; copied out and freed here."""
        m, n = A.shape
        rv = np.array(A, dtype=np.float32, order="F", copy=True)
        tau_p = _FP()
        self.lib.mmqr(_fptr(rv), ctypes.byref(tau_p), m, n)
        rp, cp = self.panel_dims(m, n)
        cnt = rp * cp * self.PC
        tau = np.ctypeslib.as_array(tau_p, shape=(cnt,)).copy()
        self.libc.free(ctypes.cast(tau_p, ctypes.c_void_p))
        return rv, tau

    def explicitQR(self, RV: np.ndarray, tau: np.ndarray):
        m, n = RV.shape
        Q = np.empty((m, m), dtype=np.float32, order="F")
        R = np.empty((m, n), dtype=np.float32, order="F")
        rv = np.asfortranarray(RV, dtype=np.float32)
        self.lib.explicitQR(_fptr(rv), _fptr(np.ascontiguousarray(tau)), _fptr(Q), _fptr(R), m, n)
        return Q, R

    def dgemm(self, A: np.ndarray, B: np.ndarray) -> np.ndarray:
        k, m = A.shape
        _, n = B.shape
        C = np.empty((k, n), dtype=np.float32, order="F")
        self.lib.dgemm(_fptr(np.asfortranarray(A, dtype=np.float32)),
                       _fptr(np.asfortranarray(B, dtype=np.float32)), _fptr(C), k, m, n)
        return C


# --------------------------------------------------------------------------------------
# glibc rand() restated in numpy (TYPE_3 additive-feedback generator, r[i] = r[i-3] +
# r[i-31]), so the reference's srand(12) input recipe (qr.c:468-474) is reproducible on
# any libc.  tests/test_oracle.py checks it against the C library's own stream.
# --------------------------------------------------------------------------------------
def glibc_rand(seed: int, count: int) -> np.ndarray:
    r = np.zeros(count + 344, dtype=np.int64)
    r[0] = seed if seed != 0 else 1
    for i in range(1, 31):
        hi, lo = divmod(int(r[i - 1]), 127773)
        w = 16807 * lo - 2836 * hi
        r[i] = w + 2147483647 if w < 0 else w
    for i in range(31, 34):
        r[i] = r[i - 31]
    M32 = (1 << 32) - 1
    # r[i] = (r[i-3] + r[i-31]) mod 2^32: the lag-3 term blocks vectorisation beyond 3 lanes,
    # so run the recurrence in python ints on a list (fast enough for test-sized inputs).
    rl = [int(x) & M32 for x in r[:34]] + [0] * (count + 310)
    for i in range(34, count + 344):
        rl[i] = (rl[i - 3] + rl[i - 31]) & M32
    out = np.array(rl[344:344 + count], dtype=np.uint64)
    return (out >> np.uint64(1)).astype(np.int64)


def rand_matrix(m: int, n: int, seed: int = 12) -> np.ndarray:
    """The reference's input recipe without libc: (float)rand()/RAND_MAX, column-major."""
    vals = glibc_rand(seed, m * n)
    a = (vals.astype(np.float32) / np.float32(2147483647.0)).astype(np.float32)
    return np.asfortranarray(a.reshape((n, m)).T)
