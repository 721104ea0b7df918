"""cuda-qr_b200 -- B200 (sm_100a) drop-in for the QR hot path of brian-kelley/CUDA-QR.

The product is ``libcudaqr_b200.so`` (hand-written CUDA behind the C ABI declared in
``include/cudaqr_b200.h``).  This module is the thin host-side mirror of the reference's
interface over that ABI:

* legacy, host-pointer entry points with the reference's names and argument meaning
  (``mmqr``, ``explicitQR``, ``dgemm``, ``identity``, ``getPanelDims``, ``printMat``;
  qr.c:15-18,47,55 / qr.cu:30-31,49,475) operating on column-major numpy arrays;
* ``Context``: the device-resident API on raw device pointers (torch tensors are only the
  allocator here -- ``tensor.data_ptr()`` is what crosses the boundary).

There is NO CPU fallback: if the shared library is missing this import raises, and every
compute call needs a CUDA device.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("CQR_LIB") or os.path.join(_HERE, "libcudaqr_b200.so")   # CQR_LIB: an instrumented build (tools/mma_trace.py)
_FP = ctypes.POINTER(ctypes.c_float)
_IP = ctypes.POINTER(ctypes.c_int)
_VP = ctypes.c_void_p

OPT_GEMM, OPT_OUTER_BLOCK, OPT_TILE_ROWS, OPT_SPLITK, OPT_LOOKAHEAD, OPT_PANEL, OPT_FLAT_TSQR, OPT_PARTITION = 1, 2, 3, 4, 5, 6, 7, 8
GEMM_SIMT, GEMM_TF32X3 = 0, 1
# OPT_FLAT_TSQR values: the leaf of the R-only cqr_tsqr_r on >= 16384 rows
TSQR_LEAF_TILE, TSQR_LEAF_FLAT, TSQR_LEAF_MMA, TSQR_LEAF_PAIR, TSQR_LEAF_GRAM = 0, 1, 2, 3, 4
TSQR_LEAF_DEFAULT = TSQR_LEAF_GRAM

# Every symbol include/cudaqr_b200.h declares (tests check the .so exports all of them).
EXPORTS = [
    "getPanelDims", "mmqr", "mmqr_alloc", "explicitQR", "dgemm", "identity", "printMat",
    "cqr_create", "cqr_destroy", "cqr_set_stream", "cqr_set_option", "cqr_get_option", "cqr_synchronize",
    "cqr_error_string", "cqr_launch_count", "cqr_profile_begin", "cqr_profile_end", "cqr_profile_timeline", "cqr_reserve", "cqr_geqrf", "cqr_geqrf_partial", "cqr_extract_r", "cqr_form_q",
    "cqr_apply_q", "cqr_solve_ls", "cqr_tsqr_r", "cqr_tsqr_gram_info", "cqr_tsqr_factor", "cqr_tsqr_form_q", "cqr_stack_qr", "cqr_stack_form_q",
    "cqr_geqrf_batched", "cqr_gemm", "cqr_gemm_tf32x3", "cqr_set_identity", "cqr_version",
    "cqr_compare_cusolver_sgeqrf", "mmqr_reference_format", "cqr_mmqr_reference_format",
    "cqr_dist_export", "cqr_dist_attach", "cqr_dist_detach", "cqr_tsqr_dist_r",
    "mmqr_f64", "explicitQR_f64", "cqr_dgeqrf", "cqr_dform_q", "cqr_dapply_q", "cqr_dextract_r",
]


def build_library(force: bool = False) -> str:
    """nvcc-compile csrc/*.cu for sm_100a into libcudaqr_b200.so (cross-compiles without a GPU)."""
    src_dir = os.path.join(_HERE, "csrc")
    if force:
        subprocess.check_call(["make", "-C", src_dir, "clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", src_dir, "-j8"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class CudaQRError(RuntimeError):
    pass


def _load() -> ctypes.CDLL:
    if not os.path.exists(_LIB_PATH):
        raise ImportError(
            f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C cuda-qr_b200/csrc). There is no CPU fallback.")
    lib = ctypes.CDLL(_LIB_PATH)
    i, ll, f = ctypes.c_int, ctypes.c_longlong, ctypes.c_float
    lib.getPanelDims.argtypes = [i, i, _IP, _IP]
    lib.mmqr.argtypes = [_FP, _FP, i, i]
    lib.mmqr_alloc.argtypes = [_FP, ctypes.POINTER(_FP), i, i]
    lib.explicitQR.argtypes = [_FP, _FP, _FP, _FP, i, i]
    lib.dgemm.argtypes = [_FP, _FP, _FP, i, i, i]
    lib.identity.argtypes = [_FP, i]
    lib.printMat.argtypes = [_FP, i, i]
    lib.cqr_create.argtypes = [ctypes.POINTER(_VP), i]
    lib.cqr_destroy.argtypes = [_VP]
    lib.cqr_set_stream.argtypes = [_VP, _VP]
    lib.cqr_set_option.argtypes = [_VP, i, i]
    lib.cqr_get_option.argtypes = [_VP, i, _IP]
    lib.cqr_synchronize.argtypes = [_VP]
    lib.cqr_error_string.argtypes = [i]
    lib.cqr_error_string.restype = ctypes.c_char_p
    lib.cqr_launch_count.argtypes = [_VP]
    lib.cqr_launch_count.restype = ll
    lib.cqr_reserve.argtypes = [_VP, ctypes.c_size_t]
    lib.cqr_profile_begin.argtypes = [_VP]
    lib.cqr_profile_end.argtypes = [_VP, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                    ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ll), i]
    lib.cqr_profile_timeline.argtypes = [_VP, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                         ctypes.POINTER(ctypes.c_int), i]
    lib.cqr_geqrf.argtypes = [_VP, _VP, i, i, i, _VP]
    lib.cqr_geqrf_partial.argtypes = [_VP, _VP, i, i, i, i, _VP]
    lib.cqr_extract_r.argtypes = [_VP, _VP, i, i, i, _VP, i, i]
    lib.cqr_form_q.argtypes = [_VP, _VP, i, i, i, _VP, _VP, i, i]
    lib.cqr_apply_q.argtypes = [_VP, i, _VP, i, i, i, _VP, _VP, i, i]
    lib.cqr_solve_ls.argtypes = [_VP, _VP, i, i, i, _VP, _VP, i, i]
    lib.cqr_tsqr_r.argtypes = [_VP, _VP, i, ll, i, _VP, i]
    lib.cqr_tsqr_factor.argtypes = [_VP, _VP, i, ll, i, _VP, i]
    lib.cqr_tsqr_gram_info.argtypes = [_VP, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)]
    lib.cqr_tsqr_form_q.argtypes = [_VP, _VP, i, _VP, i]
    lib.cqr_stack_qr.argtypes = [_VP, _VP, i, i, i, _VP, _VP, i]
    lib.cqr_stack_form_q.argtypes = [_VP, _VP, i, i, i, _VP, _VP, i, _VP, i]
    lib.cqr_geqrf_batched.argtypes = [_VP, _VP, i, ll, i, i, i, _VP]
    lib.cqr_gemm.argtypes = [_VP, i, i, i, i, f, _VP, i, _VP, i, f, _VP, i]
    lib.cqr_gemm_tf32x3.argtypes = [_VP, i, i, i, i, f, _VP, i, _VP, i, f, _VP, i]
    lib.cqr_set_identity.argtypes = [_VP, _VP, i, i, i]
    lib.cqr_version.restype = ctypes.c_char_p
    lib.cqr_compare_cusolver_sgeqrf.argtypes = [_FP, _FP, i, i]
    lib.mmqr_reference_format.argtypes = [_FP, _FP, i, i]
    lib.cqr_mmqr_reference_format.argtypes = [_VP, _VP, i, i, i, _VP]
    _DP = ctypes.POINTER(ctypes.c_double)
    lib.mmqr_f64.argtypes = [_DP, _DP, i, i]
    lib.explicitQR_f64.argtypes = [_DP, _DP, _DP, _DP, i, i]
    lib.cqr_dgeqrf.argtypes = [_VP, _VP, i, i, i, _VP]
    lib.cqr_dform_q.argtypes = [_VP, _VP, i, i, i, _VP, _VP, i, i]
    lib.cqr_dapply_q.argtypes = [_VP, i, _VP, i, i, i, _VP, _VP, i, i]
    lib.cqr_dextract_r.argtypes = [_VP, _VP, i, i, i, _VP, i, i]
    lib.cqr_dist_export.argtypes = [_VP, ctypes.c_char_p]
    lib.cqr_dist_attach.argtypes = [_VP, i, i, ctypes.c_char_p]
    lib.cqr_dist_detach.argtypes = [_VP]
    lib.cqr_tsqr_dist_r.argtypes = [_VP, _VP, i, ll, i, _VP, i]
    return lib


lib = _load()


def version() -> str:
    return lib.cqr_version().decode()


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise CudaQRError(f"{what} failed: {lib.cqr_error_string(rc).decode()} (status {rc})")


def _host(a: np.ndarray, shape=None) -> np.ndarray:
    if not (isinstance(a, np.ndarray) and a.dtype == np.float32 and (a.flags["F_CONTIGUOUS"] or a.ndim == 1)):
        raise TypeError("expected a column-major (Fortran-order) float32 numpy array")
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError(f"expected shape {shape}, got {a.shape}")
    return a


def _p(a: np.ndarray):
    return a.ctypes.data_as(_FP)


# ----------------------------------------------------------------------------------------------
# Legacy interface (host buffers; same names / argument meaning / blocking behaviour as the reference)
# ----------------------------------------------------------------------------------------------
def getPanelDims(m: int, n: int):
    """qr.cu:49-55 -> (rowPanels, colPanels); tau buffers hold rowPanels*colPanels*4 floats (qr.cu:764)."""
    rp, cp = ctypes.c_int(), ctypes.c_int()
    lib.getPanelDims(m, n, ctypes.byref(rp), ctypes.byref(cp))
    return rp.value, cp.value


def tau_size(m: int, n: int) -> int:
    rp, cp = getPanelDims(m, n)
    return rp * cp * 4


def mmqr(mat: np.ndarray, tau: np.ndarray | None = None) -> np.ndarray:
    """qr.cu:475 -- factor `mat` (m x n, column-major float32) IN PLACE; returns tau (caller-sized
    via getPanelDims when given).  Includes the H2D / D2H copies, like the reference."""
    _host(mat)
    m, n = mat.shape
    if m < n or n < 1:
        raise ValueError("mmqr needs m >= n >= 1 (qr.cu:736)")
    if tau is None:
        tau = np.empty(tau_size(m, n), dtype=np.float32)
    elif tau.size < tau_size(m, n) or tau.dtype != np.float32:
        raise ValueError("tau must hold rowPanels*colPanels*4 float32 values (getPanelDims)")
    lib.mmqr(_p(mat), _p(tau), m, n)
    return tau


def mmqr_reference_format(mat: np.ndarray, tau: np.ndarray | None = None) -> np.ndarray:
    """mmqr leaving the REFERENCE's storage (window reflector segments + the tau grid of qr.c:300-304, PR = 64, PC = 4) so
    that the reference's own explicitQR consumes it; legal shapes only (m = 64 + 60 k, 4 | n, n <= m)."""
    _host(mat)
    m, n = mat.shape
    if m < 64 or n < 4 or n > m or (m - 64) % 60 or n % 4:
        raise ValueError("the reference's window grid needs m = 64 + 60 k and n a multiple of 4, n <= m")
    if tau is None:
        tau = np.empty(tau_size(m, n), dtype=np.float32)
    elif tau.size < tau_size(m, n) or tau.dtype != np.float32:
        raise ValueError("tau must hold rowPanels*colPanels*4 float32 values (getPanelDims)")
    lib.mmqr_reference_format(_p(mat), _p(tau), m, n)
    return tau


def mmqr_f64(mat: np.ndarray, tau: np.ndarray | None = None) -> np.ndarray:
    """mmqr in double precision (the reference's contemplated `Scalar double`, qr.c:9): factor the column-major float64
    `mat` IN PLACE, return tau (rowPanels*colPanels*4 doubles, the first n used)."""
    if not (isinstance(mat, np.ndarray) and mat.dtype == np.float64 and mat.flags["F_CONTIGUOUS"]):
        raise TypeError("expected a column-major (Fortran-order) float64 numpy array")
    m, n = mat.shape
    if m < n or n < 1:
        raise ValueError("mmqr needs m >= n >= 1 (qr.cu:736)")
    if tau is None:
        tau = np.empty(tau_size(m, n), dtype=np.float64)
    dp = ctypes.POINTER(ctypes.c_double)
    lib.mmqr_f64(mat.ctypes.data_as(dp), tau.ctypes.data_as(dp), m, n)
    return tau


def explicitQR_f64(A: np.ndarray, tau: np.ndarray):
    """(Q m x m, R m x n) in double precision from mmqr_f64's storage."""
    m, n = A.shape
    Q = np.empty((m, m), dtype=np.float64, order="F")
    R = np.empty((m, n), dtype=np.float64, order="F")
    dp = ctypes.POINTER(ctypes.c_double)
    t = np.ascontiguousarray(tau, dtype=np.float64)
    lib.explicitQR_f64(A.ctypes.data_as(dp), t.ctypes.data_as(dp), Q.ctypes.data_as(dp), R.ctypes.data_as(dp), m, n)
    return Q, R


def explicitQR(A: np.ndarray, tau: np.ndarray):
    """qr.c:330 -- (Q m x m, R m x n) from mmqr's in-place storage and tau."""
    _host(A)
    m, n = A.shape
    Q = np.empty((m, m), dtype=np.float32, order="F")
    R = np.empty((m, n), dtype=np.float32, order="F")
    lib.explicitQR(_p(A), _p(np.ascontiguousarray(tau, dtype=np.float32)), _p(Q), _p(R), m, n)
    return Q, R


def dgemm(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """qr.c:443 -- C(k x n) = A(k x m) B(m x n), fp32."""
    _host(A), _host(B)
    k, m = A.shape
    if B.shape[0] != m:
        raise ValueError("inner dimensions differ")
    n = B.shape[1]
    C = np.empty((k, n), dtype=np.float32, order="F")
    lib.dgemm(_p(A), _p(B), _p(C), k, m, n)
    return C


def identity(m: int) -> np.ndarray:
    """qr.c:316."""
    A = np.empty((m, m), dtype=np.float32, order="F")
    lib.identity(_p(A), m)
    return A


def printMat(mat: np.ndarray) -> None:
    """qr.c:21 (prints through C stdio)."""
    _host(mat)
    lib.printMat(_p(mat), mat.shape[0], mat.shape[1])


# ----------------------------------------------------------------------------------------------
# Device-resident API
# ----------------------------------------------------------------------------------------------
def _dptr(t) -> int:
    """Device pointer of a torch CUDA tensor (or a raw int)."""
    if isinstance(t, int):
        return t
    if t is None:
        return 0
    if not t.is_cuda or str(t.dtype) != "torch.float32":
        raise TypeError("expected a float32 CUDA tensor")
    return t.data_ptr()


def colmajor(m: int, n: int, device="cuda", ld: int | None = None, dtype=None):
    """Allocate an m x n column-major fp32 (or `dtype`) device matrix as a torch view (ld >= m): returns a tensor
    `a` with a[i, j] at offset i + j*ld, i.e. a = storage(n, ld).T[:m]."""
    import torch
    ld = ld or m
    return torch.empty((n, ld), dtype=dtype or torch.float32, device=device).t()[:m]


def _dptr64(t) -> int:
    if t is None:
        return 0
    if not t.is_cuda or str(t.dtype) != "torch.float64":
        raise TypeError("expected a float64 CUDA tensor")
    return t.data_ptr()


def to_colmajor(x, ld: int | None = None):
    """Copy a 2-D torch tensor into column-major device storage."""
    out = colmajor(x.shape[0], x.shape[1], device=x.device, ld=ld)
    out.copy_(x)
    return out


def _ld(t) -> int:
    if t.dim() != 2 or t.stride(0) != 1:
        raise ValueError("expected a column-major 2-D tensor (stride(0) == 1); see colmajor()")
    return t.stride(1) if t.shape[1] > 1 else max(t.shape[0], t.stride(1))


class Context:
    """Owns a cqr_context (workspace + stream) on one device."""

    def __init__(self, device: int = 0, stream: int | None = None):
        h = _VP()
        _check(lib.cqr_create(ctypes.byref(h), device), "cqr_create")
        self.h = h
        self.device = device
        if stream is not None:
            self.set_stream(stream)

    def close(self):
        if getattr(self, "h", None):
            lib.cqr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int):
        _check(lib.cqr_set_stream(self.h, _VP(cuda_stream)), "cqr_set_stream")

    def use_torch_stream(self):
        import torch
        self.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def set_option(self, opt: int, value: int):
        _check(lib.cqr_set_option(self.h, opt, value), "cqr_set_option")

    def get_option(self, opt: int) -> int:
        v = ctypes.c_int()
        _check(lib.cqr_get_option(self.h, opt, ctypes.byref(v)), "cqr_get_option")
        return v.value

    def synchronize(self):
        _check(lib.cqr_synchronize(self.h), "cqr_synchronize")

    def launch_count(self) -> int:
        return int(lib.cqr_launch_count(self.h))

    PROF_CLASSES = ("panel", "gemm_tn", "gemm_nn", "misc", "chain_tn", "chain_nn", "chain_misc")

    def profile_begin(self):
        _check(lib.cqr_profile_begin(self.h), "cqr_profile_begin")

    def profile_end(self) -> dict:
        """{class: {"ms", "flops", "bytes", "launches"}} for the launches since profile_begin()."""
        n = len(self.PROF_CLASSES)
        ms, fl, by = (ctypes.c_double * n)(), (ctypes.c_double * n)(), (ctypes.c_double * n)()
        la = (ctypes.c_longlong * n)()
        _check(lib.cqr_profile_end(self.h, ms, fl, by, la, n), "cqr_profile_end")
        return {k: {"ms": ms[j], "flops": fl[j], "bytes": by[j], "launches": int(la[j])}
                for j, k in enumerate(self.PROF_CLASSES)}

    def profile_timeline(self, cap: int = 1 << 16):
        """[(start_ms, end_ms, class)] of every bracketed launch group since profile_begin(); call before profile_end()."""
        t0, t1, cl = (ctypes.c_double * cap)(), (ctypes.c_double * cap)(), (ctypes.c_int * cap)()
        n = lib.cqr_profile_timeline(self.h, t0, t1, cl, cap)
        if n < 0:
            raise RuntimeError("cqr_profile_timeline failed")
        return [(t0[i], t1[i], self.PROF_CLASSES[cl[i]]) for i in range(n)]

    def reserve(self, nbytes: int):
        _check(lib.cqr_reserve(self.h, nbytes), "cqr_reserve")

    # -- blocked Householder QR -----------------------------------------------------------
    def geqrf(self, A, tau):
        m, n = A.shape
        _check(lib.cqr_geqrf(self.h, _dptr(A), _ld(A), m, n, _dptr(tau)), "cqr_geqrf")

    def geqrf_partial(self, A, tau, nfact: int):
        """QR of the first nfact columns of A, Q^T applied to all of A's columns (cqr_geqrf_partial)."""
        m, n = A.shape
        _check(lib.cqr_geqrf_partial(self.h, _dptr(A), _ld(A), m, n, nfact, _dptr(tau)), "cqr_geqrf_partial")

    def extract_r(self, A, R):
        m, n = A.shape
        _check(lib.cqr_extract_r(self.h, _dptr(A), _ld(A), m, n, _dptr(R), _ld(R), R.shape[0]), "cqr_extract_r")

    def form_q(self, A, tau, Q):
        m, n = A.shape
        _check(lib.cqr_form_q(self.h, _dptr(A), _ld(A), m, n, _dptr(tau), _dptr(Q), _ld(Q), Q.shape[1]), "cqr_form_q")

    def apply_q(self, A, tau, C, trans: bool):
        m, n = A.shape
        _check(lib.cqr_apply_q(self.h, 1 if trans else 0, _dptr(A), _ld(A), m, n, _dptr(tau), _dptr(C), _ld(C),
                               C.shape[1]), "cqr_apply_q")

    def solve_ls(self, A, tau, B):
        """min ||A x - b|| for every column b of B (m x nrhs) from geqrf's output: X is left in B[:n]."""
        m, n = A.shape
        _check(lib.cqr_solve_ls(self.h, _dptr(A), _ld(A), m, n, _dptr(tau), _dptr(B), _ld(B), B.shape[1]), "cqr_solve_ls")

    # -- TSQR -----------------------------------------------------------------------------
    def tsqr_r(self, A, R):
        m, n = A.shape
        _check(lib.cqr_tsqr_r(self.h, _dptr(A), _ld(A), m, n, _dptr(R), _ld(R)), "cqr_tsqr_r")

    def tsqr_gram_info(self):
        """(bound, householder) of the Gram leaf (OPT_FLAT_TSQR = 4) of the last tsqr_r: bound = n ||Rs^-1||_F^2 >= cond_2 of the
        unit-diagonal Gram matrix (-1: breakdown / out-of-range scale), householder = True when the gated Householder leaf ran."""
        b, g = ctypes.c_double(0.0), ctypes.c_int(0)
        _check(lib.cqr_tsqr_gram_info(self.h, ctypes.byref(b), ctypes.byref(g)), "cqr_tsqr_gram_info")
        return b.value, bool(g.value)

    def tsqr_factor(self, A, R):
        m, n = A.shape
        _check(lib.cqr_tsqr_factor(self.h, _dptr(A), _ld(A), m, n, _dptr(R), _ld(R)), "cqr_tsqr_factor")

    # -- double precision (f64_qr.cu) -----------------------------------------------------------
    def dgeqrf(self, A, tau):
        m, n = A.shape
        _check(lib.cqr_dgeqrf(self.h, _dptr64(A), _ld(A), m, n, _dptr64(tau)), "cqr_dgeqrf")

    def dform_q(self, A, tau, Q):
        m, n = A.shape
        _check(lib.cqr_dform_q(self.h, _dptr64(A), _ld(A), m, n, _dptr64(tau), _dptr64(Q), _ld(Q), Q.shape[1]), "cqr_dform_q")

    def dapply_q(self, A, tau, C, trans: bool):
        m, n = A.shape
        _check(lib.cqr_dapply_q(self.h, 1 if trans else 0, _dptr64(A), _ld(A), m, n, _dptr64(tau), _dptr64(C), _ld(C), C.shape[1]),
               "cqr_dapply_q")

    def dextract_r(self, A, R):
        m, n = A.shape
        _check(lib.cqr_dextract_r(self.h, _dptr64(A), _ld(A), m, n, _dptr64(R), _ld(R), R.shape[0]), "cqr_dextract_r")

    # -- row-partitioned TSQR across GPUs, R tree over peer memory (cqr_dist_*) -----------------
    def dist_export(self) -> bytes:
        buf = ctypes.create_string_buffer(64)
        _check(lib.cqr_dist_export(self.h, buf), "cqr_dist_export")
        return buf.raw

    def dist_attach(self, rank: int, world: int, handles):
        blob = b"".join(handles)
        if len(blob) != 64 * world:
            raise ValueError("need one 64-byte handle per rank")
        _check(lib.cqr_dist_attach(self.h, rank, world, blob), "cqr_dist_attach")

    def dist_detach(self):
        _check(lib.cqr_dist_detach(self.h), "cqr_dist_detach")

    def tsqr_dist_r(self, A, R):
        m, n = A.shape
        _check(lib.cqr_tsqr_dist_r(self.h, _dptr(A), _ld(A), m, n, _dptr(R), _ld(R)), "cqr_tsqr_dist_r")

    def tsqr_form_q(self, Q, X=None):
        _check(lib.cqr_tsqr_form_q(self.h, _dptr(X), _ld(X) if X is not None else 0, _dptr(Q), _ld(Q)),
               "cqr_tsqr_form_q")

    def stack_qr(self, Rs, n: int, tau, R):
        rows = Rs.shape[0]
        _check(lib.cqr_stack_qr(self.h, _dptr(Rs), _ld(Rs), rows // n, n, _dptr(tau), _dptr(R), _ld(R)), "cqr_stack_qr")

    def stack_form_q(self, Rs, n: int, tau, Qs, X=None):
        rows = Rs.shape[0]
        _check(lib.cqr_stack_form_q(self.h, _dptr(Rs), _ld(Rs), rows // n, n, _dptr(tau), _dptr(X),
                                    _ld(X) if X is not None else 0, _dptr(Qs), _ld(Qs)), "cqr_stack_form_q")

    # -- batched --------------------------------------------------------------------------
    def geqrf_batched(self, A3, tau):
        """A3: torch tensor of shape (batch, n, lda) holding column-major matrices (A3[b].T[:m] is matrix b)."""
        batch, n, lda = A3.shape
        m = lda
        _check(lib.cqr_geqrf_batched(self.h, _dptr(A3), lda, A3.stride(0), m, n, batch, _dptr(tau)),
               "cqr_geqrf_batched")

    def gemm(self, A, B, D, trans_a: bool = False, alpha: float = 1.0, beta: float = 0.0):
        M, N = D.shape
        K = A.shape[0] if trans_a else A.shape[1]
        _check(lib.cqr_gemm(self.h, 1 if trans_a else 0, M, N, K, alpha, _dptr(A), _ld(A), _dptr(B), _ld(B), beta,
                            _dptr(D), _ld(D)), "cqr_gemm")

    def gemm_tf32x3(self, A, B, D, trans_a: bool = False, alpha: float = 1.0, beta: float = 0.0):
        """D = op(A) B on the tcgen05 3xTF32 kernels; raises if the shape cannot take the TMA path."""
        M, N = D.shape
        K = A.shape[0] if trans_a else A.shape[1]
        _check(lib.cqr_gemm_tf32x3(self.h, 1 if trans_a else 0, M, N, K, alpha, _dptr(A), _ld(A), _dptr(B), _ld(B),
                                   beta, _dptr(D), _ld(D)), "cqr_gemm_tf32x3")

    def set_identity(self, A):
        _check(lib.cqr_set_identity(self.h, _dptr(A), _ld(A), A.shape[0], A.shape[1]), "cqr_set_identity")
