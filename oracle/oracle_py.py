"""Python face of the oracle.  TEST INFRASTRUCTURE ONLY (see oracle/dgemm_oracle.c).

  dgemm / tensor_stream / fill_uniform / rel_frobenius   ctypes over oracle/liboracle.so
                                                         (the C restatement)
  cpu_product                                            numpy (OpenBLAS dgemm, all host cores):
                                                         stands in for "Eigen's multithreaded CPU
                                                         product" -- Eigen itself is not installed
                                                         in the image and cannot be fetched
  Reference                                              ctypes over oracle/_ref/libeigencuda_ref.so,
                                                         the UNMODIFIED reference sources compiled
                                                         against cuBLAS (needs a GPU to run)
All matrices are float64, Fortran order (column-major, ld = rows), like Eigen::MatrixXd.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "liboracle.so")
REF_LIB = os.path.join(_HERE, "_ref", "libeigencuda_ref.so")

SEED = 20200210  # fixed operand matrix_id 0, tensor element i -> i + 1

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise RuntimeError(f"{LIB} missing: run `make -C oracle` or __graft_entry__.build()")
        L = C.CDLL(LIB)
        i64, dbl, vp, u64 = C.c_int64, C.c_double, C.c_void_p, C.c_uint64
        gemm_args = [C.c_int, C.c_int, i64, i64, i64, dbl, vp, i64, vp, i64, dbl, vp, i64]
        for name in ("oracle_dgemm", "oracle_dgemm_ld", "oracle_dgemm_fast"):
            getattr(L, name).restype = C.c_int
            getattr(L, name).argtypes = gemm_args
        L.oracle_gemm_shape_check.restype = C.c_int
        L.oracle_gemm_shape_check.argtypes = [i64, i64, i64, i64]
        L.oracle_fill_uniform.restype = None
        L.oracle_fill_uniform.argtypes = [u64, u64, vp, i64]
        L.oracle_fill_uniform_strided.restype = None
        L.oracle_fill_uniform_strided.argtypes = [u64, u64, u64, u64, vp, i64]
        L.oracle_uniform_at.restype = dbl
        L.oracle_uniform_at.argtypes = [u64, u64, u64]
        L.oracle_rel_frobenius.restype = dbl
        L.oracle_rel_frobenius.argtypes = [vp, vp, i64]
        L.oracle_is_approx.restype = C.c_int
        L.oracle_is_approx.argtypes = [vp, vp, i64, dbl]
        L.oracle_tensor_stream.restype = C.c_int
        L.oracle_tensor_stream.argtypes = [C.c_int, C.c_int, vp, i64, i64, C.POINTER(vp),
                                           C.POINTER(vp), i64, i64, i64]
        L.oracle_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _f(a):
    a = np.asarray(a, dtype=np.float64)
    return np.asfortranarray(a)


_FLAVOUR = {"plain": "oracle_dgemm", "ld": "oracle_dgemm_ld", "fast": "oracle_dgemm_fast"}


class ShapeMismatch(RuntimeError):
    """The std::runtime_error of reference src/cudapipeline.cc:26-28."""


def dgemm(A, B, transa=False, transb=False, alpha=1.0, beta=0.0, Cin=None, flavour="plain"):
    """alpha * op(A) @ op(B) + beta * C through the C restatement."""
    A, B = _f(A), _f(B)
    m, k = (A.shape[1], A.shape[0]) if transa else A.shape
    kb, n = (B.shape[1], B.shape[0]) if transb else B.shape
    if lib().oracle_gemm_shape_check(m, k, kb, n):
        raise ShapeMismatch("Shape mismatch in Cublas gemm")
    out = np.zeros((m, n), order="F") if Cin is None else _f(Cin).copy(order="F")
    rc = getattr(lib(), _FLAVOUR[flavour])(int(transa), int(transb), m, n, k, alpha, A.ctypes.data,
                                           max(A.shape[0], 1), B.ctypes.data, max(B.shape[0], 1),
                                           beta, out.ctypes.data, max(m, 1))
    assert rc == 0, rc
    return out


KINDS = {"right": 0, "left": 1, "left_t": 2, "basis": 3}


def tensor_stream(kind, F, tensor, flavour="plain"):
    """The reference's streaming loop (README.md:62-66) on the CPU."""
    F = _f(F)
    ts = [_f(t) for t in tensor]
    n = len(ts)
    if n == 0:
        return []
    tr, tc = ts[0].shape
    mode = KINDS[kind] if isinstance(kind, str) else int(kind)
    shape = {0: (tr, F.shape[1]), 1: (F.shape[0], tc), 2: (F.shape[1], tc),
             3: (F.shape[1], F.shape[1])}[mode]
    res = [np.zeros(shape, order="F") for _ in range(n)]
    ip = (C.c_void_p * n)(*[t.ctypes.data for t in ts])
    op = (C.c_void_p * n)(*[r.ctypes.data for r in res])
    fl = {"plain": 0, "ld": 1, "fast": 2}[flavour]
    rc = lib().oracle_tensor_stream(mode, fl, F.ctypes.data, F.shape[0], F.shape[1], ip, op, n, tr, tc)
    if rc == 1:
        raise ShapeMismatch("Shape mismatch in Cublas gemm")
    assert rc == 0, rc
    return res


def fill_uniform(seed, matrix_id, shape):
    out = np.empty(shape, dtype=np.float64, order="F")
    lib().oracle_fill_uniform(seed, matrix_id, out.ctypes.data, out.size)
    return out


def uniform_strided(seed, matrix_id, start, stride, count):
    out = np.empty(count, dtype=np.float64)
    lib().oracle_fill_uniform_strided(seed, matrix_id, start, stride, out.ctypes.data, count)
    return out


def rel_frobenius(a, b):
    a, b = _f(a), _f(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(lib().oracle_rel_frobenius(a.ctypes.data, b.ctypes.data, a.size))


def is_approx(a, b, prec=1e-12):
    """Eigen's isApprox (reference src/tests/test_dot.cc:31,50,87-89)."""
    return rel_frobenius(a, b) <= prec


def num_threads():
    return int(lib().oracle_num_threads())


def cpu_product(A, B):
    """OpenBLAS dgemm through numpy, all host cores -- the stand-in for Eigen's CPU product."""
    return np.asfortranarray(A @ B)


class Reference:
    """The unmodified reference (oracle/_ref/libeigencuda_ref.so): cuBLAS on the GPU through
    eigencuda::CudaPipeline / CudaMatrix, driven by oracle/ref_driver.cc."""

    def __init__(self):
        if not os.path.exists(REF_LIB):
            raise RuntimeError(f"{REF_LIB} missing (built only where /root/reference exists)")
        L = C.CDLL(REF_LIB)
        vp, lg, dbl = C.c_void_p, C.c_long, C.c_double
        L.ref_roundtrip.restype = C.c_int
        L.ref_roundtrip.argtypes = [vp, lg, lg, vp]
        L.ref_gemm.restype = C.c_int
        L.ref_gemm.argtypes = [vp, lg, lg, vp, lg, lg, vp]
        L.ref_tensor_stream.restype = C.c_int
        L.ref_tensor_stream.argtypes = [C.c_int, vp, lg, lg, C.POINTER(vp), C.POINTER(vp), lg, lg,
                                        lg, C.POINTER(dbl)]
        L.ref_tensor_stream_steps.restype = C.c_int
        L.ref_tensor_stream_steps.argtypes = [C.c_int, vp, lg, lg, C.POINTER(vp), C.POINTER(vp), lg, lg,
                                              lg, C.c_int, C.c_int, C.POINTER(dbl)]
        L.ref_gemm_time.restype = C.c_int
        L.ref_gemm_time.argtypes = [lg, C.c_int, C.c_int, C.POINTER(dbl)]
        L.ref_count_available_gpus.restype = lg
        self.L = L

    @staticmethod
    def available():
        return os.path.exists(REF_LIB)

    def roundtrip(self, B):
        B = _f(B)
        out = np.zeros_like(B, order="F")
        rc = self.L.ref_roundtrip(B.ctypes.data, B.shape[0], B.shape[1], out.ctypes.data)
        assert rc == 0
        return out

    def gemm(self, A, B):
        A, B = _f(A), _f(B)
        out = np.zeros((A.shape[0], B.shape[1]), order="F")
        rc = self.L.ref_gemm(A.ctypes.data, A.shape[0], A.shape[1], B.ctypes.data, B.shape[0],
                             B.shape[1], out.ctypes.data)
        if rc == 1:
            raise ShapeMismatch("Shape mismatch in Cublas gemm")
        assert rc == 0, rc
        return out

    def tensor_stream(self, side, F, tensor):
        """side 0: T_i * F (the reference's order), side 1: F * T_i.  Returns (results, seconds)."""
        F = _f(F)
        ts = [_f(t) for t in tensor]
        n = len(ts)
        tr, tc = ts[0].shape
        shape = (tr, F.shape[1]) if side == 0 else (F.shape[0], tc)
        res = [np.zeros(shape, order="F") for _ in range(n)]
        ip = (C.c_void_p * n)(*[t.ctypes.data for t in ts])
        op = (C.c_void_p * n)(*[r.ctypes.data for r in res])
        el = C.c_double(0)
        rc = self.L.ref_tensor_stream(side, F.ctypes.data, F.shape[0], F.shape[1], ip, op, n, tr, tc,
                                      C.byref(el))
        if rc == 1:
            raise ShapeMismatch("Shape mismatch in Cublas gemm")
        assert rc == 0, rc
        return res, el.value

    def tensor_stream_steps(self, side, F, tensor_ptrs, count, t_shape, warmup, steps, check_index=None):
        """Benchmark form: `tensor_ptrs` is a ctypes array of `count` host addresses of column-major
        t_shape matrices; runs warmup + steps passes of the reference loop, returns (seconds per timed
        pass, result matrix number `check_index` of the last pass or None)."""
        F = _f(F)
        tr, tc = t_shape
        shape = (tr, F.shape[1]) if side == 0 else (F.shape[0], tc)
        op = (C.c_void_p * count)()
        keep = None
        if check_index is not None:
            keep = np.zeros(shape, order="F")
            op[check_index] = keep.ctypes.data
        el = (C.c_double * steps)()
        rc = self.L.ref_tensor_stream_steps(side, F.ctypes.data, F.shape[0], F.shape[1], tensor_ptrs, op,
                                            count, tr, tc, warmup, steps, el)
        if rc == 1:
            raise ShapeMismatch("Shape mismatch in Cublas gemm")
        assert rc == 0, rc
        return list(el), keep

    def gemm_time_ms(self, n, warmup=3, reps=10):
        ms = C.c_double(0)
        rc = self.L.ref_gemm_time(n, warmup, reps, C.byref(ms))
        assert rc == 0, rc
        return ms.value

    def count_available_gpus(self):
        return int(self.L.ref_count_available_gpus())
