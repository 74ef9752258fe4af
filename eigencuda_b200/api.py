"""Host-side mirror of the reference's operator interface for the hot path.

Same names, argument meaning and error behaviour as `eigencuda::CudaPipeline`
(reference include/cudapipeline.hpp:20-42, src/cudapipeline.cc:5-32) and `eigencuda::CudaMatrix`
(reference include/cudamatrix.hpp:27-60, src/cudamatrix.cc:20-79), with `numpy` float64
Fortran-ordered arrays standing where the C++ API has `Eigen::MatrixXd` (column-major, leading
dimension = rows).  Everything goes through the C ABI (include/eigencuda_b200.h); nothing is
computed in Python.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, lib

Index = int  # `using Index = Eigen::Index`, include/cudamatrix.hpp:24


def _as_eigen(a):
    """Column-major float64 view/copy of `a`, like the storage of an Eigen::MatrixXd."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim != 2:
        raise ValueError("expected a 2-D matrix")
    return np.asfortranarray(a)


def count_available_gpus():
    """reference include/cudamatrix.hpp:25, src/cudamatrix.cc:14-18 (without the inverted ternary)."""
    return int(lib.ec_device_count())


class CudaMatrix:
    """reference include/cudamatrix.hpp:27-60.

    CudaMatrix(matrix, stream)        -> allocate + async H2D      (src/cudamatrix.cc:20-31)
    CudaMatrix(nrows, ncols, stream)  -> allocate, uninitialised   (src/cudamatrix.cc:33-37)
    Move-only in C++; here the handle is owned by the Python object and freed on `close`/GC.
    """

    def __init__(self, *args):
        self._h = C.c_void_p()
        if len(args) == 2:
            matrix, stream = args
            m = _as_eigen(matrix)
            check(lib.ec_matrix_create_from_host(_stream_arg(stream), m.ctypes.data, m.shape[0],
                                                 m.shape[1], C.byref(self._h)))
        elif len(args) == 3:
            nrows, ncols, stream = args
            check(lib.ec_matrix_alloc(_stream_arg(stream), int(nrows), int(ncols), C.byref(self._h)))
        else:
            raise TypeError("CudaMatrix(matrix, stream) or CudaMatrix(nrows, ncols, stream)")

    def rows(self):
        return int(lib.ec_matrix_rows(self._h))

    def cols(self):
        return int(lib.ec_matrix_cols(self._h))

    def size(self):
        return int(lib.ec_matrix_size(self._h))

    def data(self):
        """Device address of the first element (double*)."""
        return lib.ec_matrix_data(self._h) or 0

    def copy_to_gpu(self, A):
        """reference src/cudamatrix.cc:47-51: async H2D into the existing buffer; `A` is reusable
        as soon as the call returns."""
        a = _as_eigen(A)
        check(lib.ec_matrix_upload(self._h, a.ctypes.data, a.size))

    def to_eigen(self):
        """`operator Eigen::MatrixXd() const`, reference src/cudamatrix.cc:39-45: blocking D2H."""
        out = np.empty((self.rows(), self.cols()), dtype=np.float64, order="F")
        check(lib.ec_matrix_download(self._h, out.ctypes.data))
        return out

    def __array__(self, dtype=None, copy=None):
        a = self.to_eigen()
        return a if dtype is None else a.astype(dtype)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib.ec_matrix_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _stream_arg(stream):
    if isinstance(stream, CudaPipeline):
        return stream.get_stream()
    if isinstance(stream, C.c_void_p):
        return stream
    return C.c_void_p(int(stream) if stream else 0)


class CudaPipeline:
    """reference include/cudapipeline.hpp:20-42.  Owns one stream (no cuBLAS handle: the products
    are launches of this package's own sm_100a kernels).  Non-copyable, like the reference."""

    def __init__(self, device=-1, stream=None):
        self._h = C.c_void_p()
        if stream is None:
            check(lib.ec_pipeline_create(int(device), C.byref(self._h)))
        else:
            check(lib.ec_pipeline_create_on_stream(int(device), C.c_void_p(int(stream)), C.byref(self._h)))

    def __copy__(self):
        raise TypeError("CudaPipeline is not copyable (reference include/cudapipeline.hpp:28-29)")

    __deepcopy__ = __copy__

    def get_stream(self):
        """reference include/cudapipeline.hpp:34.  The cudaStream_t as an integer address."""
        return C.c_void_p(lib.ec_pipeline_stream(self._h))

    def device(self):
        return int(lib.ec_pipeline_device(self._h))

    def synchronize(self):
        check(lib.ec_pipeline_synchronize(self._h))

    def gemm(self, A, B, Cm):
        """reference src/cudapipeline.cc:17-32: C = A * B; raises (std::runtime_error there,
        EigenCudaError(RuntimeError) here) when A.cols() != B.rows(), before any device work."""
        check(lib.ec_gemm(self._h, A._h, B._h, Cm._h))

    # ---- additive API: the parameters the reference hard-codes ----
    def dgemm(self, transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, Cp, ldc):
        """Raw device-pointer form of the cublasDgemm call at reference src/cudapipeline.cc:29-31."""
        check(lib.ec_dgemm(self._h, int(transa), int(transb), m, n, k, float(alpha), int(A), lda,
                           int(B), ldb, float(beta), int(Cp), ldc))

    def dgemm_strided_batched(self, transa, transb, m, n, k, alpha, A, lda, sa, B, ldb, sb, beta,
                              Cp, ldc, sc, batch):
        check(lib.ec_dgemm_strided_batched(self._h, int(transa), int(transb), m, n, k, float(alpha),
                                           int(A), lda, sa, int(B), ldb, sb, float(beta), int(Cp),
                                           ldc, sc, batch))

    def tensor_stream(self, kind, F, tensor, results=None, devices=None):
        """The reference's streaming loop (README.md:62-66, src/tests/test_dot.cc:81-85) as one
        overlapped job.  `tensor` is a sequence of equally-shaped matrices (a Python list of
        Fortran-ordered float64 arrays == std::vector<Eigen::MatrixXd>), `F` the fixed matrix.
        Returns the list of results; `results` may supply preallocated outputs (float64,
        Fortran-contiguous, writeable, of the result shape -- anything else raises ValueError
        instead of being written through).  `devices`: None = this pipeline's device; a list of
        device indices (or "all") = the multi-GPU scheduler (ec_tensor_stream_run_multi): the
        tensor is sharded into contiguous ranges, F broadcast once with NCCL."""
        F = _as_eigen(F)
        ts = [_as_eigen(t) for t in tensor]
        n = len(ts)
        if n == 0:
            return []
        t_rows, t_cols = ts[0].shape
        if any(t.shape != (t_rows, t_cols) for t in ts):
            raise ValueError("all tensor elements must have the same shape")
        r_shape = result_shape(kind, F.shape, (t_rows, t_cols))
        if results is None:
            results = [np.empty(r_shape, dtype=np.float64, order="F") for _ in range(n)]
        else:
            results = list(results)     # keeps the arrays alive for the duration of the call
            if len(results) != n:
                raise ValueError(f"results has {len(results)} elements, tensor has {n}")
            for i, r in enumerate(results):
                if not isinstance(r, np.ndarray) or r.dtype != np.float64:
                    raise ValueError(f"results[{i}] must be a float64 numpy array")
                if r.shape != tuple(r_shape):
                    raise ValueError(f"results[{i}] has shape {r.shape}, the product has shape {tuple(r_shape)}")
                if not r.flags.f_contiguous:
                    raise ValueError(f"results[{i}] must be Fortran-contiguous (column-major, ld = rows)")
                if not r.flags.writeable:
                    raise ValueError(f"results[{i}] is read-only")
        in_ptrs = (C.c_void_p * n)(*[t.ctypes.data for t in ts])
        out_ptrs = (C.c_void_p * n)(*[r.ctypes.data for r in results])
        if devices is None:
            check(lib.ec_tensor_stream_run(self._h, int(kind), F.ctypes.data, F.shape[0], F.shape[1],
                                           in_ptrs, out_ptrs, n, t_rows, t_cols))
        else:
            dev_arr, ndev = _device_list(devices)
            check(lib.ec_tensor_stream_run_multi(self._h, dev_arr, ndev, int(kind), F.ctypes.data, F.shape[0],
                                                 F.shape[1], in_ptrs, out_ptrs, n, t_rows, t_cols))
        return results

    def right_multiply(self, tensor, F, devices=None):
        """results[i] = tensor[i] * F -- the order the reference's README/test use."""
        return self.tensor_stream(_lib.EC_STREAM_RIGHT, F, tensor, devices=devices)

    def left_multiply(self, F, tensor, transpose=False, devices=None):
        """results[i] = F * tensor[i]  (or F^T * tensor[i])."""
        return self.tensor_stream(_lib.EC_STREAM_LEFT_T if transpose else _lib.EC_STREAM_LEFT, F, tensor,
                                  devices=devices)

    def basis_transform(self, F, tensor, devices=None):
        """results[i] = F^T * tensor[i] * F  (BASELINE.json config 4; two products, the
        intermediate never leaves the device)."""
        return self.tensor_stream(_lib.EC_STREAM_BASIS, F, tensor, devices=devices)

    def tensor_stream_raw_multi(self, devices, kind, F_ptr, f_rows, f_cols, in_ptrs, out_ptrs, count,
                                t_rows, t_cols):
        """Pointer-level form of the multi-GPU job (ec_tensor_stream_run_multi)."""
        dev_arr, ndev = _device_list(devices)
        check(lib.ec_tensor_stream_run_multi(self._h, dev_arr, ndev, int(kind), int(F_ptr), f_rows, f_cols,
                                             in_ptrs, out_ptrs, count, t_rows, t_cols))

    def tensor_stream_raw(self, kind, F_ptr, f_rows, f_cols, in_ptrs, out_ptrs, count, t_rows, t_cols):
        """Pointer-level form: `F_ptr` may be a host or a device address (the latter after the
        multi-GPU broadcast), `in_ptrs`/`out_ptrs` ctypes arrays of host addresses."""
        check(lib.ec_tensor_stream_run(self._h, int(kind), int(F_ptr), f_rows, f_cols, in_ptrs,
                                       out_ptrs, count, t_rows, t_cols))

    def tensor_stream_device(self, kind, F_dev, f_rows, f_cols, tensor_dev, results_dev, count,
                             t_rows, t_cols):
        check(lib.ec_tensor_stream_run_device(self._h, int(kind), int(F_dev), f_rows, f_cols,
                                              int(tensor_dev), int(results_dev), count, t_rows, t_cols))

    def set_stream_config(self, chunk_matrices=0, ring_depth=0, staging_threads=0):
        check(lib.ec_pipeline_set_stream_config(self._h, chunk_matrices, ring_depth, staging_threads))

    def fill_uniform(self, seed, first_id, dst, count, batch=1):
        check(lib.ec_fill_uniform_batched(self._h, seed, first_id, int(dst), count, batch))

    def fill_uniform_block(self, seed, matrix_id, dst, rows, cols, ld_dst, row0, col0, global_ld):
        check(lib.ec_fill_uniform_block(self._h, seed, matrix_id, int(dst), rows, cols, ld_dst, row0,
                                        col0, global_ld))

    def set_sm_margin(self, sms):
        """Leave `sms` SMs free for concurrent kernels (NCCL panel broadcasts during SUMMA)."""
        check(lib.ec_pipeline_set_sm_margin(self._h, int(sms)))

    def set_schedule(self, stream_k=True, one_cta_per_tile=False):
        """Schedule policy of the GEMM kernels (see ec_pipeline_set_schedule in the C header)."""
        check(lib.ec_pipeline_set_schedule(self._h, int(bool(stream_k)), int(bool(one_cta_per_tile))))

    def set_backend(self, backend):
        """EC_BACKEND_FP64_DMMA (native, default) or EC_BACKEND_OZAKI_I8 (int8 tensor-core emulation
        with the looser bound stated in DESIGN.md)."""
        check(lib.ec_pipeline_set_backend(self._h, int(backend)))

    def get_backend(self):
        return int(lib.ec_pipeline_get_backend(self._h))

    def launch_count(self):
        return int(lib.ec_pipeline_launch_count(self._h))

    def contention_events(self):
        """How often the dispatcher found the GPU shared and switched to dynamic schedules."""
        return int(lib.ec_pipeline_contention_events(self._h))

    def last_kernel(self):
        return lib.ec_pipeline_last_kernel(self._h).decode()

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib.ec_pipeline_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceGrid:
    """P x Q grid of the box's GPUs for the 2D block-partitioned DGEMM (ec_grid; BASELINE.json
    configs[4]).  devices: list of device indices or "all"; P = Q = 0 picks the squarest grid."""

    def __init__(self, devices="all", P=0, Q=0):
        self._h = C.c_void_p()
        dev_arr, ndev = _device_list(devices)
        check(lib.ec_grid_create(dev_arr, ndev, int(P), int(Q), C.byref(self._h)))
        p, q = C.c_int(), C.c_int()
        check(lib.ec_grid_shape(self._h, C.byref(p), C.byref(q)))
        self.P, self.Q = p.value, q.value

    def matrix(self, rows, cols):
        return DistMatrix(self, rows, cols)

    def dgemm(self, A, B, Cm, panel=0):
        """Cm = A * B (SUMMA, NCCL panel broadcasts).  Returns device milliseconds, max over the GPUs.
        Raises like the reference's gemm on an inner-dimension mismatch, before any device work."""
        ms = C.c_double()
        check(lib.ec_dgemm_2d(self._h, A._h, B._h, Cm._h, int(panel), C.byref(ms)))
        return ms.value

    def set_backend(self, backend):
        check(lib.ec_grid_set_backend(self._h, int(backend)))

    def set_schedule(self, stream_k=False, one_cta_per_tile=True):
        check(lib.ec_grid_set_schedule(self._h, int(bool(stream_k)), int(bool(one_cta_per_tile))))

    def launch_count(self):
        return int(lib.ec_grid_launch_count(self._h))

    def broadcast_count(self):
        return int(lib.ec_grid_broadcast_count(self._h))

    def last_kernel(self):
        return lib.ec_grid_last_kernel(self._h).decode()

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib.ec_grid_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DistMatrix:
    """A rows x cols FP64 matrix cut into the grid's P x Q blocks, block (pr, pc) resident on
    device (pr, pc), column-major (ec_dist_matrix)."""

    def __init__(self, grid, rows, cols):
        self._h = C.c_void_p()
        self.grid, self.rows, self.cols = grid, int(rows), int(cols)
        check(lib.ec_dist_matrix_create(grid._h, self.rows, self.cols, C.byref(self._h)))

    def upload(self, a):
        a = _as_eigen(a)
        if a.shape != (self.rows, self.cols):
            raise ValueError(f"expected a {self.rows} x {self.cols} matrix")
        check(lib.ec_dist_matrix_upload(self._h, a.ctypes.data, max(a.shape[0], 1)))
        return self

    def download(self):
        out = np.empty((self.rows, self.cols), dtype=np.float64, order="F")
        check(lib.ec_dist_matrix_download(self._h, out.ctypes.data, max(self.rows, 1)))
        return out

    def fill_uniform(self, seed, matrix_id):
        check(lib.ec_dist_matrix_fill_uniform(self._h, seed, matrix_id))
        return self

    def get(self, i, j):
        v = C.c_double()
        check(lib.ec_dist_matrix_get(self._h, int(i), int(j), C.byref(v)))
        return v.value

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib.ec_dist_matrix_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def dgemm_2d(A, B, devices="all"):
    """C = A * B for host matrices through the 2D-partitioned path (ec_dgemm_2d_host)."""
    A, B = _as_eigen(A), _as_eigen(B)
    if A.shape[1] != B.shape[0]:
        raise _lib.EigenCudaError(_lib.EC_ERR_SHAPE, "Shape mismatch in Cublas gemm")
    out = np.empty((A.shape[0], B.shape[1]), dtype=np.float64, order="F")
    dev_arr, ndev = _device_list(devices)
    check(lib.ec_dgemm_2d_host(dev_arr, ndev, A.shape[0], B.shape[1], A.shape[1], A.ctypes.data, max(A.shape[0], 1),
                               B.ctypes.data, max(B.shape[0], 1), out.ctypes.data, max(A.shape[0], 1)))
    return out


def _device_list(devices):
    """("all" | iterable of device indices) -> (ctypes int array or None, ndev); ndev 0 = every device."""
    if isinstance(devices, str):
        if devices == "auto":
            return None, _lib.EC_DEVICES_AUTO
        if devices != "all":
            raise ValueError('devices must be a list of device indices, "all" or "auto"')
        return None, 0
    devs = [int(d) for d in devices]
    if not devs:
        raise ValueError("empty device list")
    return (C.c_int * len(devs))(*devs), len(devs)


def debug_hold_sms(device, stream, sms, microseconds):
    """Test hook (ec_debug_hold_sms): occupy `sms` SMs of `device` for `microseconds` on `stream`."""
    check(lib.ec_debug_hold_sms(int(device), _stream_arg(stream), int(sms), int(microseconds)))


def plan_shard(count, ndev, idx):
    """[lo, hi) of a `count`-element tensor that device number `idx` of `ndev` takes in the multi-GPU job."""
    lo, hi = C.c_int64(), C.c_int64()
    check(lib.ec_plan_shard(int(count), int(ndev), int(idx), C.byref(lo), C.byref(hi)))
    return lo.value, hi.value


def plan_chunks(count, chunk, ramp=True):
    """Chunk boundaries of a tensor-stream job (ec_plan_chunks): chunk c = [starts[c], starts[c+1])."""
    n = lib.ec_plan_chunks(int(count), int(chunk), int(bool(ramp)), None, 0)
    if n < 0:
        check(_lib.EC_ERR_ARG)
    buf = (C.c_int64 * n)()
    lib.ec_plan_chunks(int(count), int(chunk), int(bool(ramp)), buf, n)
    return list(buf)


def plan_panels(k, P, Q, panel=0):
    """k-panel boundaries of the 2D-partitioned DGEMM on a P x Q grid (ec_plan_panels); needs no device."""
    n = lib.ec_plan_panels(int(k), int(P), int(Q), int(panel), None, 0)
    if n < 0:
        check(_lib.EC_ERR_ARG)
    buf = (C.c_int64 * n)()
    lib.ec_plan_panels(int(k), int(P), int(Q), int(panel), buf, n)
    return list(buf)


def nccl_version():
    """(major, minor, patch) of the NCCL the multi-GPU entry points bind at run time (ec_nccl_version);
    needs no device.  Raises EigenCudaError (EC_ERR_UNSUPPORTED) with the loader's reason when it cannot
    be loaded."""
    v = C.c_int()
    check(lib.ec_nccl_version(C.byref(v)))
    return v.value // 10000, v.value // 100 % 100, v.value % 100


def pick_stream_devices(devices="all"):
    """The subset of `devices` the host streams through fastest, by measurement
    (ec_pick_stream_devices).  Returns (device list, GB/s each way of that set, GB/s of all)."""
    dev_arr, ndev = _device_list(devices)
    cap = ndev if ndev > 0 else max(1, count_available_gpus())
    out, nout, gbs = (C.c_int * cap)(), C.c_int(), (C.c_double * 2)()
    check(lib.ec_pick_stream_devices(dev_arr, max(ndev, 0), out, C.byref(nout), gbs))
    return list(out[: nout.value]), gbs[0], gbs[1]


def host_link_probe(devices="all", nbytes=256 << 20, reps=5):
    """Measured host-link ceiling in GB/s summed over `devices`: plain pinned cudaMemcpyAsync on all
    of them at once (ec_host_link_probe).  Returns {"h2d", "d2h", "bidir_each_way"}."""
    dev_arr, ndev = _device_list(devices)
    out = (C.c_double * 3)()
    check(lib.ec_host_link_probe(dev_arr, ndev, int(nbytes), int(reps), out))
    return {"h2d": out[0], "d2h": out[1], "bidir_each_way": out[2]}


def host_link_probe_buffers(devices, host_in_ptr, host_out_ptr, bytes_per_device, piece_bytes=64 << 20, reps=2):
    """The link ceiling measured on the caller's own pinned buffers (ec_host_link_probe_buffers): as
    much unique host memory as the job itself touches.  host_out is overwritten."""
    dev_arr, ndev = _device_list(devices)
    out = (C.c_double * 3)()
    check(lib.ec_host_link_probe_buffers(dev_arr, ndev, C.c_void_p(int(host_in_ptr)), C.c_void_p(int(host_out_ptr)),
                                         int(bytes_per_device), int(piece_bytes), int(reps), out))
    return {"h2d": out[0], "d2h": out[1], "bidir_each_way": out[2]}


def plan_dgemm(m, n, k, batch=1, sm_count=148):
    """(tile, stream_k) the dispatcher's cost model picks: tile 1 = 128x128, 2 = 64x64, 3 = 64x32,
    4 = 32x32.  Needs no device (ec_plan_dgemm)."""
    tile, sk = C.c_int(), C.c_int()
    check(lib.ec_plan_dgemm(int(sm_count), int(m), int(n), int(k), int(batch), C.byref(tile), C.byref(sk)))
    return tile.value, bool(sk.value)


def result_shape(kind, f_shape, t_shape):
    """Shape of one result of a tensor-stream job; raises like the reference's gemm on an inner
    dimension mismatch (reference src/cudapipeline.cc:26-28)."""
    (fr, fc), (tr, tc) = f_shape, t_shape
    if kind == _lib.EC_STREAM_RIGHT:
        ok, shape = tc == fr, (tr, fc)
    elif kind == _lib.EC_STREAM_LEFT:
        ok, shape = fc == tr, (fr, tc)
    elif kind == _lib.EC_STREAM_LEFT_T:
        ok, shape = fr == tr, (fc, tc)
    elif kind == _lib.EC_STREAM_BASIS:
        ok, shape = fr == tr and tc == fr, (fc, fc)
    else:
        raise ValueError(f"unknown tensor-stream kind {kind}")
    if not ok:
        raise _lib.EigenCudaError(_lib.EC_ERR_SHAPE, "Shape mismatch in Cublas gemm")
    return shape


class PinnedBuffer:
    """Page-locked host memory (what reference CHANGELOG.md:32 says 0.1.0 defaulted to), exposed
    as a numpy array; DMA'd directly by the tensor-stream engine without staging."""

    def __init__(self, shape, order="F"):
        self.shape = tuple(int(s) for s in np.atleast_1d(shape))
        n = int(np.prod(self.shape))
        self._ptr = C.c_void_p()
        check(lib.ec_host_alloc(max(n, 1) * 8, C.byref(self._ptr)))
        buf = (C.c_double * max(n, 1)).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=np.float64, count=n).reshape(self.shape, order=order)

    @property
    def ptr(self):
        return self._ptr.value

    def close(self):
        if self._ptr:
            self.array = None
            lib.ec_host_free(self._ptr)
            self._ptr = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class host_pin:
    """Pins numpy arrays the caller already owns, in place, while the context is open
    (ec_host_register): transfers from / to them are plain DMA and
    skip the staging copy.  The arrays must stay alive and must not be reallocated meanwhile."""

    def __init__(self, *arrays):
        self._arrays = [a for a in arrays if a.size]
        self._done = []

    def __enter__(self):
        for a in self._arrays:
            if not (a.flags.f_contiguous or a.flags.c_contiguous):
                raise ValueError("host_pin needs contiguous arrays")
            check(lib.ec_host_register(C.c_void_p(a.ctypes.data), a.nbytes))
            self._done.append(a)
        return self

    def __exit__(self, *exc):
        for a in self._done:
            lib.ec_host_unregister(C.c_void_p(a.ctypes.data))
        self._done = []
        return False
