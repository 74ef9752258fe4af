"""ctypes binding of the C ABI in include/eigencuda_b200.h.

The shared object is built in-tree (eigencuda_b200/lib/libeigencuda_b200.so) by
`make -C eigencuda_b200/csrc` or `__graft_entry__.build()`.  There is no fallback: when the
library is missing, importing this module raises, and when no B200 is present every compute
entry point returns an error that the wrappers turn into RuntimeError.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# EIGENCUDA_B200_LIB points the binding at another build of the same library (A/B measurements)
LIB_PATH = os.environ.get("EIGENCUDA_B200_LIB") or os.path.join(_HERE, "lib", "libeigencuda_b200.so")

EC_OK, EC_ERR_SHAPE, EC_ERR_NOMEM, EC_ERR_COPY, EC_ERR_CUDA, EC_ERR_ARG, EC_ERR_UNSUPPORTED = range(7)
EC_OP_N, EC_OP_T = 0, 1
EC_BACKEND_FP64_DMMA, EC_BACKEND_OZAKI_I8 = 0, 1
EC_STREAM_RIGHT, EC_STREAM_LEFT, EC_STREAM_LEFT_T, EC_STREAM_BASIS = 0, 1, 2, 3
EC_DEVICES_AUTO = -1

_vp, _i64, _int, _dbl, _u64 = C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_uint64
_dp = C.c_void_p  # device / host double* passed as raw addresses

# name -> (restype, argtypes); mirrors include/eigencuda_b200.h line by line.
SIGNATURES = {
    "ec_version": (C.c_char_p, []),
    "ec_last_error_string": (C.c_char_p, []),
    "ec_device_count": (_i64, []),
    "ec_pipeline_create": (_int, [_int, C.POINTER(_vp)]),
    "ec_pipeline_create_on_stream": (_int, [_int, _vp, C.POINTER(_vp)]),
    "ec_pipeline_destroy": (_int, [_vp]),
    "ec_pipeline_stream": (_vp, [_vp]),
    "ec_pipeline_stream_ref": (_vp, [_vp]),
    "ec_pipeline_device": (_int, [_vp]),
    "ec_pipeline_synchronize": (_int, [_vp]),
    "ec_pipeline_set_backend": (_int, [_vp, _int]),
    "ec_pipeline_get_backend": (_int, [_vp]),
    "ec_matrix_alloc": (_int, [_vp, _i64, _i64, C.POINTER(_vp)]),
    "ec_matrix_create_from_host": (_int, [_vp, _dp, _i64, _i64, C.POINTER(_vp)]),
    "ec_matrix_free": (_int, [_vp]),
    "ec_matrix_upload": (_int, [_vp, _dp, _i64]),
    "ec_matrix_download": (_int, [_vp, _dp]),
    "ec_matrix_rows": (_i64, [_vp]),
    "ec_matrix_cols": (_i64, [_vp]),
    "ec_matrix_size": (_i64, [_vp]),
    "ec_matrix_data": (_vp, [_vp]),
    "ec_gemm": (_int, [_vp, _vp, _vp, _vp]),
    "ec_dgemm": (_int, [_vp, _int, _int, _i64, _i64, _i64, _dbl, _dp, _i64, _dp, _i64, _dbl, _dp, _i64]),
    "ec_dgemm_strided_batched": (_int, [_vp, _int, _int, _i64, _i64, _i64, _dbl, _dp, _i64, _i64,
                                        _dp, _i64, _i64, _dbl, _dp, _i64, _i64, _i64]),
    "ec_tensor_stream_run": (_int, [_vp, _int, _dp, _i64, _i64, C.POINTER(_vp), C.POINTER(_vp),
                                    _i64, _i64, _i64]),
    "ec_tensor_stream_run_multi": (_int, [_vp, C.POINTER(_int), _int, _int, _dp, _i64, _i64, C.POINTER(_vp),
                                          C.POINTER(_vp), _i64, _i64, _i64]),
    "ec_grid_create": (_int, [C.POINTER(_int), _int, _int, _int, C.POINTER(_vp)]),
    "ec_grid_destroy": (_int, [_vp]),
    "ec_grid_shape": (_int, [_vp, C.POINTER(_int), C.POINTER(_int)]),
    "ec_grid_set_backend": (_int, [_vp, _int]),
    "ec_grid_set_schedule": (_int, [_vp, _int, _int]),
    "ec_grid_launch_count": (_i64, [_vp]),
    "ec_grid_broadcast_count": (_i64, [_vp]),
    "ec_grid_last_kernel": (C.c_char_p, [_vp]),
    "ec_dist_matrix_create": (_int, [_vp, _i64, _i64, C.POINTER(_vp)]),
    "ec_dist_matrix_destroy": (_int, [_vp]),
    "ec_dist_matrix_upload": (_int, [_vp, _dp, _i64]),
    "ec_dist_matrix_download": (_int, [_vp, _dp, _i64]),
    "ec_dist_matrix_fill_uniform": (_int, [_vp, _u64, _u64]),
    "ec_dist_matrix_get": (_int, [_vp, _i64, _i64, C.POINTER(_dbl)]),
    "ec_dgemm_2d": (_int, [_vp, _vp, _vp, _vp, _i64, C.POINTER(_dbl)]),
    "ec_dgemm_2d_host": (_int, [C.POINTER(_int), _int, _i64, _i64, _i64, _dp, _i64, _dp, _i64, _dp, _i64]),
    "ec_plan_chunks": (_i64, [_i64, _i64, _int, C.POINTER(_i64), _i64]),
    "ec_nccl_version": (_int, [C.POINTER(_int)]),
    "ec_pick_stream_devices": (_int, [C.POINTER(_int), _int, C.POINTER(_int), C.POINTER(_int), C.POINTER(_dbl)]),
    "ec_plan_panels": (_i64, [_i64, _int, _int, _i64, C.POINTER(_i64), _i64]),
    "ec_plan_shard": (_int, [_i64, _int, _int, C.POINTER(_i64), C.POINTER(_i64)]),
    "ec_host_link_probe": (_int, [C.POINTER(_int), _int, C.c_size_t, _int, C.POINTER(_dbl)]),
    "ec_host_link_probe_buffers": (_int, [C.POINTER(_int), _int, _vp, _vp, C.c_size_t, C.c_size_t, _int, C.POINTER(_dbl)]),
    "ec_tensor_stream_run_device": (_int, [_vp, _int, _dp, _i64, _i64, _dp, _dp, _i64, _i64, _i64]),
    "ec_pipeline_set_sm_margin": (_int, [_vp, _int]),
    "ec_pipeline_set_schedule": (_int, [_vp, _int, _int]),
    "ec_pipeline_set_stream_config": (_int, [_vp, _int, _int, _int]),
    "ec_host_alloc": (_int, [C.c_size_t, C.POINTER(_vp)]),
    "ec_host_free": (_int, [_vp]),
    "ec_plan_dgemm": (_int, [_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.POINTER(_int), C.POINTER(_int)]),
    "ec_host_register": (_int, [_vp, C.c_size_t]),
    "ec_host_unregister": (_int, [_vp]),
    "ec_fill_uniform": (_int, [_vp, _u64, _u64, _dp, _i64]),
    "ec_fill_uniform_batched": (_int, [_vp, _u64, _u64, _dp, _i64, _i64]),
    "ec_fill_uniform_block": (_int, [_vp, _u64, _u64, _dp, _i64, _i64, _i64, _i64, _i64, _i64]),
    "ec_pipeline_contention_events": (_i64, [_vp]),
    "ec_debug_hold_sms": (_int, [_int, _vp, _int, _i64]),
    "ec_debug_fastdiv": (C.c_uint32, [C.c_uint32, C.c_uint32]),
    "ec_pipeline_launch_count": (_i64, [_vp]),
    "ec_pipeline_last_kernel": (C.c_char_p, [_vp]),
}


def load(path=LIB_PATH):
    if not os.path.exists(path):
        raise RuntimeError(
            f"eigencuda_b200: CUDA extension not built ({path} missing). "
            "Run `make -C eigencuda_b200/csrc` or `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        if os.environ.get("EIGENCUDA_B200_LIB") and not hasattr(lib, name):
            continue             # an older build under A/B measurement: bind what it has
        fn = getattr(lib, name)  # AttributeError here = ABI drift, fail loudly
        fn.restype = res
        fn.argtypes = args
    return lib


lib = load()


class EigenCudaError(RuntimeError):
    """Python face of the std::runtime_error the reference throws (src/cudapipeline.cc:26-28,
    src/cudamatrix.cc:28-30,63-79)."""

    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


def check(code):
    if code != EC_OK:
        raise EigenCudaError(code, lib.ec_last_error_string().decode("utf-8", "replace"))
