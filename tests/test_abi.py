"""The C-ABI library loads and exports every symbol include/eigencuda_b200.h declares; error
paths that must not need a GPU behave (not gpu: no compute calls here)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "eigencuda_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ec_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("ec_pipeline_create", "ec_pipeline_destroy", "ec_pipeline_stream", "ec_matrix_alloc",
                 "ec_matrix_create_from_host", "ec_matrix_upload", "ec_matrix_download", "ec_gemm",
                 "ec_dgemm", "ec_dgemm_strided_batched", "ec_tensor_stream_run", "ec_device_count"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    import eigencuda_b200._lib as L
    lib = C.CDLL(L.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    # and the Python binding table covers the header exactly
    assert sorted(L.SIGNATURES) == declared_symbols()


def test_no_cublas_and_no_oracle_linked():
    import eigencuda_b200._lib as L
    out = subprocess.run(["ldd", L.LIB_PATH], capture_output=True, text=True).stdout
    assert "cublas" not in out.lower(), out
    assert "oracle" not in out.lower(), out
    assert "libcuda.so" not in out, "must not hard-link the driver (loads without a GPU)"


def test_product_sources_do_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "eigencuda_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                for line in text.splitlines():
                    s = line.strip()
                    if s.startswith(("import ", "from ", "#include")):
                        assert "oracle" not in s, (f, s)
    for f in ("cudamatrix.hpp", "cudapipeline.hpp", "eigencuda_b200.h"):
        for line in open(os.path.join(ROOT, "include", f)):
            if line.strip().startswith("#include"):
                assert "oracle" not in line and "cublas" not in line, (f, line)


def test_argument_errors_need_no_gpu():
    import eigencuda_b200 as ec
    from eigencuda_b200._lib import EC_ERR_ARG, lib
    assert lib.ec_version().startswith(b"eigencuda_b200")
    assert lib.ec_gemm(None, None, None, None) == EC_ERR_ARG
    assert b"null" in lib.ec_last_error_string()
    assert lib.ec_dgemm(None, 0, 0, 1, 1, 1, 1.0, None, 1, None, 1, 0.0, None, 1) == EC_ERR_ARG
    assert lib.ec_matrix_upload(None, None, 0) == EC_ERR_ARG
    assert lib.ec_pipeline_destroy(None) == 0 and lib.ec_matrix_free(None) == 0
    assert ec.count_available_gpus() >= 0


def test_fails_loudly_without_a_gpu():
    import eigencuda_b200 as ec
    if ec.count_available_gpus() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(ec.EigenCudaError):
        ec.CudaPipeline()


def test_missing_library_raises(tmp_path):
    import eigencuda_b200._lib as L
    with pytest.raises(RuntimeError, match="no CPU fallback|not built"):
        L.load(str(tmp_path / "nope.so"))


def test_cpp_headers_compile_against_the_abi(tmp_path):
    """include/cudamatrix.hpp + cudapipeline.hpp: caller code written like the reference's tests
    compiles unchanged (Eigen stand-in, since Eigen is not in the image)."""
    exe = tmp_path / "test_dot_compat"
    cmd = ["g++", "-std=c++14", "-O1", "-Wall", "-Werror", f"-I{ROOT}/include",
           f"-I{ROOT}/oracle/eigen_standin", "-I/usr/local/cuda/include",
           f"{ROOT}/tests/cpp/test_dot_compat.cc", "-o", str(exe),
           f"-L{ROOT}/eigencuda_b200/lib", "-leigencuda_b200",
           f"-Wl,-rpath,{ROOT}/eigencuda_b200/lib"]
    if not os.path.exists("/usr/local/cuda/include/cuda_runtime_api.h"):
        pytest.skip("CUDA headers not present")
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def test_cmake_package_builds_the_reference_test_and_is_findable(tmp_path):
    """the CMake package.  Configure this project against the in-tree library
    (EIGENCUDA_PREBUILT_LIB: no nvcc run), build `unit_test_dot` -- the reference's own test target name
    (reference src/tests/CMakeLists.txt:3-17) -- install, then configure and build a separate caller
    project that does find_package(eigencuda) and links the plain target name `eigencuda` as the
    reference's callers do (reference src/CMakeLists.txt:2)."""
    import shutil
    import eigencuda_b200._lib as L
    if not shutil.which("cmake") or not os.path.exists("/usr/local/cuda/include/cuda_runtime_api.h"):
        pytest.skip("cmake or the CUDA headers are not present")
    build, prefix, cbuild = tmp_path / "build", tmp_path / "prefix", tmp_path / "consumer_build"
    run = lambda cmd: subprocess.run(cmd, capture_output=True, text=True)
    res = run(["cmake", "-S", ROOT, "-B", str(build), "-DENABLE_TESTING=ON", f"-DEIGENCUDA_PREBUILT_LIB={L.LIB_PATH}",
               "-DCUDAToolkit_ROOT=/usr/local/cuda"])
    assert res.returncode == 0, res.stdout + res.stderr
    res = run(["cmake", "--build", str(build), "--target", "unit_test_dot", "unit_test_multi_gpu", "-j", "4"])
    assert res.returncode == 0, res.stdout + res.stderr
    assert (build / "unit_test_dot").exists() and (build / "unit_test_multi_gpu").exists()
    res = run(["cmake", "--install", str(build), "--prefix", str(prefix)])
    assert res.returncode == 0, res.stdout + res.stderr
    assert (prefix / "lib" / "cmake" / "eigencuda" / "eigencudaConfig.cmake").exists()
    assert (prefix / "include" / "cudapipeline.hpp").exists() and (prefix / "lib" / "libeigencuda_b200.so").exists()
    res = run(["cmake", "-S", os.path.join(ROOT, "tests", "cmake_consumer"), "-B", str(cbuild),
               f"-DCMAKE_PREFIX_PATH={prefix}", "-DCUDAToolkit_ROOT=/usr/local/cuda",
               f"-DEIGENCUDA_TEST_SOURCE={ROOT}/tests/cpp/test_dot_compat.cc",
               f"-DEIGEN_STANDIN_DIR={ROOT}/oracle/eigen_standin"])
    assert res.returncode == 0, res.stdout + res.stderr
    res = run(["cmake", "--build", str(cbuild)])
    assert res.returncode == 0, res.stdout + res.stderr
    out = run(["ldd", str(cbuild / "consumer")]).stdout
    assert "libeigencuda_b200.so" in out and "cublas" not in out.lower()


def _nccl_probe(env_extra):
    """ec_nccl_version in a fresh process (the binding is made once per process)."""
    code = ("import eigencuda_b200 as ec\n"
            "try:\n"
            "    print('ok', *ec.nccl_version())\n"
            "except ec.EigenCudaError as e:\n"
            "    print('err', e.code, e)\n")
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), **env_extra)
    res = subprocess.run([os.sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stderr
    return res.stdout.strip().splitlines()[-1]


def test_nccl_is_bound_at_run_time_and_reports_a_missing_library():
    """The multi-GPU entry points dlopen NCCL (csrc/multi_gpu.inl: nccl_api); a library that cannot be
    loaded must come back as EC_ERR_UNSUPPORTED with the loader's reason, not as a crash."""
    from eigencuda_b200._lib import EC_ERR_UNSUPPORTED, lib
    assert lib.ec_nccl_version(None) != 0          # null out pointer: argument error
    line = _nccl_probe({"EC_NCCL_LIB": "/nonexistent/libnccl.so.2"})
    assert line.startswith(f"err {EC_ERR_UNSUPPORTED} "), line
    assert "/nonexistent/libnccl.so.2" in line
    # a library without the NCCL symbols is refused, naming a missing symbol
    line = _nccl_probe({"EC_NCCL_LIB": "libm.so.6"})
    assert line.startswith(f"err {EC_ERR_UNSUPPORTED} ") and "lacks nccl" in line, line
    # the default candidates: present in this image through torch's bundled copy when torch is imported first
    import glob
    import importlib.util
    spec = importlib.util.find_spec("nvidia.nccl")
    cands = glob.glob(os.path.join(list(spec.submodule_search_locations)[0], "lib", "libnccl.so.2")) if spec else []
    if not cands:
        pytest.skip("no libnccl.so.2 to bind in this environment")
    line = _nccl_probe({"EC_NCCL_LIB": cands[0]})
    assert line.startswith("ok 2 "), line


def test_every_entry_point_survives_null_and_zero_arguments():
    """No exported function may crash the caller's process on a null handle, a null pointer or zero sizes:
    each one is called that way (in a child process, so that a crash is a test failure and not the end of the
    run) and has to come back -- with an error code where the header promises one."""
    code = r'''
import ctypes as C
import eigencuda_b200._lib as L
for name, (res, args) in L.SIGNATURES.items():
    vals = [0.0 if a is C.c_double else (None if (a is C.c_void_p or issubclass(a, C._Pointer)) else 0) for a in args]
    r = getattr(L.lib, name)(*vals)
    print(name, r, flush=True)
print("survived")
'''
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    res = subprocess.run([os.sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=ROOT)
    assert res.returncode == 0 and res.stdout.strip().endswith("survived"), res.stdout[-400:] + res.stderr[-400:]
    got = dict(line.split(" ", 1) for line in res.stdout.strip().splitlines()[:-1])
    import eigencuda_b200._lib as L
    for name in ("ec_gemm", "ec_dgemm", "ec_matrix_upload", "ec_tensor_stream_run", "ec_tensor_stream_run_multi",
                 "ec_dgemm_2d", "ec_pipeline_create", "ec_grid_create", "ec_pick_stream_devices", "ec_nccl_version"):
        assert int(got[name]) != L.EC_OK, name
    for name in ("ec_pipeline_destroy", "ec_matrix_free", "ec_grid_destroy", "ec_dist_matrix_destroy", "ec_host_free"):
        assert int(got[name]) == L.EC_OK, name        # destroying nothing is not an error (unique_ptr deleters do it)


def test_device_needing_entry_points_report_the_missing_gpu():
    """No CPU fallback anywhere: on a machine without a GPU every entry point that needs one comes back with an
    error code and a message, valid arguments notwithstanding."""
    import numpy as np
    import eigencuda_b200 as ec
    from eigencuda_b200._lib import EC_OK, lib
    if ec.count_available_gpus() > 0:
        pytest.skip("a GPU is present")
    a = np.zeros((4, 4), order="F")
    h, out3, devs, ndev = C.c_void_p(), (C.c_double * 3)(), (C.c_int * 8)(), C.c_int()
    calls = {
        "ec_matrix_alloc": lambda: lib.ec_matrix_alloc(None, 4, 4, C.byref(h)),
        "ec_matrix_create_from_host": lambda: lib.ec_matrix_create_from_host(None, a.ctypes.data, 4, 4, C.byref(h)),
        "ec_host_alloc": lambda: lib.ec_host_alloc(1024, C.byref(h)),
        "ec_host_register": lambda: lib.ec_host_register(a.ctypes.data, a.nbytes),
        "ec_host_link_probe": lambda: lib.ec_host_link_probe(None, 0, 1 << 20, 1, out3),
        "ec_pick_stream_devices": lambda: lib.ec_pick_stream_devices(None, 0, devs, C.byref(ndev), None),
        "ec_grid_create": lambda: lib.ec_grid_create(None, 0, 0, 0, C.byref(h)),
        "ec_dgemm_2d_host": lambda: lib.ec_dgemm_2d_host(None, 0, 4, 4, 4, a.ctypes.data, 4, a.ctypes.data, 4, a.ctypes.data, 4),
    }
    for name, call in calls.items():
        assert call() != EC_OK, name
        assert lib.ec_last_error_string(), name
        assert not h.value, name
    with pytest.raises(ec.EigenCudaError):
        ec.dgemm_2d(a, a)
    with pytest.raises(ec.EigenCudaError):
        ec.host_link_probe()
