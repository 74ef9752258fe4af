set -x
mkdir -p gpurun_out/f1
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/f1/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f1/pytest.log
B="--steps 10 --warmup 3 --no-cpu --no-peaks --no-e2e --no-extras --no-2d"
EC_FUSE_EPILOGUE=1 timeout 200 python bench.py $B > gpurun_out/f1/bench_fuse1.json 2> gpurun_out/f1/bench_fuse1.err
EC_FUSE_EPILOGUE=0 timeout 200 python bench.py $B > gpurun_out/f1/bench_fuse0.json 2> gpurun_out/f1/bench_fuse0.err
tail -3 gpurun_out/f1/pytest.log
