set -x
mkdir -p gpurun_out/f3
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/f3/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f3/pytest.log
B="--steps 10 --warmup 3 --no-cpu --no-peaks --no-e2e --no-extras --no-2d"
timeout 200 python bench.py $B > gpurun_out/f3/bench_coords.json 2> gpurun_out/f3/bench_coords.err
timeout 300 ncu --set full --import-source on --clock-control none -k regex:dgemm_tma -s 3 -c 1 -f -o gpurun_out/f3/prof_dgemm python bench.py --steps 2 --warmup 3 --no-cpu --no-peaks --no-e2e --no-sweep --no-extras --no-2d > gpurun_out/f3/ncu.log 2>&1
tail -3 gpurun_out/f3/pytest.log
