mkdir -p gpurun_out/f11
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/f11/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f11/pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu --no-2d --no-peaks > gpurun_out/f11/bench.json 2> gpurun_out/f11/bench.err; echo "bench rc=$?" >> gpurun_out/f11/bench.err
tail -2 gpurun_out/f11/pytest.log; tail -1 gpurun_out/f11/bench.err
