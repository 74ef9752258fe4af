mkdir -p gpurun_out/f9
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/f9/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f9/pytest.log
timeout 200 python tools/sk_probe.py 1792,2304,2560,2816,3072,3328,3584,4096,6144 > gpurun_out/f9/sk_probe.txt 2>&1
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-peaks --no-e2e --no-extras --no-2d > gpurun_out/f9/bench.json 2> gpurun_out/f9/bench.err
tail -3 gpurun_out/f9/pytest.log
