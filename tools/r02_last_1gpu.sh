# Last check of the round on one B200 (6 GPU-minutes left): GPU suite, smoke, the driver's bench line.
mkdir -p gpurun_out/last
timeout 150 python -m pytest tests -m gpu -x -q > gpurun_out/last/pytest_1gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/last/pytest_1gpu.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/last/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/last/smoke.log
timeout 150 python bench.py --steps 20 --warmup 5 > gpurun_out/last/bench_1gpu.json 2> gpurun_out/last/bench_1gpu.err; echo "bench rc=$?" >> gpurun_out/last/bench_1gpu.err
tail -2 gpurun_out/last/pytest_1gpu.log; tail -2 gpurun_out/last/smoke.log; tail -1 gpurun_out/last/bench_1gpu.err
