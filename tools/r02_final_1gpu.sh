# Round-2 final evidence on one B200: GPU suite, smoke, the driver's bench line, the ncu launch list and one full capture
# of the headline launch (each ncu pass only after the same command exited 0 without the profiler).
mkdir -p gpurun_out/final
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/final/pytest_1gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final/pytest_1gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final/smoke.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/final/bench_1gpu.json 2> gpurun_out/final/bench_1gpu.err; echo "bench rc=$?" >> gpurun_out/final/bench_1gpu.err
Q="--steps 2 --warmup 3 --no-cpu --no-peaks --no-e2e --no-sweep --no-extras --no-2d"
timeout 100 python bench.py $Q > gpurun_out/final/plain.log 2>&1 && \
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final/ncu_launches.csv python bench.py $Q > gpurun_out/final/ncu1.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:dgemm_tma -s 3 -c 1 -f -o gpurun_out/final/prof_dgemm python bench.py $Q > gpurun_out/final/ncu2.log 2>&1
tail -2 gpurun_out/final/pytest_1gpu.log; tail -2 gpurun_out/final/smoke.log; tail -1 gpurun_out/final/bench_1gpu.err
