set -x
mkdir -p gpurun_out/f2
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/f2/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f2/pytest.log
B="--steps 10 --warmup 3 --no-cpu --no-peaks --no-e2e --no-extras --no-2d"
timeout 200 python bench.py $B > gpurun_out/f2/bench_head1.json 2> gpurun_out/f2/bench_head1.err
cp eigencuda_b200/lib/libeigencuda_b200.so /tmp/lib_head1.so
cp tools/bin/lib_head0.so eigencuda_b200/lib/libeigencuda_b200.so
timeout 200 python bench.py $B > gpurun_out/f2/bench_head0.json 2> gpurun_out/f2/bench_head0.err
cp /tmp/lib_head1.so eigencuda_b200/lib/libeigencuda_b200.so
timeout 200 python bench.py $B --no-sweep > gpurun_out/f2/bench_head1b.json 2> gpurun_out/f2/bench_head1b.err
tail -3 gpurun_out/f2/pytest.log
