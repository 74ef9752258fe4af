mkdir -p gpurun_out/f13
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/f13/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f13/pytest.log
B="--steps 20 --warmup 5 --no-cpu --no-peaks --no-e2e --no-2d --no-sweep --no-extras"
cp eigencuda_b200/lib/libeigencuda_b200.so /tmp/keep.so
for v in keep e0 keep e0; do
  if [ $v = keep ]; then cp /tmp/keep.so eigencuda_b200/lib/libeigencuda_b200.so; else cp tools/bin/lib_$v.so eigencuda_b200/lib/libeigencuda_b200.so; fi
  timeout 200 python bench.py $B >> gpurun_out/f13/all_$v.jsonl 2>> gpurun_out/f13/bench_$v.err
done
cp /tmp/keep.so eigencuda_b200/lib/libeigencuda_b200.so
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-peaks --no-2d --no-sweep --no-e2e > gpurun_out/f13/bench_extras.json 2> gpurun_out/f13/bench_extras.err
tail -3 gpurun_out/f13/pytest.log
