mkdir -p gpurun_out/f14
B="--steps 20 --warmup 5 --no-cpu --no-peaks --no-e2e --no-2d --no-extras"
cp eigencuda_b200/lib/libeigencuda_b200.so /tmp/keep.so
for v in e0 keep e0 keep; do
  if [ $v = keep ]; then cp /tmp/keep.so eigencuda_b200/lib/libeigencuda_b200.so; else cp tools/bin/lib_$v.so eigencuda_b200/lib/libeigencuda_b200.so; fi
  timeout 200 python bench.py $B >> gpurun_out/f14/all_$v.jsonl 2>> gpurun_out/f14/bench_$v.err
done
cp /tmp/keep.so eigencuda_b200/lib/libeigencuda_b200.so
timeout 200 python -m pytest tests -m gpu -x -q -k "fused or dynamic or stream_k or square or layouts or readme or schedule" > gpurun_out/f14/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f14/pytest.log
tail -2 gpurun_out/f14/pytest.log
