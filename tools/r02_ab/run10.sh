mkdir -p gpurun_out/f10
B="--steps 20 --warmup 5 --no-cpu --no-peaks --no-e2e --no-extras --no-2d --no-sweep"
cp eigencuda_b200/lib/libeigencuda_b200.so /tmp/lib_u1.so
cp tools/bin/lib_u0.so /tmp/
timeout 200 python -m pytest tests -m gpu -x -q -k "fused or dynamic or stream_k or slot_release or square or layouts" > gpurun_out/f10/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f10/pytest.log
for v in u1 u0 u1 u0; do
  cp /tmp/lib_$v.so eigencuda_b200/lib/libeigencuda_b200.so
  timeout 200 python bench.py $B >> gpurun_out/f10/all_$v.jsonl 2>> gpurun_out/f10/bench_$v.err
done
cp /tmp/lib_u1.so eigencuda_b200/lib/libeigencuda_b200.so
tail -3 gpurun_out/f10/pytest.log
