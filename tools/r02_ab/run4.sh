set -x
mkdir -p gpurun_out/f4
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/f4/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f4/pytest.log
B="--steps 10 --warmup 3 --no-cpu --no-peaks --no-e2e --no-extras --no-2d"
cp eigencuda_b200/lib/libeigencuda_b200.so /tmp/lib_v2.so
for v in v2 v0 v1 v2; do
  if [ $v = v2 ]; then cp /tmp/lib_v2.so eigencuda_b200/lib/libeigencuda_b200.so; else cp tools/bin/lib_$v.so eigencuda_b200/lib/libeigencuda_b200.so; fi
  timeout 200 python bench.py $B > gpurun_out/f4/bench_$v.json 2> gpurun_out/f4/bench_$v.err
  cat gpurun_out/f4/bench_$v.json >> gpurun_out/f4/all_$v.jsonl
done
cp /tmp/lib_v2.so eigencuda_b200/lib/libeigencuda_b200.so
python tools/small_probe.py > gpurun_out/f4/small_probe_v2.txt 2>&1
tail -3 gpurun_out/f4/pytest.log
