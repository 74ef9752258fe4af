set -x
mkdir -p gpurun_out/f5
B="--steps 20 --warmup 5 --no-cpu --no-peaks --no-e2e --no-extras --no-2d --no-sweep"
cp eigencuda_b200/lib/libeigencuda_b200.so /tmp/lib_v2.so
for v in v2 v1 v2 v1 v2 v1; do
  if [ $v = v2 ]; then cp /tmp/lib_v2.so eigencuda_b200/lib/libeigencuda_b200.so; else cp tools/bin/lib_$v.so eigencuda_b200/lib/libeigencuda_b200.so; fi
  timeout 200 python bench.py $B >> gpurun_out/f5/all_$v.jsonl 2>> gpurun_out/f5/bench_$v.err
done
cp /tmp/lib_v2.so eigencuda_b200/lib/libeigencuda_b200.so
