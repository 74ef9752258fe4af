set -x
mkdir -p gpurun_out/f6
B="--steps 20 --warmup 5 --no-cpu --no-peaks --no-e2e --no-extras --no-2d --no-sweep"
cp eigencuda_b200/lib/libeigencuda_b200.so /tmp/lib_r2.so
cp tools/bin/lib_r4.so tools/bin/lib_r8.so /tmp/
for v in r2 r4 r8 r2 r4 r8; do
  cp /tmp/lib_$v.so eigencuda_b200/lib/libeigencuda_b200.so
  timeout 200 python bench.py $B >> gpurun_out/f6/all_$v.jsonl 2>> gpurun_out/f6/bench_$v.err
done
cp /tmp/lib_r2.so eigencuda_b200/lib/libeigencuda_b200.so
