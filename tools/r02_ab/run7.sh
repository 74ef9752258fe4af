set -x
mkdir -p gpurun_out/f7
B="--steps 20 --warmup 5 --no-cpu --no-peaks --no-e2e --no-extras --no-2d --no-sweep"
cp eigencuda_b200/lib/libeigencuda_b200.so /tmp/lib_p5.so
cp tools/bin/lib_p0.so tools/bin/lib_p6.so /tmp/
for v in p5 p0 p6 p5 p0 p6; do
  cp /tmp/lib_$v.so eigencuda_b200/lib/libeigencuda_b200.so
  timeout 200 python bench.py $B >> gpurun_out/f7/all_$v.jsonl 2>> gpurun_out/f7/bench_$v.err
done
cp /tmp/lib_p5.so eigencuda_b200/lib/libeigencuda_b200.so
