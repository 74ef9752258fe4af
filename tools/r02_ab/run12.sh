mkdir -p gpurun_out/f12
B="--steps 20 --warmup 5 --no-cpu --no-peaks --no-e2e --no-2d --no-sweep"
cp eigencuda_b200/lib/libeigencuda_b200.so /tmp/keep.so
for v in e0 e1 e3 e0 e1 e3; do
  cp tools/bin/lib_$v.so eigencuda_b200/lib/libeigencuda_b200.so
  timeout 200 python bench.py $B --no-extras >> gpurun_out/f12/all_$v.jsonl 2>> gpurun_out/f12/bench_$v.err
done
cp /tmp/keep.so eigencuda_b200/lib/libeigencuda_b200.so
