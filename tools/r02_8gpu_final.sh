#!/bin/bash
# Final 8-GPU session of round 2: the scheduler's device pick, smoke(), the multi-GPU tests that
# exercise it, and the bench line at N = 8.
set -x
mkdir -p gpurun_out
python -c "
import eigencuda_b200 as ec, time, json
t = time.time(); picked, g1, g2 = ec.pick_stream_devices('all')
print(json.dumps({'picked': picked, 'gbs_each_way_picked': g1, 'gbs_each_way_all': g2, 'seconds': time.time() - t}))
" | tee gpurun_out/r02_device_pick_8gpu.txt
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 400 python -m pytest tests -m gpu -x -q -k "multi_gpu_readme or pick_stream or multi_gpu_cpp or multi_gpu_counts" 2>&1 | tail -4 | tee gpurun_out/r02_pytest_8gpu_final.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err
tail -3 gpurun_out/r02_bench_8gpu.err
