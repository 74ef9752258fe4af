#!/bin/bash
# One 8-GPU session: link ceilings for 1/2/4/8 GPUs, the 2D-partitioned DGEMM at n = 65536, the
# multi-GPU tests, the bench line at N = 8 (and N = 4).  Writes into gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8; free -g | head -2; nproc
python tools/link_probe_subsets.py | tee gpurun_out/r02_link_probe_8gpu.txt
python tools/dgemm_2d_run.py --gpus 8 2>&1 | tail -1 | tee -a gpurun_out/r02_dgemm2d_8gpu.txt
python tools/dgemm_2d_run.py --gpus 8 --panel 2048 2>&1 | tail -1 | tee -a gpurun_out/r02_dgemm2d_8gpu.txt
EC_2D_NCCL_CTAS=8 python tools/dgemm_2d_run.py --gpus 8 2>&1 | tail -1 | tee -a gpurun_out/r02_dgemm2d_8gpu.txt
python tools/dgemm_2d_run.py --gpus 8 --backend ozaki 2>&1 | tail -1 | tee -a gpurun_out/r02_dgemm2d_8gpu.txt
python tools/dgemm_2d_run.py --gpus 4 2>&1 | tail -1 | tee -a gpurun_out/r02_dgemm2d_8gpu.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "dgemm_2d or multi_gpu_cpp or multi_gpu_readme" 2>&1 | tail -5 | tee gpurun_out/r02_pytest_8gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err
tail -3 gpurun_out/r02_bench_8gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 10 --warmup 3 --no-2d > gpurun_out/r02_bench_4gpu.json 2> gpurun_out/r02_bench_4gpu.err
tail -3 gpurun_out/r02_bench_4gpu.err
python tools/engine_sweep.py --gpus 8 --count 1024 --chunks 16,32,64 --rings 3 2>&1 | tee gpurun_out/r02_engine_sweep_8gpu.txt
