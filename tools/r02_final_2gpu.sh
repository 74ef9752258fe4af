# Round-2 final check on two B200s of one box: the whole GPU suite (multi-GPU tests included) and the bench line at N = 2.
mkdir -p gpurun_out/final
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/final/pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final/pytest_2gpu.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/final/bench_2gpu.json 2> gpurun_out/final/bench_2gpu.err; echo "bench rc=$?" >> gpurun_out/final/bench_2gpu.err
tail -2 gpurun_out/final/pytest_2gpu.log; tail -1 gpurun_out/final/bench_2gpu.err
