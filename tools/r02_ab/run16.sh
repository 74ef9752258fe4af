mkdir -p gpurun_out/f16
timeout 200 python -m pytest tests -m gpu -x -q -k "fused or skinny or square or layouts or ragged or edge" > gpurun_out/f16/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f16/pytest.log
PROBE_MIN_KBLOCKS=2 timeout 100 python tools/fuse_k_probe.py > gpurun_out/f16/fuse_k_probe_default.txt 2>&1
tail -2 gpurun_out/f16/pytest.log
