mkdir -p gpurun_out/f8
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/f8/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f8/pytest.log
timeout 200 python tools/sk_probe.py > gpurun_out/f8/sk_probe.txt 2>&1
tail -3 gpurun_out/f8/pytest.log
