python -m pytest tests/test_binned32_gpu.py tests/test_band_gpu.py -m gpu -x -q > gpurun_out/t15.log 2>&1; tail -3 gpurun_out/t15.log
python tools/fp32_probe.py 2>&1 | grep -E "loop peak|us/eval|rel gap|per eval"
