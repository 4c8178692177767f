python -m pytest tests/test_binned32_gpu.py tests/test_band_gpu.py -m gpu -q > gpurun_out/t5.log 2>&1; tail -8 gpurun_out/t5.log
PROBE_LOOPS=0 python tools/fp32_probe.py > gpurun_out/p5.log 2>&1; cat gpurun_out/p5.log
