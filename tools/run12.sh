python -m pytest tests -m gpu -x -q > gpurun_out/t12.log 2>&1; tail -4 gpurun_out/t12.log
PROBE_LOOPS=0 python tools/fp32_probe.py 2>&1 | grep -E "us/eval|rel gap"
python tools/c5_probe.py 2>&1 | tail -5
