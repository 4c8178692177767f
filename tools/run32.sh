python -m pytest tests -m gpu -x -q 2>&1 | tail -2
PROBE_LOOPS=0 python tools/fp32_probe.py 2>&1 | grep -E " R 256|rel gap"
PROBE_R=100 PROBE_LOOPS=0 python tools/fp32_probe.py 2>&1 | grep -E "binned32 R 100"
C5_N=3000000 python tools/c5_probe.py 2>&1 | grep -E "pinned|density|classify" | tail -3
