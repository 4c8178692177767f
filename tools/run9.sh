python -m pytest tests/test_band_gpu.py tests/test_binned32_gpu.py -m gpu -q > gpurun_out/t9.log 2>&1; tail -4 gpurun_out/t9.log
PROBE_LOOPS=0 python tools/fp32_probe.py 2>&1 | grep -A1 "^binned32"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 tools/mgpu_check.py > gpurun_out/m9.log 2>&1; grep -E "^C5|steps|older|single|ok" gpurun_out/m9.log
