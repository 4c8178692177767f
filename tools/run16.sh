python -m pytest tests/test_band_gpu.py tests/test_nccl_gpu.py -m gpu -x -q > gpurun_out/t16.log 2>&1; tail -3 gpurun_out/t16.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 tools/mgpu_check.py > gpurun_out/m16.log 2>&1; grep -E "^C5|steps|kernel ms|single|older|ok|Error|error" gpurun_out/m16.log | cut -c1-420
