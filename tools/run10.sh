PROBE_LOOPS=0 python tools/fp32_probe.py 2>&1 | grep -E "us/eval"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 tools/mgpu_check.py > gpurun_out/m10.log 2>&1; grep -E "^C5|steps|kernel ms|single|older|ok|Error|error" gpurun_out/m10.log | cut -c1-600
