python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 tools/mgpu_check.py > gpurun_out/m33.log 2>&1; echo rc=$?; grep -v "^W\|^\[W" gpurun_out/m33.log | grep "C5 n\|ok\|rror\|steps" | head -12
