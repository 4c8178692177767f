python -m pytest tests/test_nccl_gpu.py -m gpu -q > gpurun_out/t8.log 2>&1; tail -5 gpurun_out/t8.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 tools/mgpu_check.py > gpurun_out/m8.log 2>&1; grep -v "^\[W\|NCCL version" gpurun_out/m8.log | tail -25
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29656 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/b8_n2.json 2> gpurun_out/b8_n2.err; echo "bench rc=$?"; tail -3 gpurun_out/b8_n2.err
