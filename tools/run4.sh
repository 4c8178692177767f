python -m pytest tests/test_binned32_gpu.py tests/test_band_gpu.py -m gpu -q > gpurun_out/t4.log 2>&1; tail -12 gpurun_out/t4.log
for a in binned32 binned; do python tools/overlap_probe.py $a; KDEJS_DENS_BLOCKS=5 python tools/overlap_probe.py $a; KDEJS_DENS_BLOCKS=4 python tools/overlap_probe.py $a; done > gpurun_out/p4.log 2>&1; cat gpurun_out/p4.log
CMD="python bench.py --algo binned32 --steps 10 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/plain4.json 2> gpurun_out/plain4.err && ncu --set full --clock-control none --import-source on -k regex:k_density_binned32 -s 10 -c 1 -o gpurun_out/prof_b32_r2a $CMD > gpurun_out/ncu4.log 2>&1; tail -3 gpurun_out/ncu4.log
