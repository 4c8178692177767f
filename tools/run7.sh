PROBE_LOOPS=0 python tools/fp32_probe.py > gpurun_out/p7.log 2>&1; cat gpurun_out/p7.log
CMD="python bench.py --algo binned32 --steps 10 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/plain7.json 2> gpurun_out/plain7.err && ncu --set full --clock-control none --import-source on -k regex:"k_scale_count|k_place|k_tile_classify|k_density_binned32|k_minmax|k_js_tiles" -s 60 -c 8 -o gpurun_out/prof_small_r2b $CMD > gpurun_out/ncu7.log 2>&1; tail -2 gpurun_out/ncu7.log | cut -c1-200
