#!/bin/bash
# Run ON the GPU box (through gpurun): bench lines + ncu launch list + ncu full captures -> gpurun_out/.
# usage: tools/capture_profiles.sh <tag>
set -u
TAG="${1:-rX}"
OUT=gpurun_out
python bench.py > $OUT/bench_${TAG}_n1.json 2> $OUT/bench_${TAG}_n1.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_${TAG}_reference.json 2> $OUT/bench_${TAG}_reference.err
CMD="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
$CMD > $OUT/plain_${TAG}.log 2>&1 || exit 2
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/launches_${TAG}.csv $CMD > $OUT/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_density_binned -s 10 -c 2 -o $OUT/prof_binned_${TAG} $CMD > $OUT/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_density_dense32 -c 1 -o $OUT/prof_dense32_${TAG} $CMD > $OUT/ncu_d.log 2>&1
ncu --set full --clock-control none -k regex:"k_minmax|k_scale_count|k_scan|k_place|k_tile_classify|k_set_totals|k_js" -s 21 -c 7 -o $OUT/prof_small_${TAG} $CMD > $OUT/ncu_s.log 2>&1
for f in ncu_b ncu_d ncu_s; do tail -n 1 $OUT/$f.log; done
# fine-grid path (128x128 cells, cell moments): config C5 at 3 M + 3 M points, R = 1024
C5CMD="python tools/c5_probe.py"
C5_N=3000000 $C5CMD > $OUT/plain_c5_${TAG}.log 2>&1 || exit 3
C5_N=3000000 ncu --set full --clock-control none --import-source on -k regex:"k_density_binned|k_cell_moments|k_scan_wide" -c 3 -o $OUT/prof_c5_${TAG} $C5CMD > $OUT/ncu_c5.log 2>&1
tail -n 1 $OUT/ncu_c5.log
