#!/bin/bash
# Run ON the GPU box (through gpurun): bench lines + ncu launch list + ncu full captures -> gpurun_out/.
# usage: tools/capture_profiles.sh <tag>          (every ncu pass runs only after the same command exited 0 without ncu)
set -u
TAG="${1:-rX}"
OUT=gpurun_out
python bench.py > $OUT/bench_${TAG}_n1.json 2> $OUT/bench_${TAG}_n1.err || exit 1
python bench.py --algo binned --steps 200 --no-cpu-baseline > $OUT/bench_${TAG}_n1_binned.json 2> $OUT/bench_${TAG}_n1_binned.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_${TAG}_reference.json 2> $OUT/bench_${TAG}_reference.err
python tools/fp32_probe.py > $OUT/fp32_probe_${TAG}.log 2>&1
CMD="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > $OUT/plain_${TAG}.log 2>&1 || exit 2
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_${TAG}.csv $CMD > $OUT/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_density_binned32 -s 10 -c 2 -o $OUT/prof_binned32_${TAG} $CMD > $OUT/ncu_b32.log 2>&1
ncu --set full --clock-control none -k regex:'k_minmax|k_scale_count|k_scan|k_place|k_tile_classify|k_set_totals|k_js|^k_density_binned$' -s 27 -c 9 -o $OUT/prof_small_${TAG} $CMD > $OUT/ncu_s.log 2>&1
CMD64="python bench.py --algo binned --steps 10 --warmup 3 --no-cpu-baseline --no-extras"
$CMD64 > $OUT/plain64_${TAG}.log 2>&1 || exit 3
ncu --set full --clock-control none --import-source on -k 'regex:^k_density_binned$' -s 10 -c 2 -o $OUT/prof_binned_${TAG} $CMD64 > $OUT/ncu_b64.log 2>&1
# fine-grid path (config C5's geometry at 3 M + 3 M points): moment cells, wide scan, heavy tiles in parts
C5_N=3000000 python tools/c5_probe.py > $OUT/plain_c5_${TAG}.log 2>&1 || exit 4
C5_N=3000000 ncu --set full --clock-control none --import-source on -k 'regex:^k_density_binned$|k_cell_moments|k_scan_wide|k_scale_count' -s 4 -c 4 -o $OUT/prof_c5_${TAG} python tools/c5_probe.py > $OUT/ncu_c5.log 2>&1
python tools/band_sort_probe.py 1250000 > $OUT/band_sort_probe_${TAG}.log 2>&1
tools/ubench/match_probe > $OUT/match_probe_${TAG}.log 2>&1
# what a wave costs whatever its size, and what that does to the host batch pipeline (first / middle / last wave sizes)
WAVES=1,2,4,8,12,16,24,32,64 python tools/wave_probe.py > $OUT/wave_probe_${TAG}.log 2>&1
python tools/tail_probe.py "3D=0" "3D=1" "3D=1;KDEJS_TAIL_WAVE=6" "3D=1;KDEJS_WAVE_PLAN=1,2,4,7,8,7" "3D=1;KDEJS_WAVE_PLAN=2,6,12,12,12,10,10" "3D=1;KDEJS_FIRST_WAVE=64;KDEJS_HOST_WAVE=64" > $OUT/tail_probe_${TAG}.log 2>&1
for f in ncu_b32 ncu_s ncu_b64 ncu_c5; do tail -n 1 $OUT/$f.log; done
# tables on disk -> arg-min: host parser vs the parser on the device
python tools/ingest_probe.py 2048 > $OUT/ingest_probe_${TAG}.log 2>&1
# the device table parser alone: a call over 64 tables held as text, then its three kernels under ncu
TEXT_WAVES=16 python tools/tsv_probe.py 64 > $OUT/tsv_probe_${TAG}.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_tsv -s 6 -c 3 -o $OUT/prof_tsv_${TAG} python tools/tsv_probe.py 64 > $OUT/ncu_tsv.log 2>&1
tail -n 1 $OUT/ncu_tsv.log
