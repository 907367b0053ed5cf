#!/bin/bash
# Round-2 captures on one B200 (run under gpurun from the repository root; outputs land in gpurun_out/ and are
# turned into the tracked summaries by profiles/extract_ncu.py and profiles/sass_summary.py afterwards).
# ncu passes never produce bench values; the bench lines come from the plain runs at the end.
set -u
mkdir -p gpurun_out
Q="--quick --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 6 --warmup 3 $Q > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:measure_kernel -s 4 -c 1 -f -o gpurun_out/r2_measure_f32 \
    python bench.py --steps 3 --warmup 3 $Q > gpurun_out/ncu_f32.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:measure_kernel -s 4 -c 1 -f -o gpurun_out/r2_measure_f64arith \
    python bench.py --steps 3 --warmup 3 $Q --arith f64 > gpurun_out/ncu_f64arith.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:measure_kernel -s 4 -c 1 -f -o gpurun_out/r2_measure_f64 \
    python bench.py --steps 3 --warmup 3 $Q --arith f64 --dtype f64 > gpurun_out/ncu_f64.log 2>&1
ncu --set full --clock-control none -k regex:"copy_blocks|motion_kernel|weight_scan|thresholds|resample_plan|assign_kernel|free_list_fused" \
    -s 21 -c 7 -f -o gpurun_out/r2_others python bench.py --steps 3 --warmup 3 $Q > gpurun_out/ncu_others.log 2>&1
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err
ls -la gpurun_out | tail -20
