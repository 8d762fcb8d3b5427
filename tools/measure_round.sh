#!/bin/bash
# One measurement pass on a GPU box: bench line (+ other workloads), ncu launch list of the same command, one
# --set full capture of the three kernels of the step and of the splat kernels.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 600 gpurun_out/bench_final.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 5 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"warp_t_fwd_lean|warp_t_bwd_tile|flow_stats" -s 3 -c 3 -f \
    -o gpurun_out/prof_final_step python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"splat" -s 12 -c 6 -f \
    -o gpurun_out/prof_final_splat python tools/splat_probe.py 16 > /dev/null 2>&1
ls -la gpurun_out/prof_final_*.ncu-rep
