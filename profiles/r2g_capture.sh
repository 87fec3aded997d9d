#!/bin/bash
# Final round-2 evidence run (gpurun, one GPU): bench line, reference arm, ncu launch list, ncu --set full of the step kernels.
set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2g_bench_ref.json 2> gpurun_out/r2g_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g_launches_raw.csv \
    python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/r2g_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"big2s|hidden_fwd12|hidden_bwd2|head_bwd|prep_operands|gram16|loss_dF16|weight_stats|fwd_plan" \
    -c 14 -o gpurun_out/r2g_prof -f python bench.py --steps 1 --warmup 3 --points 65536 --no-configs --no-cpu-baseline > gpurun_out/r2g_ncu_full.log 2>&1
python profiles/summarize.py gpurun_out/r2g_prof.ncu-rep gpurun_out/r2g_launches_raw.csv r2g > gpurun_out/r2g_summarize.log 2>&1
cp profiles/r2g_ncu_full.csv profiles/r2g_launches.csv gpurun_out/ 2>/dev/null
ls -la gpurun_out | tail -8
