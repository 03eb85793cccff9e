#!/bin/bash
# Round-2 evidence of the 1-D flow (run through gpurun from the repo root): bench lines with and without the activation
# checkpoint, the ncu launch list of one step and `ncu --set full` captures of the layer GEMM and the weight-gradient GEMM.
set -x
O=gpurun_out/r02_1d; mkdir -p $O
timeout 600 python bench.py --workload lih1d --steps 10 --warmup 3 --no-extra-blocks > $O/bench_lih1d.json 2> $O/bench_lih1d.err
timeout 600 python bench.py --workload lih1d --steps 10 --warmup 3 --no-cpu-baseline --no-extra-blocks --recompute > $O/bench_lih1d_recompute.json 2> $O/bench_lih1d_recompute.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/ncu_launches_lih1d.csv \
  python bench.py --workload lih1d --steps 1 --warmup 3 --no-cpu-baseline --no-extra-blocks > $O/ncu_launches.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:mlp_layer_gemm_tc -s 700 -c 3 -f -o $O/prof_mlp_gemm \
  python bench.py --workload lih1d --steps 1 --warmup 3 --no-cpu-baseline --no-extra-blocks > $O/ncu_gemm.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:mlp_dw_gemm_ts -s 40 -c 1 -f -o $O/prof_mlp_dw \
  python bench.py --workload lih1d --steps 1 --warmup 3 --no-cpu-baseline --no-extra-blocks > $O/ncu_dw.log 2>&1
ls -la $O
