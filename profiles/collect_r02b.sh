#!/bin/bash
# Re-collects, with the final tree of round 2, the single-GPU evidence that the depth-generic kernels and the in-run traffic
# measurement touch (run through gpurun from the repo root; outputs in gpurun_out/r02b/): default bench line, recompute mode,
# GNN depths 3 / 1, the launch list of the depth-3 step and `ncu --set full` captures of the depth-generic kernels.
set -x
O=gpurun_out/r02b; mkdir -p $O
python bench.py --steps 20 --warmup 3 > $O/bench_1gpu.json 2> $O/bench_1gpu.err
timeout 600 python bench.py --steps 5 --warmup 3 --recompute --no-cpu-baseline > $O/bench_recompute.json 2> $O/bench_recompute.err
timeout 600 python bench.py --steps 10 --warmup 3 --nn 3 --no-cpu-baseline > $O/bench_h2o_nn3.json 2> $O/bench_h2o_nn3.err
timeout 600 python bench.py --steps 10 --warmup 3 --nn 1 --no-cpu-baseline > $O/bench_h2o_nn1.json 2> $O/bench_h2o_nn1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/ncu_launches_nn3.csv \
  python bench.py --nn 3 --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches_nn3.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gnn_deep_ -s 4 -c 3 -f -o $O/prof_gnn_deep \
  python bench.py --nn 3 --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_gnn_deep.log 2>&1
ls -la $O
