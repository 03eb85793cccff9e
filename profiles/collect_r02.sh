#!/bin/bash
# Collects the round-2 single-GPU evidence on the GPU box (run through gpurun from the repo root):
#   bench lines of every workload, the ncu launch list of the default bench command, and `ncu --set full` captures of the
#   GNN kernels, the Hartree pair kernels and the fused functionals kernel.  Outputs land in gpurun_out/r02/.
set -x
O=gpurun_out/r02; mkdir -p $O
python bench.py --steps 20 --warmup 3 > $O/bench_1gpu.json 2> $O/bench_1gpu.err
for wl in h2 lih c6h6 lih1d; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_$wl.json 2> $O/bench_$wl.err
done
timeout 600 python bench.py --steps 5 --warmup 3 --recompute --no-cpu-baseline > $O/bench_recompute.json 2> $O/bench_recompute.err
timeout 600 python bench.py --steps 5 --warmup 3 --recompute --ffma --no-cpu-baseline > $O/bench_recompute_ffma.json 2> $O/bench_recompute_ffma.err
timeout 600 python bench.py --steps 5 --warmup 3 --ffma --no-cpu-baseline > $O/bench_ffma.json 2> $O/bench_ffma.err
# launch list (cold-cache, serialised: shares only) -- after the plain command above exited 0
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ncu_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-blocks > $O/ncu_launches.log 2>&1
# full captures
timeout 900 ncu --set full --import-source on --clock-control none -k regex:gnn_ -s 6 -c 2 -f -o $O/prof_gnn \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-blocks > $O/ncu_gnn.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:hartree_pairs_kernel -c 6 -f -o $O/prof_pairs \
  python tests/bench_pairs.py 262144 > $O/ncu_pairs.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:functionals_kernel -s 3 -c 1 -f -o $O/prof_functionals \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-blocks > $O/ncu_functionals.log 2>&1
ls -la $O
