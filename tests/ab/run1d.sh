timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mlp1d or 1d" 2>&1 | tail -5
show() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], d['ms_per_step'], d['roofline']['kernel_ms'])" $1; }
timeout 200 python bench.py --workload lih1d --steps 10 --warmup 3 --no-cpu-baseline --no-extra-blocks > gpurun_out/lih1d_act.json 2>gpurun_out/lih1d_act.err; show gpurun_out/lih1d_act.json
timeout 200 python bench.py --workload lih1d --steps 10 --warmup 3 --no-cpu-baseline --no-extra-blocks --recompute > gpurun_out/lih1d_rc.json 2>gpurun_out/lih1d_rc.err; show gpurun_out/lih1d_rc.json
OFDFT_B200_LIB=$PWD/ofdft_nflows_b200/variants/lib_sg1.so timeout 200 python bench.py --workload lih1d --steps 10 --warmup 3 --no-cpu-baseline --no-extra-blocks > gpurun_out/lih1d_sg1.json 2>gpurun_out/lih1d_sg1.err; show gpurun_out/lih1d_sg1.json
tail -3 gpurun_out/lih1d_act.err

