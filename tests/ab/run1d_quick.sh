show() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['kernel_ms'].items() if k.startswith('mlp')}, d['energy_terms_last_step'][:3])" $1; }
for v in "$@"; do
  if [ "$v" = base ]; then unset OFDFT_B200_LIB; else export OFDFT_B200_LIB=$PWD/ofdft_nflows_b200/variants/lib_$v.so; fi
  timeout 200 python bench.py --workload lih1d --steps 10 --warmup 3 --no-cpu-baseline --no-extra-blocks > gpurun_out/lih1d_$v.json 2>gpurun_out/lih1d_$v.err; show gpurun_out/lih1d_$v.json
done
