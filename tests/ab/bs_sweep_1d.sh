for bs in 256 512 768 1024; do
  timeout 200 python bench.py --workload lih1d --bs $bs --steps 5 --warmup 3 --no-cpu-baseline --no-extra-blocks > gpurun_out/lih1d_bs$bs.json 2>gpurun_out/lih1d_bs$bs.err
  python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d['ms_per_step'],3), round(d['value']), {k:round(v,3) for k,v in d['roofline']['kernel_ms'].items() if k.startswith('mlp')})" gpurun_out/lih1d_bs$bs.json
done
