# usage: bash tools/run_exp.sh VARIANT...   (timing experiments, see tools/build_exp.sh)
for v in "$@"; do
  PCOE_LIB=$PWD/gpurun_exp/libpcoe_$v.so timeout 200 python -m pytest tests/test_sa_gpu.py -m gpu -q -k bf16 2>&1 | tail -1
  PCOE_LIB=$PWD/gpurun_exp/libpcoe_$v.so timeout 120 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/exp_$v.json 2> gpurun_out/exp_$v.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/exp_$v.json').read().strip().splitlines()[-1])
print('$v', round(d['value']), ' '.join(k+'='+str(round(x['ms_per_step']*1000,1)) for k,x in d['kernels'].items() if k.startswith('sa1_') or k.startswith('sa2_')))
"
done
