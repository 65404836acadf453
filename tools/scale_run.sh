#!/usr/bin/env bash
# usage (on a multi-GPU box): bash tools/scale_run.sh TAG "N:config ..."   e.g.  bash tools/scale_run.sh r2 "1:c3 8:c3 8:c2"
# One bench.py run per entry (torchrun for N > 1, as the driver launches it); JSON lines go to gpurun_out/TAG_<cfg>_g<N>.json
tag=$1; shift
for e in $1; do
  n=${e%%:*}; cfg=${e##*:}
  out=gpurun_out/${tag}_${cfg}_g${n}.json
  if [ "$n" = 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 --config $cfg --no-cpu-baseline --no-modes-leg --no-sampling-leg > $out 2> ${out%.json}.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 50 --warmup 5 --config $cfg --no-cpu-baseline --no-modes-leg --no-sampling-leg > $out 2> ${out%.json}.err
  fi
  python - <<PY
import json
try:
    d = json.loads([l for l in open("$out").read().strip().splitlines() if l.startswith("{")][-1])
    print("$cfg", "N=$n", "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "launches/step", d.get("gpu_launches_per_step"))
except Exception as ex:
    print("$cfg N=$n FAILED", ex)
PY
done
