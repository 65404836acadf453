run() { tag=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 50 --warmup 5 --config c2 --no-cpu-baseline --no-modes-leg --no-sampling-leg "$@" 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('$tag', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value']))"; }
run skip --skip-allreduce
run peer --exchange peer
run peer_noov --exchange peer --no-overlap
run nccl --exchange nccl
run nccl_noov --exchange nccl --no-overlap
run skip2 --skip-allreduce
