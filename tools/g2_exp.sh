# N=2 NCCL configuration sweep (diagnostic): step time with one all-reduce after backward (--no-overlap) and with the
# default two overlapped buckets
run() { # label, then VAR=VALUE..., then -- bench args
  lbl=$1; shift
  envs=(); while [[ $# -gt 0 && "$1" != "--" ]]; do envs+=("$1"); shift; done; shift || true
  env "${envs[@]}" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29520 + RANDOM % 100)) bench.py --gpus 2 --steps 100 --warmup 5 --timed-only "$@" 2>/dev/null | tail -1 | cut -c1-64 | sed "s/^/$lbl /"
}
run default_noov X=1 -- --no-overlap
run simple_noov NCCL_PROTO=Simple -- --no-overlap
run ll128_noov NCCL_PROTO=LL128 -- --no-overlap
run ll_noov NCCL_PROTO=LL -- --no-overlap
run nvls0_noov NCCL_NVLS_ENABLE=0 -- --no-overlap
run minch16_noov NCCL_MIN_NCHANNELS=16 -- --no-overlap
run ll128_ov NCCL_PROTO=LL128 --
run minch16_ov NCCL_MIN_NCHANNELS=16 --
