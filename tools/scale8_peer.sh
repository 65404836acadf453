# 8-GPU comparison of the two gradient exchanges on one box (run with gpurun --gpus 8)
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/peer_bench.py 2>&1 | grep -E "^world" > gpurun_out/peer_bench_g8.txt
cat gpurun_out/peer_bench_g8.txt
PCOE_EXCHANGE=peer bash tools/scale_run.sh s8_peer "8:c2 8:c3"
PCOE_EXCHANGE=nccl bash tools/scale_run.sh s8_nccl "8:c2 8:c3"
PCOE_EXCHANGE=peer PCOE_PEER_MULTICAST=0 bash tools/scale_run.sh s8_peer_nomc "8:c2"
bash tools/scale_run.sh s8 "1:c2 1:c3"
PCOE_EXCHANGE=peer timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 tools/dp_overlap_check.py 2>&1 | grep -E "^\[peer\]|dp_overlap"
