run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29800+RANDOM%100)) bench.py --gpus 8 --steps 20 --warmup 5 "$@"; }
run > gpurun_out/r2m_b8_merge.json 2>/dev/null
PVCR_DP_MERGE_TAIL=0 run > gpurun_out/r2m_b8_nomerge.json 2>/dev/null
echo done
