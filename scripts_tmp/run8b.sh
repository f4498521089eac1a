run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29800+RANDOM%100)) bench.py --gpus 8 --steps 20 --warmup 5 "$@"; }
PVCR_DP_SKIP=0,1,2,3 run > gpurun_out/r2j_skipall.json 2>/dev/null
PVCR_DP_SKIP=0 run > gpurun_out/r2j_skip0.json 2>/dev/null
PVCR_DP_SKIP=1 run > gpurun_out/r2j_skip1.json 2>/dev/null
PVCR_DP_SKIP=2 run > gpurun_out/r2j_skip2.json 2>/dev/null
PVCR_DP_SKIP=3 run > gpurun_out/r2j_skip3.json 2>/dev/null
PVCR_DP_SKIP=2,3 run > gpurun_out/r2j_skip23.json 2>/dev/null
PVCR_DP_SKIP=0,1 run > gpurun_out/r2j_skip01.json 2>/dev/null
echo done
