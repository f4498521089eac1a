TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2h_tests.log
B="bench.py --steps 20 --warmup 5"
timeout 300 python $B --no-eager --no-greedy --no-optimizer --no-cpu-baseline > gpurun_out/r2h_b1.json 2>/dev/null
timeout 300 $TR --master-port 29710 $B --gpus 2 > gpurun_out/r2h_b2.json 2> /dev/null
timeout 300 python $B --workload cfg3 > gpurun_out/r2h_b1_cfg3.json 2>/dev/null
PVCR_NO_TMA_XCHG=1 timeout 300 python $B --workload cfg3 > gpurun_out/r2h_b1_cfg3_notma.json 2>/dev/null
timeout 300 $TR --master-port 29711 $B --gpus 2 --workload cfg3 > gpurun_out/r2h_b2_cfg3.json 2> /dev/null
PVCR_PHASE_DEC_BWD=1 timeout 300 python tests/gpu_probe_phases.py > gpurun_out/r2h_ph_dec_bwd.txt 2>&1
cat gpurun_out/r2h_tests.log | tail -4
