TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_s2vtatt.py tests/test_gpu_rationale.py tests/test_gpu_data_parallel.py tests/test_optim.py -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2k_tests.log
B="bench.py --steps 20 --warmup 5"
timeout 300 python $B --no-eager --no-greedy --no-optimizer --no-cpu-baseline > gpurun_out/r2k_b1.json 2>/dev/null
timeout 300 $TR --master-port 29710 $B --gpus 2 > gpurun_out/r2k_b2.json 2> gpurun_out/r2k_b2.err
PVCR_DP_MERGE_TAIL=0 timeout 300 $TR --master-port 29711 $B --gpus 2 > gpurun_out/r2k_b2_nomerge.json 2> /dev/null
timeout 300 $TR --master-port 29712 $B --gpus 2 --workload cfg3 > gpurun_out/r2k_b2_cfg3.json 2> /dev/null
tail -4 gpurun_out/r2k_tests.log
