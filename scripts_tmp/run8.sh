run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29800+RANDOM%100)) bench.py --gpus $n --steps 20 --warmup 5 "$@"; }
run 8 > gpurun_out/r2i_b8.json 2> gpurun_out/r2i_b8.err
run 8 --nccl-ctas 24 > gpurun_out/r2i_b8_c24.json 2>/dev/null
run 4 > gpurun_out/r2i_b4.json 2>/dev/null
run 8 --workload cfg3 > gpurun_out/r2i_b8_cfg3.json 2>/dev/null
run 4 --workload cfg3 > gpurun_out/r2i_b4_cfg3.json 2>/dev/null
python bench.py --impl reference --gpus 1 --steps 3 --warmup 3 > gpurun_out/r2i_ref1.json 2>/dev/null
run 8 --impl reference --steps 3 --warmup 3 > gpurun_out/r2i_ref8.json 2>/dev/null
echo done
