B="bench.py --steps 30 --warmup 5 --no-eager --no-greedy --no-optimizer --no-cpu-baseline"
for i in 1 2; do
timeout 300 python $B > gpurun_out/r2l_a$i.json 2>/dev/null
PVCR_NO_EMB_FIRST=1 timeout 300 python $B > gpurun_out/r2l_b$i.json 2>/dev/null
PVCR_NO_EMB_EARLY_ZERO=1 timeout 300 python $B > gpurun_out/r2l_c$i.json 2>/dev/null
PVCR_NO_EMB_FIRST=1 PVCR_NO_EMB_EARLY_ZERO=1 timeout 300 python $B > gpurun_out/r2l_d$i.json 2>/dev/null
done
echo done
