set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2n_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2n_ref.json 2>/dev/null
for w in "" DEC_BWD GRU; do
  if [ -z "$w" ]; then timeout 200 python tests/gpu_probe_phases.py > gpurun_out/r2n_ph_dec_fwd.txt 2>&1;
  else env PVCR_PHASE_$w=1 timeout 200 python tests/gpu_probe_phases.py > gpurun_out/r2n_ph_$w.txt 2>&1; fi
done
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-greedy --no-eager --no-optimizer"
$CMD > gpurun_out/r2n_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r2n_ncu1.log 2>&1
$CMD > gpurun_out/r2n_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"dec_persist_bwd|dec_persist_fwd|gru_persist_fwd|gru_persist_bwd|EpiCeFwd|EpiCeBwd" -s 12 -c 6 -o gpurun_out/r02_prof $CMD > gpurun_out/r2n_ncu2.log 2>&1
tail -3 gpurun_out/r2n_tests.log
