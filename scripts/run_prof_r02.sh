# Round-2 evidence run on one B200 (everything lands in gpurun_out/; copy what is kept into profiles/).
set -x
timeout 120 scripts/micro/mma_rate > gpurun_out/r02_mma_rate_micro.txt 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02_final.log 2>&1
tail -c 400 gpurun_out/bench_r02_final.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_reference.log 2>&1
for c in 1 2; do timeout 400 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_final_cfg$c.log 2>&1; done
timeout 500 python bench.py --config 4 --batch 8192 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_final_cfg4.log 2>&1
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/plain1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_cfg3_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/ncu1.log 2>&1
python scripts/bench_small.py 3 2048 2 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_synth_tc -c 1 -o /tmp/tc python scripts/bench_small.py 3 2048 2 > gpurun_out/ncu2.log 2>&1
python scripts/ncu_summary.py /tmp/tc.ncu-rep k_synth > gpurun_out/tc_summary.txt 2>&1
ncu -i /tmp/tc.ncu-rep --page raw --csv > gpurun_out/tc_raw.csv 2>/dev/null
ncu -i /tmp/tc.ncu-rep --page source --csv > gpurun_out/tc_source.csv 2>/dev/null
ls -la gpurun_out/
