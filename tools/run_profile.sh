# Round-2 measurement pass on one B200 (run under gpurun): tests, bench lines, launch lists, ncu --set full captures.
# Every ncu command is preceded by the same command without ncu (&&) as the profiling recipe asks.
set -x
T=${1:-r2}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${T}_pytest.log
python bench.py --steps 2000 --warmup 50 > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_reference_arm.err
for c in 2s 2x 3 4 5; do
  python bench.py --config $c > gpurun_out/${T}_bench_config$c.json 2> gpurun_out/${T}_bench_config$c.err
done
SHORT="--steps 5 --warmup 3 --latency-calls 10 --no-cpu-baseline --no-extras"
python bench.py $SHORT > gpurun_out/${T}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py $SHORT > gpurun_out/${T}_ncu_1.log 2>&1
python bench.py $SHORT > gpurun_out/${T}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tron1_solve_kernel -s 6 -c 2 -o gpurun_out/${T}_prof_c2 -f python bench.py $SHORT > gpurun_out/${T}_ncu_2.log 2>&1
for c in 2s 3 4; do
  python bench.py --config $c $SHORT > gpurun_out/${T}_plain.log 2>&1 && \
  ncu --set full --clock-control none -k regex:tron1_solve_kernel -s 6 -c 2 -o gpurun_out/${T}_prof_c$c -f python bench.py --config $c $SHORT > gpurun_out/${T}_ncu_c$c.log 2>&1
done
python bench.py --config 5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:tron1_rollout_kernel -s 2 -c 1 -o gpurun_out/${T}_prof_c5 -f python bench.py --config 5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_ncu_c5.log 2>&1
# the reports are tens of MB each and gpurun_out/ is capped at 64 MiB: keep the raw-metric page (and the source page of the
# two kernels under work) as CSV, drop the reports
for f in gpurun_out/${T}_prof_*.ncu-rep; do
  b=${f%.ncu-rep}
  ncu -i $f --page raw --csv > $b.raw.csv 2>/dev/null
done
for c in c2 c4; do
  ncu -i gpurun_out/${T}_prof_$c.ncu-rep --page source --csv > gpurun_out/${T}_prof_$c.source.csv 2>/dev/null
  gzip -f gpurun_out/${T}_prof_$c.source.csv
done
ls -la gpurun_out/
rm -f gpurun_out/*.ncu-rep gpurun_out/libmpc_b200_timing.so
