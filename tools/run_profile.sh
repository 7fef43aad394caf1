set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_r1f.log
python bench.py --steps 2000 --warmup 50 > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err
python bench.py --steps 5 --warmup 3 --latency-calls 10 --no-cpu-baseline > gpurun_out/plain_r1f.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 5 --warmup 3 --latency-calls 10 --no-cpu-baseline > gpurun_out/ncu_r1f_1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tron1_solve_kernel -s 6 -c 2 -o gpurun_out/prof_r1f -f python bench.py --steps 5 --warmup 3 --latency-calls 10 --no-cpu-baseline > gpurun_out/ncu_r1f_2.log 2>&1
cat gpurun_out/pytest_r1f.log
cut -c1-300 gpurun_out/bench_r1f.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r1f_ref.json 2> gpurun_out/bench_r1f_ref.err
python tools/bench_configs.py > gpurun_out/configs_r1f.jsonl 2> gpurun_out/configs_r1f.err
python tools/host_path_probe.py > gpurun_out/host_probe_r1f.log 2>&1
