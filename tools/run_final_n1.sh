# One-GPU evidence batch of a round: tests, smoke, bench line, per-kernel table, per-config table, launch list.
R=${1:-r02}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${R}_gpu_tests.txt; cat gpurun_out/${R}_gpu_tests.txt
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -4
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err; tail -2 gpurun_out/${R}_bench_n1.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${R}_bench_reference_arm.json 2> /dev/null
timeout 400 python bench_kernels.py > gpurun_out/${R}_kernels.jsonl 2> gpurun_out/${R}_kernels.err; tail -2 gpurun_out/${R}_kernels.err
timeout 400 python bench_configs.py > gpurun_out/${R}_configs.jsonl 2> gpurun_out/${R}_configs.err; tail -2 gpurun_out/${R}_configs.err
timeout 200 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${R}_launches_final.csv \
  python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
wc -l gpurun_out/${R}_launches_final.csv
