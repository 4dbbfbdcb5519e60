# One-GPU evidence batch of a round: tests, smoke, bench line, per-kernel table, per-config table, launch list, stem ncu.
R=${1:-r02}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${R}_gpu_tests.txt; cat gpurun_out/${R}_gpu_tests.txt
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -6 | tee gpurun_out/${R}_smoke.txt
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err; tail -2 gpurun_out/${R}_bench_n1.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${R}_bench_reference_arm.json 2> /dev/null
timeout 400 python bench_kernels.py > gpurun_out/${R}_kernels.jsonl 2> gpurun_out/${R}_kernels.err; tail -2 gpurun_out/${R}_kernels.err
timeout 400 python bench_configs.py > gpurun_out/${R}_configs.jsonl 2> gpurun_out/${R}_configs.err; tail -2 gpurun_out/${R}_configs.err
timeout 300 python bench_configs.py --bf16 --only C1,C2,C4 > gpurun_out/${R}_configs_bf16.jsonl 2>> gpurun_out/${R}_configs.err
timeout 300 python tools/run_backbone_bf16.py > gpurun_out/${R}_backbone_bf16.txt 2>&1
timeout 200 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${R}_launches_final.csv \
  python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
wc -l gpurun_out/${R}_launches_final.csv
timeout 100 python tools/run_conv1_tc_only.py > gpurun_out/plain2.log 2>&1 && \
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv1_tc -s 2 -c 1 -f -o gpurun_out/${R}_stem \
  python tools/run_conv1_tc_only.py > gpurun_out/ncu_stem.log 2>&1
tail -2 gpurun_out/ncu_stem.log
