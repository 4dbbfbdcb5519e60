# usage: bash tools/run_n2.sh [tag]  -- 2-GPU box: NCCL test, N = 1 and N = 2 lines of the collective-bearing modes
R=${1:-r2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${R}_gpu_dist_tests.txt
timeout 200 $TR --master-port 29602 bench.py --gpus 2 --mode train --steps 20 --warmup 3 > gpurun_out/${R}_train_n2.json 2> gpurun_out/${R}_train_n2.err; tail -2 gpurun_out/${R}_train_n2.err
timeout 300 $TR --master-port 29603 bench.py --gpus 2 --mode eval10k --episodes 2000 > gpurun_out/${R}_eval10k_n2.json 2> gpurun_out/${R}_eval10k_n2.err; tail -2 gpurun_out/${R}_eval10k_n2.err
timeout 200 python bench.py --mode train --steps 20 --warmup 3 > gpurun_out/${R}_train_n1.json 2> gpurun_out/${R}_train_n1.err
timeout 300 python bench.py --mode eval10k --episodes 2000 > gpurun_out/${R}_eval10k_n1.json 2> gpurun_out/${R}_eval10k_n1.err
timeout 300 python bench.py --mode eval10k --episodes 2000 --bf16-backbone > gpurun_out/${R}_eval10k_bf16_n1.json 2> gpurun_out/${R}_eval10k_bf16_n1.err
timeout 300 $TR --master-port 29604 bench.py --gpus 2 --steps 20 --warmup 3 --no-extras > gpurun_out/${R}_bench_n2.json 2> gpurun_out/${R}_bench_n2.err; tail -2 gpurun_out/${R}_bench_n2.err
cut -c1-300 gpurun_out/${R}_train_n1.json gpurun_out/${R}_train_n2.json gpurun_out/${R}_eval10k_n1.json gpurun_out/${R}_eval10k_n2.json gpurun_out/${R}_eval10k_bf16_n1.json gpurun_out/${R}_bench_n2.json
