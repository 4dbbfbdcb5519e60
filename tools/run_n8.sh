# usage: bash tools/run_n8.sh N [tag]  -- the path's multi-GPU lines on one box with N GPUs (gpurun --gpus N)
N=${1:-8}
R=${2:-r2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29612 bench.py --gpus $N --steps 20 --warmup 3 --no-extras > gpurun_out/${R}_bench_n$N.json 2> gpurun_out/${R}_bench_n$N.err; tail -1 gpurun_out/${R}_bench_n$N.err
timeout 300 $TR --master-port 29613 bench.py --gpus $N --mode train --steps 20 --warmup 3 > gpurun_out/${R}_train_n$N.json 2> gpurun_out/${R}_train_n$N.err; tail -1 gpurun_out/${R}_train_n$N.err
timeout 400 $TR --master-port 29614 bench.py --gpus $N --mode eval10k --episodes 10000 --episodes-per-step 5 > gpurun_out/${R}_eval10k_n$N.json 2> gpurun_out/${R}_eval10k_n$N.err; tail -1 gpurun_out/${R}_eval10k_n$N.err
timeout 400 $TR --master-port 29615 bench.py --gpus $N --mode eval10k --episodes 10000 --episodes-per-step 5 --bf16-backbone > gpurun_out/${R}_eval10k_bf16_n$N.json 2> gpurun_out/${R}_eval10k_bf16_n$N.err; tail -1 gpurun_out/${R}_eval10k_bf16_n$N.err
cut -c1-400 gpurun_out/${R}_bench_n$N.json gpurun_out/${R}_train_n$N.json gpurun_out/${R}_eval10k_n$N.json gpurun_out/${R}_eval10k_bf16_n$N.json
