"""What can the host side of this box deliver?  Plain pinned cudaMemcpyAsync host->device, one process per GPU,
all ranks at once, with and without binding each rank to the CPUs NVML reports as local to its GPU.

    python tools/h2d_probe.py                                    # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        tools/h2d_probe.py                                       # N GPUs, every rank copying at the same time

Prints one JSON line (rank 0): per-rank GB/s (slowest rank), aggregate GB/s, for buffer sizes 64 MB / 256 MB / 1 GB,
unbound and bound, plus `nvidia-smi topo -m` when available.  One cudaMemcpyAsync per buffer -- no batched copies."""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from bench import bind_to_gpu_cpus


def main():
    real_stdout = os.fdopen(os.dup(1), "w")  # NCCL prints its banner on fd 1: keep the JSON line alone on stdout
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    results = []
    for bound in (False, True):
        cpus = bind_to_gpu_cpus(local) if bound else None
        for mb in (64, 256, 1024):
            n = mb * 2 ** 20 // 4
            src = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(2)]  # allocated after binding
            for s in src:
                s.fill_(1.0)
            dst = torch.empty(n, dtype=torch.float32, device=dev)
            dst.copy_(src[0], non_blocking=True)
            reps = max(4, 4096 // mb)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record()
            for i in range(reps):
                dst.copy_(src[i % 2], non_blocking=True)
            e1.record()
            barrier()
            ms = reduce_max(e0.elapsed_time(e1))
            gbs = reps * n * 4 / (ms * 1e-3) / 1e9
            results.append({"bound_to_local_cpus": bound, "n_local_cpus": len(cpus) if cpus else None, "buffer_mb": mb,
                            "gbs_per_gpu_slowest_rank": gbs, "gbs_aggregate": gbs * world})
            del src, dst
    if rank == 0:
        topo = None
        try:
            topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        except Exception:
            pass
        real_stdout.write(json.dumps({"n_gpus": world, "host_cpus": os.cpu_count(), "results": results, "topo": topo}) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
