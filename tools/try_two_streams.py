"""Experiment: does running two half-batches on two streams overlap the LSU-bound log-mel kernel with the
TMEM/tensor-bound backbone kernels?  (development helper)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_fewshot_b200 import model as arch
from audio_fewshot_b200.frontend import LogMelFrontEnd
dev = torch.device("cuda", 0)
torch.manual_seed(0)
W, S, Q = 5, 5, 15
emb = arch.Conv64F(is_flatten=True, num_channels=1)
model = arch.ProtoNet(way_num=W, shot_num=S, query_num=Q, test_way=W, test_shot=S, test_query=Q, emb_func=emb, device=dev).to(dev).eval()
front = LogMelFrontEnd(hop_length=512, n_mels=128, mean=-15.0, std=26.0).to(dev).eval()

def run(E_total, n_streams, iters=30):
    E = E_total // n_streams
    wavs = [[torch.randn(E * 100, 80000, device=dev) * 0.1 for _ in range(2)] for _ in range(n_streams)]
    rep = torch.ones(E * W * Q, dtype=torch.long)
    streams = [torch.cuda.Stream() for _ in range(n_streams)]
    def step(i):
        for s, st in enumerate(streams):
            with torch.cuda.stream(st):
                img = front(wavs[s][i % 2])
                model.set_forward([img, None, rep, E * W * S])
    with torch.no_grad():
        for i in range(3):
            step(i)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for st in streams:
            st.wait_event(t0)
        for i in range(iters):
            step(i)
        for st in streams:
            torch.cuda.current_stream().wait_stream(st)
        t1.record()
        torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / iters
    print("E_total %d streams %d: %.3f ms/step, %.0f episodes/s" % (E_total, n_streams, ms, E_total / ms * 1e3), flush=True)

for E_total in (8, 16, 32):
    for ns in (1, 2, 4):
        run(E_total, ns)
