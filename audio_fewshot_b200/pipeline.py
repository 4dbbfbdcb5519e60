"""Waveform -> logits episodic pipeline: the public call a user of this repo makes.

    pipe = EpisodePipeline(frontend, model)          # model: ProtoNet / DN4 / DeepBDC (eval)
    output, acc = pipe(wav, repeats, support_size)   # wav: [N, L] fp32, pinned host or CUDA

One call = H2D of the waveforms (when they arrive on the host) -> fused log-mel kernel ->
emb_func (cuDNN) -> head kernel -> vote/accuracy kernel.  Nothing synchronises with the host;
`acc` is a 0-dim CUDA tensor.  With `use_graph=True` the device work of a fixed batch shape is
captured once into a CUDA graph (the per-episode launch sequence is short and launch-bound for
small batches) and replayed from a static input buffer.
"""
import torch


class EpisodePipeline:
    def __init__(self, frontend, model, use_graph=False, channels_last=False):
        self.frontend = frontend
        self.model = model
        self.use_graph = use_graph
        self.channels_last = channels_last
        self._graphs = {}

    def _device_forward(self, wav_dev, repeats, support_size, first_clip_index=0):
        image = self.frontend(wav_dev, first_clip_index=first_clip_index)
        if self.channels_last:
            image = image.contiguous(memory_format=torch.channels_last)
        target = None
        return self.model.set_forward([image, target, repeats, support_size])

    @torch.no_grad()
    def __call__(self, wav, repeats, support_size, first_clip_index=0):
        dev = self.frontend.mean.device
        if not self.use_graph:
            wav_dev = wav.to(dev, non_blocking=True)
            return self._device_forward(wav_dev, repeats, support_size, first_clip_index)
        key = (tuple(wav.shape), support_size, None if repeats is None else bytes(repeats.cpu().numpy().tobytes()))
        entry = self._graphs.get(key)
        if entry is None:
            static_in = torch.empty(wav.shape, dtype=torch.float32, device=dev)
            static_in.copy_(wav, non_blocking=True)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):  # warm-up outside capture: plans, cuDNN heuristics, tables
                for _ in range(2):
                    self._device_forward(static_in, repeats, support_size, first_clip_index)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self._device_forward(static_in, repeats, support_size, first_clip_index)
            entry = (graph, static_in, out)
            self._graphs[key] = entry
        graph, static_in, out = entry
        static_in.copy_(wav, non_blocking=True)
        graph.replay()
        return out
