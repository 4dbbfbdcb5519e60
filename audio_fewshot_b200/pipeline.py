"""Waveform -> logits episodic pipeline: the public call a user of this repo makes.

    pipe = EpisodePipeline(frontend, model)          # model: ProtoNet / DN4 / DeepBDC (eval)
    output, acc = pipe(wav, repeats, support_size)   # wav: [N, L] fp32 or int16 PCM, pinned host or CUDA

One call = H2D of the waveforms (when they arrive on the host) -> fused log-mel kernel ->
emb_func (cuDNN) -> head kernel -> vote/accuracy kernel.  Nothing synchronises with the host;
`acc` is a 0-dim CUDA tensor.  With `use_graph=True` the device work of a fixed batch shape is
captured once into a CUDA graph (the per-episode launch sequence is short and launch-bound for
small batches) and replayed from a static input buffer.  The graph bakes in everything the launches take by value:
the cache key therefore covers the batch shape, the episode layout, `first_clip_index`, the front-end's seed and
augmentation switch and the version counters of every model parameter and buffer (a reloaded state_dict or an
optimizer step re-captures).  A graphed call returns the graph's STATIC output tensors -- the next replay overwrites
them; clone what must outlive it.

    for output_host, acc_host in pipe.stream(batches, repeats, support_size): ...

`stream` is the throughput form: pinned host batches are copied by a dedicated copy stream into two
rotating device buffers while the compute stream works on the previous batch, and the logits and the
accuracy come back to pinned host memory the same way, so PCIe and the SMs overlap (a 5w5s15q step
moves 256 MB host->device and computes for ~1.7 ms: the copy is the longer leg).
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

    def _state_key(self):
        """Everything a captured graph holds by value: front-end switches and the versions of the model's tensors
        (folded weights travel as kernel parameters, see Conv64F)."""
        fe = self.frontend
        versions = tuple(t._version for t in list(self.model.parameters()) + list(self.model.buffers()))
        return (bool(fe.training), int(getattr(fe, "seed", 0)), bool(self.model.training), hash(versions))

    @torch.no_grad()
    def __call__(self, wav, repeats, support_size, first_clip_index=0):
        dev = self.frontend.mean.device
        if not self.use_graph:
            wav_dev = wav.to(dev, non_blocking=True)
            return self._device_forward(wav_dev, repeats, support_size, first_clip_index)
        key = (tuple(wav.shape), wav.dtype, support_size, None if repeats is None else bytes(repeats.cpu().numpy().tobytes()),
               int(first_clip_index), self._state_key())
        entry = self._graphs.get(key)
        if entry is None:
            static_in = torch.empty(wav.shape, dtype=wav.dtype, device=dev)
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

    @torch.no_grad()
    def stream(self, batches, repeats, support_size, first_clip_index=0, depth=3):
        """Yield (output, acc) as pinned HOST tensors for every [N, L] pinned host batch of `batches`
        (all of one shape).  Results are yielded `depth` batches late, once their D2H copy has finished."""
        dev = self.frontend.mean.device
        compute = torch.cuda.current_stream(dev)
        copier = torch.cuda.Stream(dev)
        slots = []
        pending = []
        i = 0
        for wav in batches:
            if len(slots) < depth:
                slots.append({"wav": torch.empty(wav.shape, dtype=wav.dtype, device=dev), "out": None,
                              "acc": torch.empty((), dtype=torch.float32).pin_memory(),
                              "free": torch.cuda.Event(), "copied": torch.cuda.Event(), "done": torch.cuda.Event()})
            slot = slots[i % depth]
            if i >= depth:  # the slot's previous result must be handed out before it is overwritten
                prev = pending.pop(0)
                prev["done"].synchronize()
                yield prev["out"].clone(), prev["acc"].clone()
            with torch.cuda.stream(copier):
                copier.wait_event(slot["free"])  # compute finished reading this buffer (no-op the first time)
                slot["wav"].copy_(wav, non_blocking=True)
                slot["copied"].record(copier)
            compute.wait_event(slot["copied"])
            output, acc = self._device_forward(slot["wav"], repeats, support_size, first_clip_index)
            slot["free"].record(compute)
            if slot["out"] is None:
                slot["out"] = torch.empty(output.shape, dtype=output.dtype).pin_memory()
            slot["out"].copy_(output, non_blocking=True)
            slot["acc"].copy_(acc, non_blocking=True)
            slot["done"].record(compute)
            pending.append(slot)
            i += 1
        for prev in pending:
            prev["done"].synchronize()
            yield prev["out"].clone(), prev["acc"].clone()
