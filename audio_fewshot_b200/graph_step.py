"""A whole episodic training step as ONE CUDA graph.

The reference's training loop (libfewshot_core/trainer.py:186-192) runs, per batch,

    output, acc, loss = model(batch);  optimizer.zero_grad();  loss.backward();  optimizer.step()

as thousands of small launches: for MAML (config C5) every step is 2 episodes x 5 inner SGD steps, each a forward and
a create_graph=True backward through Conv64F on 25 clips, then the second-order outer backward -- 0.16 s of launch
latency for a few milliseconds of GPU work.  `GraphedTrainStep` captures exactly that sequence once, for a fixed
batch shape, into a CUDA graph and replays it from a static input buffer; the arithmetic (kernels, their order,
the Dropout Philox stream handling, the optimizer update) is PyTorch's own, only the launch path changes.

    step = GraphedTrainStep(model, optimizer, image_shape=(150, 1, 128, 157), target=target)
    output, acc, loss = step(image)          # image: CUDA or pinned-host tensor of that shape

Requirements (checked): the model is one of this package's classes in train() mode; the optimizer must not sync with
the host inside step() (torch.optim.Adam/AdamW need capturable=True; SGD is fine).  `acc` comes back as a 1-element
device tensor (model.acc_on_device is switched on), `output` / `loss` are the graph's static tensors -- copy them
if they must outlive the next call.  With torch.distributed initialised, pass reduce_gradients=True to capture the
flat NCCL gradient all-reduce (dist.all_reduce_gradients) between backward and step.

Construction runs `warmup` eager steps and the capture itself on an all-zero batch (cuDNN heuristics, episode tables,
lazily created optimizer state).  Those steps would move parameters, BatchNorm running statistics and optimizer
moments; the model's and the optimizer's state are therefore snapshotted before and restored IN PLACE after them
(the graph keeps pointing at the same tensors), so the first replay starts from exactly the caller's state.
"""
import copy

import torch

from . import dist as afs_dist


class GraphedTrainStep:
    def __init__(self, model, optimizer, image_shape, target=None, repeats=None, support_size=0, warmup=3,
                 reduce_gradients=False, dtype=torch.float32):
        if not model.training:
            raise ValueError("GraphedTrainStep captures set_forward_loss: call model.train() first")
        for group in optimizer.param_groups:
            if "capturable" in group and not group["capturable"]:
                raise ValueError("construct %s with capturable=True to use it inside a CUDA graph"
                                 % type(optimizer).__name__)
        self.model, self.optimizer = model, optimizer
        dev = torch.device(model.device)
        self.static_image = torch.zeros(image_shape, dtype=dtype, device=dev)
        self._rest = [target] if repeats is None else [target, repeats, support_size]
        self.reduce_gradients = reduce_gradients
        model.acc_on_device = True

        model_state = {k: v.detach().clone() for k, v in model.state_dict().items()}
        optim_state = copy.deepcopy(optimizer.state_dict())
        had_state = {id(p) for p in optimizer.state}

        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):  # eager warm-up: cuDNN heuristics, episode tables, optimizer state
            for _ in range(max(1, warmup)):
                self._eager_step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)

        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)  # backward inside the capture then WRITES fresh static .grad tensors
        with torch.cuda.graph(self.graph):
            self.output, self.acc, self.loss = self._eager_step()
        self._restore(model_state, optim_state, had_state)

    def _restore(self, model_state, optim_state, had_state):
        """Undo the warm-up / capture steps in place (graph-captured tensors keep their addresses)."""
        with torch.no_grad():
            for k, v in self.model.state_dict().items():
                v.copy_(model_state[k])
            # optimizer state: tensors are copied into the live (captured) ones; entries that did not exist before
            # the warm-up (lazily created moments, step counters) are reset to zero, their initial value
            saved = optim_state["state"]
            index = {}
            i = 0
            for group in self.optimizer.param_groups:
                for prm in group["params"]:
                    index[id(prm)] = i
                    i += 1
            for prm, st in self.optimizer.state.items():
                old = saved.get(index[id(prm)]) if id(prm) in had_state else None
                for name, val in st.items():
                    if torch.is_tensor(val):
                        if old is not None and torch.is_tensor(old.get(name)):
                            val.copy_(old[name])
                        else:
                            val.zero_()
                    elif old is not None and name in old:
                        st[name] = old[name]
        torch.cuda.synchronize(torch.device(self.model.device))

    def _eager_step(self):
        self.optimizer.zero_grad(set_to_none=True)
        output, acc, loss = self.model([self.static_image] + self._rest)
        loss.backward()
        if self.reduce_gradients:
            afs_dist.all_reduce_gradients(self.model.parameters())
        self.optimizer.step()
        return output, acc, loss

    def __call__(self, image):
        self.static_image.copy_(image, non_blocking=True)
        self.graph.replay()
        return self.output, self.acc, self.loss
