"""Whole-step CUDA-graph capture of a QAT training step.

The quantised layers cost ~40 us of host time per differentiable call (Python -> ctypes -> launch); at per-GPU batch
<= 64 a ResNet-50 step is host-bound.  Nothing on the module path synchronises with the host once the
observers have run (host-side init flags, device-resident scales, stream-ordered descriptor uploads), so forward +
backward + optimizer step can be captured ONCE and replayed: the host then issues one cudaGraphLaunch per step.

    step = graph_train_step(model, optimizer, criterion, sample_input, sample_target)
    loss = step(x, t)            # copies x, t into the static buffers, replays, returns the (static) loss tensor

Requirements (checked): observers initialised (run one forward first), single device, `capturable` optimizer state
(torch.optim.SGD is; Adam needs capturable=True).  DistributedDataParallel is not captured here - under DDP the step
is GPU-bound by the gradient all-reduce long before the host matters."""
import torch

__all__ = ["graph_train_step"]


def graph_train_step(model, optimizer, criterion, sample_input, sample_target, warmup=3):
    from .scalar.modules.base import QBase
    for m in model.modules():
        if isinstance(m, QBase):
            h = getattr(m, "_host_init", None) or {}
            q = m.qconfig
            if (q["input"]["enable"] and not h.get("in")) or (q["weight"]["enable"] and not h.get("wt")):
                raise RuntimeError("graph_train_step: run one forward first - the lazy observer initialisation "
                                   "(modules/base.py:82-94,107-129) reads statistics back and cannot be captured")
    static_x = sample_input.clone()
    static_t = sample_target.clone()

    def one_step():
        optimizer.zero_grad(set_to_none=True)
        loss = criterion(model(static_x), static_t)
        loss.backward()
        optimizer.step()
        return loss

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(warmup):                    # allocator warm-up, descriptor tables, cuDNN algorithm selection
            one_step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    optimizer.zero_grad(set_to_none=True)
    with torch.cuda.graph(graph):
        static_loss = one_step()

    def step(x, t):
        static_x.copy_(x, non_blocking=True)
        static_t.copy_(t, non_blocking=True)
        graph.replay()
        return static_loss

    step.graph, step.static_input, step.static_target, step.static_loss = graph, static_x, static_t, static_loss
    return step
