"""Whole-step CUDA-graph capture of a QAT training step.

The quantised layers cost ~40 us of host time per differentiable call (Python -> ctypes -> launch); at per-GPU batch
<= 64 a ResNet-50 step is host-bound.  Nothing on the module path synchronises with the host once the
observers have run (host-side init flags, device-resident scales, stream-ordered descriptor uploads), so forward +
backward + optimizer step can be captured ONCE and replayed: the host then issues one cudaGraphLaunch per step.

    step = graph_train_step(model, optimizer, criterion, sample_input, sample_target)
    loss = step(x, t)            # copies x, t into the static buffers, replays, returns the (static) loss tensor

Requirements (checked): observers initialised (run one forward first), single device per process, `capturable`
optimizer state (torch.optim.SGD is; Adam needs capturable=True).

DistributedDataParallel: the DDP step is host-bound too (the GPU idles ~7 ms of a 33 ms ResNet-50 step at 2 GPUs while
the NCCL kernels take < 1 ms: profiles/r02_ddp_gpu_busy_n2.jsonl), so the step is captured INCLUDING DDP's bucketed
gradient all-reduces.  PyTorch's rules for that: the DDP wrapper is constructed on a side stream (`wrap_ddp`), at least
11 eager iterations run before capture (bucket rebuild, the reducer's first-iteration logic - `graph_train_step` warms
up 12 when it sees a DDP model).  NCCL >= 2.9.6 captures its collectives as they are (measured here under torchrun,
which sets TORCH_NCCL_ASYNC_ERROR_HANDLING=1: profiles/r02_ddp_graph_probe_n{2,8}.jsonl); older NCCL needs that variable
set to 0 before `init_process_group`.

    model = wrap_ddp(model, device_ids=[local_rank])
    optimizer = torch.optim.SGD(model.parameters(), ...)
    step = graph_train_step(model, optimizer, criterion, x, t)"""
import torch

__all__ = ["graph_train_step", "wrap_ddp"]

_DDP_WARMUP = 12


def wrap_ddp(model, **ddp_kwargs):
    """DistributedDataParallel(model, **ddp_kwargs) constructed on a side stream, as whole-step capture requires."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ddp = torch.nn.parallel.DistributedDataParallel(model, **ddp_kwargs)
    torch.cuda.current_stream().wait_stream(side)
    return ddp


def graph_train_step(model, optimizer, criterion, sample_input, sample_target, warmup=None):
    if warmup is None:
        warmup = _DDP_WARMUP if isinstance(model, torch.nn.parallel.DistributedDataParallel) else 3
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
