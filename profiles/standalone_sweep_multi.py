"""BASELINE.json configs[4] on N GPUs: the standalone fake-quant fwd+bwd and observer kernels on a tensor of N_total
elements SHARDED over the ranks (per-tensor tensors in contiguous 1/world slices, per-channel weight
matrices by output-channel blocks), with the path's only exchanges inside the timed region:
  * per-tensor fwd+bwd : every rank runs its slice, the scale gradient (1 float) is all-reduced (SUM);
  * min/max observer   : local statistics, all-reduce (MIN/MAX/SUM) of the [1,4] statistics, finalise;
  * per-channel fwd+bwd: rows split over the ranks, no collective (every row's scale gradient is complete locally);
  * per-channel sweep  : rows split over the ranks, (scale, offset) blocks all-gathered (dist.rows_sharded).
Timing: CUDA events around `iters` back-to-back repetitions, max over ranks; GB/s on algorithmic bytes (20 B/elem
fwd+bwd fp32, 4 B/elem observers).  Strong scaling: the efficiency is (aggregate GB/s at N) / (N x GB/s at 1 GPU), where
the 1-GPU figure is this script run with one rank on the same shapes (profiles/r02_standalone_multi_n1.jsonl).

    torchrun --nproc-per-node N profiles/standalone_sweep_multi.py > profiles/r02_standalone_multi_nN.jsonl"""
import json
import math
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlmc_quant_b200 import dist as qdist  # noqa: E402
from dlmc_quant_b200 import functional as F  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) * 1e-3 / iters], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def emit(**kw):
    if rank == 0:
        print(json.dumps(kw), flush=True)


for log2 in (24, 26, 28, 30):
    n_total = 1 << log2
    n = n_total // world
    iters = max(5, min(100, int(4e9 / (n * 4))))
    g = torch.Generator(device=dev).manual_seed(2333 + rank)
    x = torch.relu(torch.randn(n, generator=g, device=dev)) * 2
    dy = torch.randn(n, generator=g, device=dev)
    stats = qdist.sync_stats(F.obs_stats(x))
    scale, off = F.minmax_from_stats(stats, 4, False)
    gq = 1 / math.sqrt(n_total * 15)
    y = torch.empty_like(x)

    def fwd_bwd():
        F.fq_forward(x, scale, off, 0, 15, F.FORM_AFFINE, g=gq)
        dx, ds = F.fq_backward(x, dy, scale, off, 0, 15, F.FORM_AFFINE, g=gq)
        qdist.allreduce_grads_(ds)
    t = timed(fwd_bwd, iters)
    emit(case="per_tensor_fwd_bwd", n_total=n_total, n_gpus=world, us=round(t * 1e6, 1),
         gbps=round(20 * n_total / t / 1e9, 1), collective="all_reduce(SUM) of 1 scale-gradient float")

    def observer():
        st = qdist.sync_stats(F.obs_stats(x))
        F.minmax_from_stats(st, 4, False)
    t = timed(observer, iters)
    emit(case="minmax_observer", n_total=n_total, n_gpus=world, us=round(t * 1e6, 1), gbps=round(4 * n_total / t / 1e9, 1),
         collective="3 all_reduces (MIN / MAX / SUM) of the [1,4] statistics")
    del x, dy, y
    torch.cuda.empty_cache()

for c_total, k in ((32768, 1152), (262144, 1152)):
    per = c_total // world
    g = torch.Generator(device=dev).manual_seed(7)
    w_full_rows = c_total                                   # replicated weights: every rank could hold all rows
    w = torch.randn(per, k, generator=g, device=dev) * 0.02
    dw = torch.randn(per, k, generator=g, device=dev)
    st = F.obs_stats(w, ch_axis=0)
    s, _ = F.minmax_from_stats(st, 4, True)
    n_total = c_total * k

    def wt():
        F.fq_forward(w, s, None, -7, 7, F.FORM_SYM, ch_axis=0)
        F.fq_backward(w, dw, s, None, -7, 7, F.FORM_SYM, ch_axis=0)
    t = timed(wt, 20)
    emit(case="per_channel_fwd_bwd", rows_total=c_total, inner=k, n_gpus=world, us=round(t * 1e6, 1),
         gbps=round(20 * n_total / t / 1e9, 1), collective="none (rows split by output channel)")
    if c_total <= 32768:
        full = torch.randn(c_total, k, generator=torch.Generator(device=dev).manual_seed(9), device=dev) * 0.02

        def sweep():
            qdist.rows_sharded(full, lambda blk: F.sweep_channel(blk, 4, True, full.shape[0]), min_rows_per_rank=1024)
        t = timed(sweep, 5)
        emit(case="per_channel_mse_sweep", rows_total=c_total, inner=k, n_gpus=world, us=round(t * 1e6, 1),
             candidate_evals_per_s=round(80 * n_total / t, 1), collective="all_gather of (scale, offset) row blocks")
        del full
    del w, dw
    torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
