"""SURVEY.md 8e, per-channel weight observers sharded by output channel: the 80-candidate sweep over the
54 ResNet-50 weight tensors (25.5 M elements, 27 560 rows), each rank sweeping ceil(C / world) rows of
every tensor, qparams all-gathered over NCCL.  Checks that the gathered qparams equal the unsharded ones
bit for bit and times both.  Launch under torchrun (N = 2, 4, 8); one JSON line from rank 0.

    python -m torch.distributed.run --nproc-per-node N profiles/sharded_observer.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(device)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    import torchvision
    from dlmc_quant_b200 import dist as qdist
    from dlmc_quant_b200 import functional as F
    from dlmc_quant_b200.scalar import ops
    torch.manual_seed(2333)                                   # identical (replicated) weights on every rank
    model = torchvision.models.resnet50()
    weights = [m.weight.detach().to(device) for m in model.modules()
               if isinstance(m, (torch.nn.Conv2d, torch.nn.Linear))]

    def sweep_one(w, mode):
        rows = w.reshape(w.shape[0], -1)
        if mode == "local":
            return F.sweep_channel(rows, 4, True)
        thr = 64 if mode == "forced" else 2048               # "default" = the library's threshold
        return qdist.rows_sharded(rows, lambda blk: F.sweep_channel(blk, 4, True, rows.shape[0]), min_rows_per_rank=thr)

    def timed(tensors, mode, reps=5):
        run = lambda: [sweep_one(w, mode) for w in tensors]
        run()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            run()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / reps], device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    def same(tensors, mode):
        return all(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
                   for a, b in ((sweep_one(w, "local"), sweep_one(w, mode)) for w in tensors))

    big = [torch.randn(32768, 1152, device=device) * 0.02]     # an embedding / large-FC sized matrix (37.7 M elements)
    out = {"op": "quantize_l2loss_channel (W4 signed, 80 candidates), rows sharded by output channel + all-gather",
           "n_gpus": world,
           "resnet50_54_tensors": {"rows": sum(w.shape[0] for w in weights), "elements": sum(w.numel() for w in weights),
                                   "bit_identical_forced_sharding": same(weights, "forced"),
                                   "ms_every_rank_sweeps_all": round(timed(weights, "local"), 3),
                                   "ms_forced_sharding_108_allgathers": round(timed(weights, "forced"), 3),
                                   "ms_library_default": round(timed(weights, "default"), 3)},
           "matrix_32768x1152": {"bit_identical": same(big, "default"),
                                 "ms_every_rank_sweeps_all": round(timed(big, "local"), 3),
                                 "ms_sharded_allgather": round(timed(big, "default"), 3)}}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
