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

    def sweep_all(sharded):
        qdist.set_enabled(sharded)
        out = [ops.quantize_l2loss_channel(w, n_bits=4, signed=True, ch_axis=0) for w in weights]
        qdist.set_enabled(True)
        return out

    def timed(sharded, reps=5):
        sweep_all(sharded)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            sweep_all(sharded)
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / reps], device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    local, shard = sweep_all(False), sweep_all(True)
    same = all(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) for a, b in zip(local, shard))
    t_local, t_shard = timed(False), timed(True)
    if rank == 0:
        print(json.dumps({"op": "quantize_l2loss_channel over the 54 ResNet-50 weight tensors (W4, 80 candidates)",
                          "n_gpus": world, "rows": sum(w.shape[0] for w in weights),
                          "elements": sum(w.numel() for w in weights), "bit_identical_to_unsharded": same,
                          "ms_unsharded_per_rank": round(t_local, 3), "ms_sharded_allgather": round(t_shard, 3),
                          "speedup": round(t_local / t_shard, 2)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
