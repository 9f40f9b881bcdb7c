"""Percentile observer under data parallelism: each rank holds a shard of the activations, the per-pass histograms
are all-reduced, and every rank must obtain the qparams of the UNION of the shards (== single-process result on the
concatenated tensor).   python -m torch.distributed.run --nproc-per-node 2 profiles/percentile_multi_gpu.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(device)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=device)
from dlmc_quant_b200 import dist as qdist  # noqa: E402
from dlmc_quant_b200.scalar import ops  # noqa: E402

full = torch.relu(torch.randn(world * 64, 32, 14, 14, generator=torch.Generator().manual_seed(2333))) * 2 + 0.25
shard = full[rank * 64:(rank + 1) * 64].to(device)
out = {}
for signed, t in ((False, shard), (True, shard - 1.0)):
    s, o = ops.quantize_percentile_tensor(t, n_bits=8, signed=signed, percentile=99.9)
    qdist.set_enabled(False)
    ref_s, ref_o = ops.quantize_percentile_tensor((full - (1.0 if signed else 0.0)).to(device), n_bits=8, signed=signed,
                                                  percentile=99.9)
    qdist.set_enabled(True)
    out["signed" if signed else "unsigned"] = bool(torch.equal(s, ref_s) and torch.equal(o.float(), ref_o.float()))
flags = torch.tensor([int(all(out.values()))], device=device)
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"n_gpus": world, "percentile_qparams_equal_union": out, "all_ranks_agree": bool(flags.item())}))
dist.destroy_process_group()
