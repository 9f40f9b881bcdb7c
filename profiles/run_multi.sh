#!/bin/bash
# Multi-GPU evidence on one box: 2-rank parity tests, the sharded standalone sweep (BASELINE configs[4]) and bench.py
# (headline metric + ResNet-50 QAT arms under DDP).
#   gpurun --gpus N --timeout 1500 -- 'bash profiles/run_multi.sh N r02 [nobench]'
N=${1:-2}
R=${2:-r02}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_parity_edges.py -q 2>&1 | tail -3 | tee $O/${R}_two_rank_tests.log
fi
timeout 900 $TR --master-port 29524 profiles/standalone_sweep_multi.py 2>$O/${R}_standalone_multi_n$N.err > $O/${R}_standalone_multi_n$N.jsonl
cat $O/${R}_standalone_multi_n$N.jsonl | cut -c1-200
if [ "$3" != "nobench" ]; then
  timeout 900 $TR --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 2>$O/${R}_bench_n$N.err | tail -1 > $O/${R}_bench_n$N.json
  cut -c1-300 $O/${R}_bench_n$N.json
fi
