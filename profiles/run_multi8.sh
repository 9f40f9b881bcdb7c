#!/bin/bash
# 8-GPU evidence, trimmed to what 8x-charged box time allows: sharded standalone sweep (BASELINE configs[4]), the
# headline metric, and the channels_last QAT arms under DDP (eager and whole-step CUDA graph).
#   gpurun --gpus 8 --timeout 900 -- 'bash profiles/run_multi8.sh 8 r02'
N=${1:-8}
R=${2:-r02}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29524 profiles/standalone_sweep_multi.py 2>$O/${R}_standalone_multi_n$N.err > $O/${R}_standalone_multi_n$N.jsonl
wc -l $O/${R}_standalone_multi_n$N.jsonl
timeout 200 $TR --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 --no-qat --e2e-steps 2 2>$O/${R}_bench_n$N.err | tail -1 > $O/${R}_bench_n$N.json
cut -c1-200 $O/${R}_bench_n$N.json
timeout 400 $TR --master-port 29522 bench.py --gpus $N --qat-only --qat-arms fp32,ours_fused,fp32_graphed,ours_fused_graphed 2>>$O/${R}_bench_n$N.err | tail -1 > $O/${R}_bench_qat_n$N.json
cut -c1-1200 $O/${R}_bench_qat_n$N.json
