#!/bin/bash
# Multi-GPU evidence on one box: bench.py, end-to-end QAT images/s (DDP) and the sharded per-channel observer.
#   gpurun --gpus N --timeout 900 -- 'bash profiles/run_multi.sh N r01'
N=${1:-2}
R=${2:-r01}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 2>$O/${R}_bench_n$N.err | tail -1 > $O/${R}_bench_n$N.json
$TR --master-port 29522 profiles/qat_images_per_s.py --batch 128 --steps 20 2>$O/${R}_qat_n$N.err | tail -1 > $O/${R}_qat_images_per_s_n$N.json
$TR --master-port 29523 profiles/sharded_observer.py 2>$O/${R}_sharded_n$N.err | tail -1 > $O/${R}_sharded_observer_n$N.json
cut -c1-260 $O/${R}_bench_n$N.json; echo; cat $O/${R}_qat_images_per_s_n$N.json $O/${R}_sharded_observer_n$N.json; tail -2 $O/${R}_sharded_n$N.err
