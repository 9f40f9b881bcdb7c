#!/bin/bash
# Runs every measurement driver on the GPU box and leaves the outputs in gpurun_out/ (copied to profiles/ afterwards).
#   gpurun --timeout 1500 -- 'bash profiles/run_evidence.sh r01'
R=${1:-r01}
O=gpurun_out
mkdir -p $O
python bench.py > $O/${R}_bench_n1.json 2> $O/${R}_bench_n1.err
python bench.py --impl reference > $O/${R}_bench_reference_n1.json 2>> $O/${R}_bench_n1.err
python profiles/standalone_sweep.py > $O/${R}_standalone_sweep.jsonl 2> $O/${R}_standalone_sweep.err
python profiles/qat_images_per_s.py --batch 128 --steps 20 2>/dev/null | tail -1 > $O/${R}_qat_images_per_s.json
python profiles/qat_images_per_s.py --batch 128 --steps 20 --channels-last 2>/dev/null | tail -1 > $O/${R}_qat_images_per_s_channels_last.json
python profiles/rootq_c1.py 2>/dev/null | tail -1 > $O/${R}_rootq_c1.json
python profiles/calibration_c3_c4.py 2>/dev/null | tail -2 > $O/${R}_calibration_c3_c4.jsonl
python profiles/prof_fq.py 67108864 3 > $O/${R}_prof_fq_plain.log 2>&1
tail -n 2 $O/${R}_bench_n1.json | cut -c1-400
