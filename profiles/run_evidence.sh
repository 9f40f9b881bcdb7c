#!/bin/bash
# Runs every single-GPU measurement driver on the GPU box and leaves the outputs in gpurun_out/ (copied to profiles/
# afterwards).  Every ncu capture comes AFTER the same program has run plain and exited 0.
#   gpurun --timeout 2400 -- 'bash profiles/run_evidence.sh r02'
R=${1:-r02}
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/${R}_smoke.log 2>&1
python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/${R}_gpu_tests_full.log
python bench.py > $O/${R}_bench_n1.json 2> $O/${R}_bench_n1.err && echo bench ok
python bench.py --impl reference > $O/${R}_bench_reference_n1.json 2>> $O/${R}_bench_n1.err
# launch list of the bench step (kernel share of the step) and DRAM bytes of every fq_bwd_flat launch (roofline.traffic)
python bench.py --steps 2 --warmup 3 --no-e2e --no-qat > /dev/null 2>&1 && \
  $NCU --metrics gpu__time_duration.sum -k regex:"fq_|grouped|finalize" -c 800 --csv --log-file $O/${R}_launches_bench_step.csv \
       python bench.py --steps 2 --warmup 3 --no-e2e --no-qat > /dev/null 2>&1
$NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:"fq_bwd_flat" -c 108 --csv \
     --log-file $O/${R}_traffic_fq_bwd_flat.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-qat > /dev/null 2>&1
# full captures: flat fake-quant kernels, fused BatchNorm kernels
python profiles/prof_fq.py 67108864 3 > $O/${R}_prof_fq_plain.log 2>&1 && \
  $NCU --set full --import-source on -k regex:fq_ -c 4 -o /tmp/${R}_fq_flat python profiles/prof_fq.py 67108864 2 > /dev/null 2>&1
python profiles/ncu_summary.py /tmp/${R}_fq_flat.ncu-rep > $O/${R}_ncu_full_fq_flat_2p26.txt 2>&1
python profiles/prof_bnq.py > $O/${R}_prof_bnq.jsonl 2> $O/${R}_prof_bnq.err && \
  $NCU --set full --import-source on -k regex:bnq -c 24 -o /tmp/${R}_bnq python profiles/prof_bnq.py --one > /dev/null 2>&1
python profiles/ncu_summary.py /tmp/${R}_bnq.ncu-rep > $O/${R}_ncu_full_bnq.txt 2>&1
# the .ncu-rep files stay on the box unless small (gpurun_out/ is capped at 64 MiB): the summaries above are what is judged
for f in /tmp/${R}_fq_flat.ncu-rep /tmp/${R}_bnq.ncu-rep; do [ -f $f ] && [ $(stat -c %s $f) -lt 25000000 ] && cp $f $O/; done
$NCU --metrics gpu__time_duration.sum,smsp__inst_executed.sum --csv --log-file $O/${R}_launches_bnq.csv python profiles/prof_bnq.py --one > /dev/null 2>&1
# this round's observer kernels: one-read k-th value, slab kernels (7x7 planes), grouped RootQ launches
$NCU --set full -k regex:"kth_" -c 3 -o /tmp/${R}_kth python profiles/prof_kth.py > /dev/null 2>&1
python profiles/ncu_summary.py /tmp/${R}_kth.ncu-rep > $O/${R}_ncu_full_kth.txt 2>&1
$NCU --set full -k regex:"slab|cmaj|l2norm_resident|sweep_channel_grouped" -c 8 -o /tmp/${R}_obs python profiles/prof_obs.py 1 > /dev/null 2>&1
python profiles/ncu_summary.py /tmp/${R}_obs.ncu-rep > $O/${R}_ncu_full_observers.txt 2>&1
$NCU --set full -k regex:"rootq_grouped" -c 4 -o /tmp/${R}_rootq python profiles/rootq_c1.py --cpu-passes 1 > /dev/null 2>&1
python profiles/ncu_summary.py /tmp/${R}_rootq.ncu-rep > $O/${R}_ncu_full_rootq_grouped.txt 2>&1
# plain drivers
python profiles/prof_obs.py 5 > $O/${R}_prof_obs_plain.log 2>&1
python profiles/prof_kth.py > $O/${R}_prof_kth.log 2>&1
python profiles/standalone_sweep.py > $O/${R}_standalone_sweep.jsonl 2> $O/${R}_standalone_sweep.err
python profiles/standalone_sweep_multi.py > $O/${R}_standalone_multi_n1.jsonl 2>/dev/null
python profiles/rootq_c1.py 2>/dev/null | tail -1 > $O/${R}_rootq_c1.json
python profiles/calibration_c3_c4.py 2>/dev/null | tail -2 > $O/${R}_calibration_c3_c4.jsonl
python profiles/fsptq_recon_c3.py --iters 200 2>/dev/null | tail -1 > $O/${R}_fsptq_recon_c3.json
python profiles/qat_kernel_breakdown.py ours_fused channels_last > $O/${R}_qat_breakdown_fused_cl.txt 2>&1
python profiles/qat_kernel_breakdown.py fp32 channels_last > $O/${R}_qat_breakdown_fp32_cl.txt 2>&1
python profiles/host_profile.py ours_fused 2>&1 | grep -v "^$" | head -60 | cut -c1-170 > $O/${R}_host_profile_fused.txt
timeout 120 tests/_build/abi_consumer > $O/${R}_abi_consumer.log 2>&1
tail -n 1 $O/${R}_bench_n1.json | cut -c1-300
