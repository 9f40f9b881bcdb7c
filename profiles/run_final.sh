#!/bin/bash
# Short end-of-round evidence pass (one GPU, ~4.5 min): smoke, the full GPU suite with skip reasons, the bench line and
# its CPU reference arm, the code-GEMM driver plain and then under ncu (full set, two shapes; launch list).
#   gpurun --timeout 560 -- 'bash profiles/run_final.sh r02'
R=${1:-r02}
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${R}_smoke.log 2>&1; echo "smoke rc=$?"
(time timeout 400 python -m pytest tests -m gpu -q -rs) > $O/${R}_gpu_tests_full.log 2>&1; echo "pytest rc=$?"
(time timeout 300 python bench.py) > $O/${R}_bench_n1.json 2> $O/${R}_bench_n1.err; echo "bench rc=$?"
timeout 200 python bench.py --impl reference > $O/${R}_bench_reference_n1.json 2>> $O/${R}_bench_n1.err; echo "ref rc=$?"
timeout 200 python profiles/prof_qgemm.py 128 --e4m3 > $O/${R}_prof_qgemm.jsonl 2> $O/${R}_prof_qgemm.err; echo "prof rc=$?"
for s in 1 3; do
  timeout 150 $NCU --set full --import-source on -k regex:qgemm_kernel -c 1 -f -o /tmp/${R}_qgemm_s$s \
      python profiles/prof_qgemm.py 128 --shape $s > /dev/null 2>&1
  python profiles/ncu_summary.py /tmp/${R}_qgemm_s$s.ncu-rep >> $O/${R}_ncu_full_qgemm.txt 2>&1
done
timeout 150 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:"qgemm|codes_" -c 40 --csv \
    --log-file $O/${R}_launches_qgemm.csv python profiles/prof_qgemm.py 128 --shape 1 > /dev/null 2>&1
tail -3 $O/${R}_gpu_tests_full.log; tail -c 400 $O/${R}_bench_n1.json
