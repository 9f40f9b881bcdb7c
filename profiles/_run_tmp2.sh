cat > /tmp/k.py <<'PY'
import torch, sys
sys.path.insert(0, '.')
from dlmc_quant_b200 import functional as F
n=1<<26
x=torch.relu(torch.randn(n, device='cuda'))*2
k_hi=int(0.9999*n); k_lo=n+1-k_hi
print(F.kth_values(x,[k_lo,k_hi],fast=True))
print(F.kth_values(x,[k_hi],abs_input=True,fast=True))
PY
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,launch__registers_per_thread --clock-control none --csv --log-file gpurun_out/r02_kth_launches.csv python /tmp/k.py > /dev/null 2>&1
python - <<'PY'
import csv
lines=[l for l in open('gpurun_out/r02_kth_launches.csv') if not l.startswith('==')]
for row in csv.DictReader(lines):
    if 'kth' in row['Kernel Name'] or 'radix' in row['Kernel Name']:
        print(row['Kernel Name'][:60], row['Grid Size'], row['Metric Name'], row['Metric Value'])
PY
