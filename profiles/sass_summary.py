"""Per-kernel SASS / resource summary of libdlmcq.so (runs on the CPU box: cuobjdump only).
    python profiles/sass_summary.py > profiles/r02_sass_summary.txt
Columns: registers, static shared memory, local (spill) bytes, SASS instruction count, and the counts of the
instructions that show how a kernel moves and computes: 128-bit global loads / stores (LDG.E.128 incl. .NA / .CONSTANT
variants, STG.E.128), bulk async copies (UBLKCP = cp.async.bulk, the TMA 1-D path) and mbarrier waits (SYNCS), packed
fp32 math (FFMA2 / FMUL2 / FADD2), reciprocal (MUFU.RCP: the one per-thread reciprocal of the exact fast division),
shared-memory atomics (ATOMS), warp shuffles; for the one contraction on the path (the integer-code GEMM) the tensor
path: UTMALDG (TMA tensor loads), UTC*MMA (tcgen05.mma: UTCIMMA = kind::i8, UTCQMMA = kind::f8f6f4), UTCBAR
(tcgen05.commit), LDTM (tcgen05.ld from tensor memory)."""
import collections
import os
import re
import subprocess
import sys

so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dlmc_quant_b200", "libdlmcq.so")
res = subprocess.run(["cuobjdump", "--dump-resource-usage", so], capture_output=True, text=True).stdout
usage = {}
name = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        name = m.group(1)
        continue
    if name and "REG:" in line:
        usage[name] = {k: int(v) for k, v in re.findall(r"(REG|SHARED|LOCAL|STACK):(\d+)", line)}
        name = None
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
counts = collections.defaultdict(collections.Counter)
name = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        op = m.group(1)
        c = counts[name]
        c["n"] += 1
        if op.startswith("LDG") and ".128" in op:
            c["LDG.128"] += 1
        elif op.startswith("LDG"):
            c["LDG.other"] += 1
        if op.startswith("STG") and ".128" in op:
            c["STG.128"] += 1
        elif op.startswith("STG"):
            c["STG.other"] += 1
        for key in ("UBLKCP", "SYNCS", "FFMA2", "FMUL2", "FADD2", "MUFU.RCP", "ATOMS", "SHFL", "UTMALDG", "UTCBAR", "LDTM"):
            if op.startswith(key):
                c[key] += 1
        if re.match(r"UTC[A-Z]*MMA", op):
            c["UTC*MMA"] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
pretty = dict(zip(counts, (d.replace("(anonymous namespace)::", "") for d in demangle)))
cols = ["LDG.128", "LDG.other", "STG.128", "STG.other", "UBLKCP", "SYNCS", "FFMA2", "FMUL2", "FADD2", "MUFU.RCP", "ATOMS", "SHFL",
        "UTMALDG", "UTC*MMA", "UTCBAR", "LDTM"]
tc = [pretty[k] for k in counts if counts[k]["UTC*MMA"]]
print("libdlmcq.so: %d kernels (sm_100a); tensor-core instructions (tcgen05.mma) only in the %d instantiations of the "
      "integer-code GEMM: %s" % (len(counts), len(tc), "; ".join(re.sub(r"\(.*", "", t) for t in sorted(tc))))
print("%-92s %4s %6s %5s %5s " % ("kernel", "REG", "SHARED", "LOCAL", "SASS") + " ".join("%9s" % c for c in cols))
for k in sorted(counts, key=lambda k: pretty[k]):
    p = re.sub(r"\(.*", "", pretty[k]).replace("dlmcq::", "").replace("void ", "").replace("__nv_bfloat16", "bf16")
    p = p.replace("(bool)", "").replace("(int)", "")
    u = usage.get(k, {})
    print("%-92s %4d %6d %5d %5d " % (p[:92], u.get("REG", -1), u.get("SHARED", -1), u.get("LOCAL", -1), counts[k]["n"]) +
          " ".join("%9d" % counts[k][c] for c in cols))
