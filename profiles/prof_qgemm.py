"""The integer-code layer product (csrc/qgemm_kernels.cu: TMA + tcgen05.mma kind::i8, TMEM accumulators) on ResNet-50's
stride-1 1x1 convolutions and its classifier at batch 128, channels-last, beside the library path the reference takes
(cuDNN convolution of the fake-quantised fp32 tensor, TF32 as torch enables by default, and strict fp32).

Per shape (m = batch*H*W rows, k = Cin, n = Cout): CUDA-event time of 20 calls after 5 warm-ups (working sets of the
big shapes are far beyond the 126 MB L2; the small ones are flagged `l2_resident`), algorithmic bytes
    code GEMM : m*k (codes) + n*k (weight codes) + m*n*4 (fp32 out)   [bf16 out: m*n*2]
    cuDNN     : m*k*4 + n*k*4 + m*n*4
and the code GEMM's bytes / time against the measured copy peak.  `codes_us` is the separate x -> code pass
(4 B read + 1 B written per element) the layer needs when its producer does not hand over codes already.
Also a full-size property check: code GEMM == strict-fp32 cuDNN on the dequantised tensors within 1e-5 of sum|terms|."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dlmc_quant_b200 import qgemm as Q  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6550.0

# (H = W, Cin, Cout) of torchvision ResNet-50's stride-1 1x1 convolutions, batch 128
SHAPES = [(56, 64, 64), (56, 64, 256), (56, 256, 64), (56, 256, 128), (28, 128, 512), (28, 512, 128), (28, 512, 256),
          (14, 256, 1024), (14, 1024, 256), (14, 1024, 512), (7, 512, 2048), (7, 2048, 512)]


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(iters):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / iters * 1e3      # us


def run(batch, hw, cin, cout, encoding, check):
    dev = torch.device("cuda")
    m, k, n = batch * hw * hw, cin, cout
    g = torch.Generator(device="cuda").manual_seed(hw * 1000 + cin)
    ca = torch.randint(0, 16, (m, k), device=dev, generator=g, dtype=torch.uint8)
    cw_i = torch.randint(-7, 8, (n, k), device=dev, generator=g, dtype=torch.int8)
    if encoding == Q.QGEMM_E4M3:
        a_b = ca.float().to(torch.float8_e4m3fn).view(torch.uint8)
        w_b = cw_i.float().to(torch.float8_e4m3fn).view(torch.uint8)
    else:
        a_b, w_b = ca, cw_i.view(torch.uint8)
    s_a = 0.11
    s_w = (torch.rand(n, device=dev, generator=g) * 0.01 + 0.002)
    alpha, beta = (s_a * s_w).float(), torch.zeros(n, device=dev)
    out32 = torch.empty(m, n, device=dev)
    out16 = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    rec = {"m": m, "k": k, "n": n, "encoding": "i8" if encoding == Q.QGEMM_I8 else "e4m3"}
    rec["qgemm_f32_us"] = timeit(lambda: Q.qgemm(a_b, w_b, alpha, beta, encoding=encoding, out=out32))
    rec["qgemm_bf16_us"] = timeit(lambda: Q.qgemm(a_b, w_b, alpha, beta, encoding=encoding, out=out16,
                                                  out_dtype=torch.bfloat16))
    by32 = m * k + n * k + m * n * 4
    by16 = m * k + n * k + m * n * 2
    rec["qgemm_f32_gbs"] = by32 / rec["qgemm_f32_us"] * 1e-3
    rec["qgemm_bf16_gbs"] = by16 / rec["qgemm_bf16_us"] * 1e-3
    rec["qgemm_f32_frac_of_copy_peak"] = rec["qgemm_f32_gbs"] / PEAK
    rec["qgemm_bf16_frac_of_copy_peak"] = rec["qgemm_bf16_gbs"] / PEAK
    rec["tops"] = 2.0 * m * n * k / rec["qgemm_f32_us"] * 1e-6
    rec["l2_resident"] = by32 < 100e6
    if encoding == Q.QGEMM_I8:
        # the library path on the same values: fake-quantised fp32 tensors, channels-last 1x1 convolution
        x = (ca.float() * s_a).view(batch, hw, hw, k).permute(0, 3, 1, 2)          # NCHW view, channels-last storage
        w = (cw_i.float() * s_w[:, None]).view(n, k, 1, 1).contiguous(memory_format=torch.channels_last)
        torch.backends.cudnn.allow_tf32 = True
        rec["cudnn_tf32_us"] = timeit(lambda: torch.nn.functional.conv2d(x, w))
        torch.backends.cudnn.allow_tf32 = False
        rec["cudnn_fp32_us"] = timeit(lambda: torch.nn.functional.conv2d(x, w))
        rec["cudnn_bytes"] = m * k * 4 + n * k * 4 + m * n * 4
        rec["speedup_vs_cudnn_tf32"] = rec["cudnn_tf32_us"] / rec["qgemm_f32_us"]
        xs = torch.empty(m, k, device=dev).copy_(ca)
        sc, of = torch.tensor([s_a], device=dev), torch.zeros(1, device=dev)
        rec["codes_us"] = timeit(lambda: Q.codes_forward(xs, sc, of, 0, 15, 1, 1e-4))
        rec["codes_gbs"] = m * k * 5 / rec["codes_us"] * 1e-3
        if check:
            ref = torch.nn.functional.conv2d(x, w).permute(0, 2, 3, 1).reshape(m, n)
            got = Q.qgemm(a_b, w_b, alpha, beta, encoding=encoding)
            mag = (15 * s_a) * (7 * s_w) * k
            rec["max_err_over_bound"] = float(((got - ref).abs() / mag).max())
            rec["property_ok"] = rec["max_err_over_bound"] <= 1e-5
    return rec


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    encs = [Q.QGEMM_I8] + ([Q.QGEMM_E4M3] if "--e4m3" in sys.argv else [])
    if "--shape" in sys.argv:                       # one shape only, few calls: the command profiled under ncu
        hw, cin, cout = SHAPES[int(sys.argv[sys.argv.index("--shape") + 1])]
        dev = torch.device("cuda")
        m = batch * hw * hw
        a = torch.randint(0, 16, (m, cin), device=dev, dtype=torch.uint8)
        w = torch.randint(-7, 8, (cout, cin), device=dev, dtype=torch.int8).view(torch.uint8)
        al, be = torch.full((cout,), 1e-3, device=dev), torch.zeros(cout, device=dev)
        out = torch.empty(m, cout, device=dev)
        for _ in range(3):
            Q.qgemm(a, w, al, be, out=out)
        torch.cuda.synchronize()
        print(json.dumps({"shape": [m, cin, cout], "us": timeit(lambda: Q.qgemm(a, w, al, be, out=out), 10, 2)}))
        return
    print(json.dumps({"device": torch.cuda.get_device_name(0), "copy_peak_gbs": PEAK, "batch": batch}))
    for enc in encs:
        for hw, cin, cout in SHAPES:
            try:
                rec = run(batch, hw, cin, cout, enc, check=True)
            except Exception as e:  # noqa: BLE001
                rec = {"shape": [hw, cin, cout], "error": f"{type(e).__name__}: {e}"[:300]}
            print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in rec.items()}), flush=True)
            torch.cuda.empty_cache()
        # the classifier: [batch, 2048] x [1000, 2048]
        try:
            dev = torch.device("cuda")
            a = torch.randint(0, 16, (batch, 2048), device=dev, dtype=torch.uint8)
            w = torch.randint(-7, 8, (1000, 2048), device=dev, dtype=torch.int8).view(torch.uint8)
            if enc == Q.QGEMM_E4M3:
                a = a.float().to(torch.float8_e4m3fn).view(torch.uint8)
                w = w.view(torch.int8).float().to(torch.float8_e4m3fn).view(torch.uint8)
            al, be = torch.full((1000,), 1e-3, device=dev), torch.zeros(1000, device=dev)
            us = timeit(lambda: Q.qgemm(a, w, al, be, encoding=enc))
            print(json.dumps({"fc": [batch, 1000, 2048], "encoding": "i8" if enc == Q.QGEMM_I8 else "e4m3",
                              "qgemm_f32_us": round(us, 2)}), flush=True)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"fc": "error", "error": str(e)[:300]}), flush=True)


if __name__ == "__main__":
    main()
