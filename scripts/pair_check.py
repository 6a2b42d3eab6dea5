"""CTA-pair GEMM (cta_group::2) smoke check against torch, run under `timeout` (dev tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from muse_maskgit_pytorch_b200 import ops

torch.manual_seed(0)
for (M, N, K) in [(256, 256, 64), (512, 512, 512), (384, 768, 192), (32768, 1536, 512), (10240, 65536, 512), (4100, 512, 1408)]:
    a = (torch.randn(M, K, device="cuda")).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.float32)
    ops.linear(a, w, out)
    torch.cuda.synchronize()
    if M * N <= 2 ** 26:
        ref = a.float() @ w.float().t()
        err = (out - ref).abs().max().item()
    else:   # spot-check rows
        idx = torch.randint(0, M, (64,), device="cuda")
        ref = a[idx].float() @ w.float().t()
        err = (out[idx] - ref).abs().max().item()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(5):
        ops.linear(a, w, out)
    ev1.record(); torch.cuda.synchronize()
    us = ev0.elapsed_time(ev1) * 1000 / 5
    print(f"M={M} N={N} K={K}  max|err|={err:.3e}  {us:8.1f} us  {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s", flush=True)
    assert err < 2e-2 * max(1.0, K ** 0.5 * 0.05), err
print("pair_check ok")
