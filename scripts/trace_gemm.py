"""Per-tile timeline of the tcgen05 GEMM (dev tool).  Builds an instrumented copy of libmmg (-DMMG_GEMM_TRACE: clock64() stamps
in the TMA / MMA / epilogue roles), runs the transformer-block GEMM shapes and prints where a tile period goes.
  build (CPU container):  python scripts/trace_gemm.py --build
  run (GPU box):          MMG_LIB=scripts/_build/libmmg_trace.so python scripts/trace_gemm.py"""
import argparse, ctypes, os, subprocess, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "scripts", "_build")
ap = argparse.ArgumentParser()
ap.add_argument("--build", action="store_true")
args = ap.parse_args()

if args.build:
    from muse_maskgit_pytorch_b200 import build as B
    B.build()
    os.makedirs(OUT, exist_ok=True)
    obj = os.path.join(OUT, "mmg_gemm_trace.o")
    subprocess.run([B.NVCC] + B.FLAGS + ["-DMMG_GEMM_TRACE", "-c", os.path.join(B.CSRC, "mmg_gemm.cu"), "-o", obj], check=True)
    objs = [os.path.join(B.OBJ, f) for f in sorted(os.listdir(B.OBJ)) if f.endswith(".o") and f != "mmg_gemm.o"] + [obj]
    subprocess.run([B.NVCC, "-shared", "-o", os.path.join(OUT, "libmmg_trace.so")] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"], check=True)
    print("built", os.path.join(OUT, "libmmg_trace.so"))
    sys.exit(0)

import torch
from muse_maskgit_pytorch_b200 import ops, _lib as L

lib = L.lib()
lib.mmg_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
NCTA, NT, NS = 160, 32, 10


def read():
    buf = np.zeros(NCTA * NT * NS, dtype=np.int64)
    assert lib.mmg_trace_read(buf.ctypes.data, buf.nbytes) == 0
    return buf.reshape(NCTA, NT, NS)


def report(name, fn, tiles_per_cta):
    lib.mmg_trace_clear()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t = read()[:148, :min(tiles_per_cta, NT)]
    pair = bool((t[1::2, :2, 3] == 0).all()) or os.environ.get("MMG_GEMM_PAIR") == "1"
    if pair:
        lead = t[0::2]                     # MMA slots are written by the leader CTA of each pair only
        t = t.copy(); t[1::2, :, 0:4] = lead[:, :, 0:4]
        name += " [CTA pairs]"
    n = t.shape[1]
    # slots: 0 mma tile start, 1 after tmem_empty wait, 2 sum full-bar wait, 3 after last commit, 4 epi before tmem_full wait,
    #        5 after, 6 release, 7 epi end, 8 producer sum empty wait, 9 producer tile end
    mid = slice(2, n - 1) if n > 4 else slice(0, n)
    period = np.diff(t[:, :, 3], axis=1)[:, 1:-1] if n > 3 else np.zeros((1, 1))
    f = lambda a: f"{np.median(a):8.0f}"
    print(f"--- {name}: tiles/CTA {n}")
    print(" tile period (MMA last-commit to last-commit)  ", f(period), "cycles")
    print(" tile period seen by the epilogue (tmem_full to tmem_full)", f(np.diff(t[:, :, 5], axis=1)[:, 1:-1] if n > 3 else np.zeros((1, 1))))
    print(" MMA: wait tmem_empty                           ", f((t[:, mid, 1] - t[:, mid, 0])))
    print(" MMA: wait full barriers (sum per tile)         ", f(t[:, mid, 2]))
    print(" MMA: issue span (after tmem wait -> last commit)", f(t[:, mid, 3] - t[:, mid, 1]))
    print(" EPI: wait tmem_full                            ", f(t[:, mid, 5] - t[:, mid, 4]))
    print(" EPI: tmem_full -> stage released               ", f(t[:, mid, 6] - t[:, mid, 5]))
    print(" EPI: tmem_full -> epilogue end                 ", f(t[:, mid, 7] - t[:, mid, 5]))
    print(" EPI: MMA last commit issued -> tmem_full seen  ", f(t[:, mid, 5] - t[:, mid, 3]))
    print(" TMA: wait empty barriers (sum per tile)        ", f(t[:, mid, 8]))
    print(" whole kernel (first MMA start -> last epi end) ", f(t[:, n - 1, 7] - t[:, 0, 0]))


dev = "cuda"
bf = torch.bfloat16
M = 32768
x = torch.randn(M, 512, device=dev).to(bf)
xf = torch.randn(M, 512, device=dev)
h = torch.randn(M, 1408, device=dev).to(bf)


def W(n, k):
    return (torch.randn(n, k, device=dev) * 0.03).to(bf)


# plain bf16 store at the qkv shape
w = W(1536, 512); out = torch.empty(M, 1536, device=dev, dtype=bf)
report("store bf16  32768x1536x512", lambda: ops.linear(x, w, out), 10)
# QKV epilogue
heads, n = 8, 256
q = torch.empty(M // n * heads, n, 64, device=dev, dtype=bf); k = torch.zeros(M // n * heads, 264, 64, device=dev, dtype=bf); v = torch.zeros_like(k)
qs = torch.ones(64, device=dev); nk = torch.zeros(heads, 64, device=dev, dtype=bf)
epi = ops.qkv_epilogue(bf, heads, n, q=q, k=k, v=v, q_scale=qs, k_scale=qs, key_off=1, null_k=nk, null_v=nk)
report("qkv         32768x1536x512", lambda: ops.linear(x, w, None, epilogue=ops.EPI_QKV, epi=epi), 10)
# residual fp32
w2 = W(512, 512)
report("resid f32   32768x512x512", lambda: ops.linear(x, w2, xf, epilogue=ops.EPI_RESIDUAL, resid=xf), 3)
# GEGLU
w1 = W(2816, 512); hh = torch.empty(M, 1408, device=dev, dtype=bf)
report("geglu       32768x2816x512", lambda: ops.linear(x, w1, hh, epilogue=ops.EPI_GEGLU), 19)
# ff2 residual
w3 = W(512, 1408)
report("resid f32   32768x512x1408", lambda: ops.linear(h, w3, xf, epilogue=ops.EPI_RESIDUAL, resid=xf), 3)
# logits
wl = W(65536, 512); e = torch.randn(10240, 512, device=dev).to(bf); lg = torch.empty(10240, 65536, device=dev)
report("logits f32  10240x65536x512", lambda: ops.linear(e, wl, lg), 32)
# long K reference
a8 = torch.randn(8192, 8192, device=dev).to(bf); w8 = W(8192, 8192); o8 = torch.empty(8192, 8192, device=dev, dtype=bf)
report("store bf16  8192^3", lambda: ops.linear(a8, w8, o8), 13)
