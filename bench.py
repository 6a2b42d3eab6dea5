"""bench.py — images/sec of MaskGit.generate() 256x256, 18 steps, CFG=3 (BASELINE.json metric, config C3), batch 64
sharded over N GPUs of one node (STRONG scaling: the global batch is fixed at 64, per-GPU batch = 64/N).

  python bench.py --gpus N --steps K --warmup W            # our arm (libmmg.so through the drop-in classes)
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm's CPU arm (oracle port) on host cores

One "step" = one full generate() call over the rank's shard (18 decode steps + VAE decode + (N>1) one NCCL all-gather of
the decoded images).  `value` is timed with the text embeddings already resident in HBM; `e2e` times the same public call
with the embeddings coming from pinned host memory and the images copied back to the host inside the timed region.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

GLOBAL_BATCH = 64
IMAGE, TIMESTEPS, COND_SCALE, TEXT_LEN = 256, 18, 3.0, 32
TR_CFG = dict(num_tokens=65536, seq_len=256, dim=512, depth=8, dim_head=64, heads=8, ff_mult=4)
VAE_CFG = dict(dim=256, codebook_size=65536)
# algorithmic dense work (SURVEY.md 8d / BASELINE.md 3): 36 forwards x 33.70 + decode 165.47 GFLOP per image
GFLOP_PER_IMAGE = 1378.7
ATTN_GFLOP_PER_IMAGE = 1.216 * 36


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops=d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0)), hbm=d.get("hbm_gbs", 6650.0), src="measured")
    return dict(tflops=1400.0, hbm=6650.0, src="fallback")


def text_embeddings(batch, seed=1):
    g = torch.Generator().manual_seed(seed)
    te = torch.randn((batch, TEXT_LEN, 512), generator=g)
    te[1::2, 24:] = 0.                      # padded positions on odd rows (exercises the context mask)
    return te


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    def __init__(self, index):
        super().__init__(daemon=True)
        self.rows, self.proc, self.index = [], None, index

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        try:
            sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
            if sm:
                out["sm_mhz"] = sm[len(sm) // 2]
                out["sm_max_mhz"] = float(self.rows[0][1])
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            out["reasons"] = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        except Exception:
            pass
        return out


def build_models(device, precision="bf16"):
    import muse_maskgit_pytorch_b200 as M
    from muse_maskgit_pytorch_b200 import t5
    t5.T5_CONFIGS["synth-512"] = {"d_model": 512}
    torch.manual_seed(0)
    vae = M.VQGanVAE(precision=precision, **VAE_CFG)
    tr = M.MaskGitTransformer(t5_name="synth-512", precision=precision, **TR_CFG)
    mg = M.MaskGit(image_size=IMAGE, transformer=tr.to(device), vae=vae.to(device)).to(device)
    return mg


def run_ours(args):
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    assert GLOBAL_BATCH % world == 0
    b = GLOBAL_BATCH // world
    from muse_maskgit_pytorch_b200 import _lib
    mg = build_models(device)
    mg.row_offset = rank * b                  # RNG keyed on the global sequence index -> ids independent of the GPU count
    mg.sampler_seed = 2
    te_all = text_embeddings(GLOBAL_BATCH)
    te_host = te_all[rank * b:(rank + 1) * b].contiguous().pin_memory()
    te_dev = te_host.to(device)
    texts = [""] * b
    gathered = torch.empty((GLOBAL_BATCH, 3, IMAGE, IMAGE), device=device) if world > 1 else None
    host_out = torch.empty((GLOBAL_BATCH if world > 1 else b, 3, IMAGE, IMAGE)).pin_memory()

    def step(e2e):
        if e2e:
            mg.transformer.encode_text = lambda t: te_host.to(device, non_blocking=True)
        else:
            mg.transformer.encode_text = lambda t: te_dev
        images = mg.generate(texts, timesteps=TIMESTEPS, cond_scale=COND_SCALE, temperature=1., topk_filter_thres=0.9)
        if world > 1:
            dist.all_gather_into_tensor(gathered, images)
            images = gathered
        if e2e:
            host_out.copy_(images, non_blocking=True)
        return images

    def timed(e2e, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(steps):
            step(e2e)
        t1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([t0.elapsed_time(t1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(ms) / steps, (_lib.launch_count() - l0) // steps

    for _ in range(max(args.warmup, 3)):
        step(False)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms, launches = timed(False, args.steps)
    clocks = sampler.stop() if sampler else None
    step(True)
    ms_e2e, _ = timed(True, args.steps)

    line = None
    if rank == 0:
        pk = peaks()
        roof = kernel_roofline(mg, texts, te_dev, pk)
        value = GLOBAL_BATCH / (ms / 1e3)
        line = {
            "metric": "images/sec MaskGit.generate() 256x256 18-step CFG=3", "value": round(value, 2), "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms, 3),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "C3: MaskGit.generate() 256x256, 18 steps, cond_scale=3, top-k 0.9, global batch 64 "
                                   "(batch-sharded, one NCCL all-gather of images per call); transformer dim512 depth8 V65536, VQGanVAE dim256; "
                                   "random-init weights, pre-computed T5 embeddings (32 positions)",
                       "global_batch": GLOBAL_BATCH, "per_gpu_batch": b, "parallelism": f"dp{world}",
                       "l2": "inputs larger than L2: every decode step streams > 2 GB of logits"},
            "e2e": {"value": round(GLOBAL_BATCH / (ms_e2e / 1e3), 2), "unit": "images/s",
                    "h2d_bytes_per_step": te_host.numel() * 4, "d2h_bytes_per_step": host_out.numel() * 4},
            "gpu_launches": (launches if launches > 0 else roof["launches_per_step_all_kernels"]) * args.steps,   # graph replays re-issue the captured kernel nodes
            "clocks": clocks,
            "roofline": roof,
            "dense_flop_frac_of_peak": round(value * GFLOP_PER_IMAGE / 1e3 / world / pk["tflops"], 4),
            "attention_gemm_roofline_frac": round(value * ATTN_GFLOP_PER_IMAGE / 1e3 / world / pk["tflops"], 5),
            "peaks": pk,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(mg)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))


def kernel_roofline(mg, texts, te_dev, pk):
    """One extra generate() with every libmmg call bracketed by CUDA events on the launching stream: duration and
    algorithmic FLOPs of the dominant kernel (the tcgen05 GEMM behind mmg_linear / mmg_conv*)."""
    from muse_maskgit_pytorch_b200 import _lib
    rec = []
    orig = _lib.call

    def prof_call(name, a, stream=None):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig(name, a, stream)
        e1.record()
        fl = 0.0
        if name == "mmg_linear":
            fl = 2.0 * a.M * a.N * a.K
        elif name == "mmg_conv2d":
            taps = {0: 1, 1: 9, 2: 16, 3: 25}[a.kind]
            s = 2 if a.kind == 2 else 1
            fl = 2.0 * a.B * (a.H // s) * (a.W // s) * a.Cout * taps * a.Cin
        elif name == "mmg_conv_transpose2d":
            fl = 2.0 * a.B * a.H * a.W * a.Cout * 16 * a.Cin
        rec.append((name, e0, e1, fl))

    _lib.call = prof_call
    graph_mode = mg.use_cuda_graph
    try:
        mg.use_cuda_graph = False              # eager pass: every kernel node of the replayed graph issued one by one
        mg.transformer.encode_text = lambda t: te_dev
        mg.generate(texts, timesteps=TIMESTEPS, cond_scale=COND_SCALE)
        torch.cuda.synchronize()
    finally:
        _lib.call = orig
        mg.use_cuda_graph = graph_mode
    tot = {}
    for name, e0, e1, fl in rec:
        d = tot.setdefault(name, [0.0, 0.0, 0])
        d[0] += e0.elapsed_time(e1); d[1] += fl; d[2] += 1
    all_ms = sum(v[0] for v in tot.values())
    gemm = [tot[k] for k in ("mmg_linear", "mmg_conv2d", "mmg_conv_transpose2d") if k in tot]
    g_ms, g_fl, g_n = sum(v[0] for v in gemm), sum(v[1] for v in gemm), sum(v[2] for v in gemm)
    achieved = g_fl / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    return {"bound": "tensor", "kernel": "tc_gemm_kernel (tcgen05; mmg_linear + mmg_conv2d + mmg_conv_transpose2d)",
            "achieved": round(achieved, 1), "peak": pk["tflops"], "unit": "TFLOP/s", "frac": round(achieved / pk["tflops"], 4),
            "traffic": None, "launches": g_n, "launches_per_step_all_kernels": len(rec), "avg_launch_us": round(1e3 * g_ms / max(g_n, 1), 2), "share_of_step": round(g_ms / max(all_ms, 1e-9), 3),
            "by_entry_point_ms": {k: round(v[0], 3) for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0])}, "peak_source": pk["src"]}


def oracle_setup(mg):
    from oracle import muse_oracle as O
    sd = {k: v.detach().float().cpu() for k, v in mg.transformer.state_dict().items()}
    vsd = {k: v.detach().float().cpu() for k, v in mg.vae.state_dict().items()}
    return O, sd, vsd


def cpu_sample(O, sd, vsd, decode_steps=2, images=1, seed=2):
    """Bounded sample of the reference algorithm on the host cores: `decode_steps` of the 18 decode steps (every step costs the
    same on the CPU: two full forwards + the sampling tail over all n x V logits) plus one VAE decode, for `images` images.
    Returns (seconds per decode step, seconds per VAE decode)."""
    te = text_embeddings(GLOBAL_BATCH)[:images]
    g = torch.Generator().manual_seed(seed)
    noise = lambda step, shape: torch.rand(shape, generator=g)
    t0 = time.perf_counter()
    ids = O.generate_ids(sd, dict(heads=8, depth=8), te, (IMAGE // 16) ** 2, TR_CFG["num_tokens"], noise, timesteps=TIMESTEPS,
                         cond_scale=COND_SCALE, max_steps=decode_steps)
    t1 = time.perf_counter()
    ids = ids.clamp(max=TR_CFG["num_tokens"] - 1)          # still-masked positions of the truncated loop -> any valid code
    O.vae_decode_from_ids(vsd, ids.view(images, IMAGE // 16, IMAGE // 16), 16)
    t2 = time.perf_counter()
    return (t1 - t0) / decode_steps, t2 - t1


def cpu_threads():
    return min(os.cpu_count() or 1, int(os.environ.get("MMG_CPU_THREADS", "32")))


def cpu_baseline(mg, decode_steps=2):
    """The reference algorithm (oracle port) on this box's host cores, bounded sample of the same workload."""
    O, sd, vsd = oracle_setup(mg)
    cores = cpu_threads()
    torch.set_num_threads(cores)
    with torch.no_grad():
        t_step, t_dec = cpu_sample(O, sd, vsd, decode_steps)
    total = TIMESTEPS * t_step + t_dec
    return {"value": round(1.0 / total, 5), "unit": "images/s", "cores": cores, "kind": "port",
            "sample": f"1 image, full config (CFG=3, V=65536, 256x256, fp32 torch-CPU oracle): {decode_steps} of {TIMESTEPS} decode steps timed "
                      f"({t_step:.2f} s each) + VAE decode ({t_dec:.2f} s); images/s = 1 / ({TIMESTEPS} x step + decode)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    # CPU arm: build the same modules on the CPU only to draw the same default-init weights (no CUDA involved)
    import muse_maskgit_pytorch_b200 as M
    from muse_maskgit_pytorch_b200 import t5
    from oracle import muse_oracle as O
    t5.T5_CONFIGS["synth-512"] = {"d_model": 512}
    torch.manual_seed(0)
    vae = M.VQGanVAE(**VAE_CFG)
    tr = M.MaskGitTransformer(t5_name="synth-512", **TR_CFG)
    sd = {k: v.detach().float() for k, v in tr.state_dict().items()}
    vsd = {k: v.detach().float() for k, v in vae.state_dict().items()}
    cores = cpu_threads()
    torch.set_num_threads(cores)
    vals = []
    with torch.no_grad():
        for _ in range(min(args.warmup, 1)):
            cpu_sample(O, sd, vsd, 1)
        t = time.perf_counter()
        for _ in range(args.steps):                      # one bench "step" = one bounded sample (2 decode steps + VAE decode of 1 image)
            t_step, t_dec = cpu_sample(O, sd, vsd, 2)
            vals.append(1.0 / (TIMESTEPS * t_step + t_dec))
        dt = time.perf_counter() - t
    v = sum(vals) / len(vals)
    print(json.dumps({
        "impl": "reference", "metric": "images/sec MaskGit.generate() 256x256 18-step CFG=3", "value": round(v, 5), "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": round(1e3 * dt / args.steps, 1),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C3 on host cores: reference algorithm (oracle port of muse_maskgit_pytorch.py:493-621), same model config; "
                               "each step times 2 of the 18 decode steps + the VAE decode of 1 image and extrapolates to a full generate()",
                   "global_batch": 1, "parallelism": "cpu"},
        "cpu_baseline": {"value": round(v, 5), "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": "1 image per step: 2 of 18 decode steps + VAE decode timed, images/s = 1 / (18 x step + decode)"},
        "e2e": {"value": round(v, 5), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--global-batch", type=int, default=GLOBAL_BATCH, help="experiments only; the benchmark config is 64")
    args = ap.parse_args()
    GLOBAL_BATCH = args.global_batch
    if args.impl == "reference":
        run_reference(args)
    else:
        with torch.no_grad():
            run_ours(args)
