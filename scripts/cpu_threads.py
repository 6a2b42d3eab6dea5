"""How the CPU oracle scales with torch threads on this box (picks the thread count for the cpu_baseline)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import muse_maskgit_pytorch_b200 as M
from muse_maskgit_pytorch_b200 import t5
from oracle import muse_oracle as O
t5.T5_CONFIGS["synth-512"] = {"d_model": 512}
torch.manual_seed(0)
vae = M.VQGanVAE(**bench.VAE_CFG); tr = M.MaskGitTransformer(t5_name="synth-512", **bench.TR_CFG)
sd = {k: v.detach().float() for k, v in tr.state_dict().items()}; vsd = {k: v.detach().float() for k, v in vae.state_dict().items()}
print("cpu_count", os.cpu_count(), flush=True)
with torch.no_grad():
    for th in (8, 16, 32, 64, 128):
        if th > (os.cpu_count() or 1):
            break
        torch.set_num_threads(th)
        t_step, t_dec = bench.cpu_sample(O, sd, vsd, 1)
        print(f"threads {th:4d}: decode step {t_step:7.2f} s, vae decode {t_dec:7.2f} s -> {1 / (18 * t_step + t_dec):.4f} img/s", flush=True)
