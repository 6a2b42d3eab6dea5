"""One eager (graph-free) generate() of the C3 workload between cudaProfilerStart/Stop, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ... python scripts/profile_step.py
usage: python scripts/profile_step.py [batch]"""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

b = int(sys.argv[1]) if len(sys.argv) > 1 else 64
with torch.no_grad():
    mg = bench.build_models(torch.device("cuda", 0))
    mg.use_cuda_graph = False
    mg.sampler_seed = 2
    te = bench.text_embeddings(b).cuda()
    mg.transformer.encode_text = lambda t: te
    mg.generate([""] * b, timesteps=bench.TIMESTEPS, cond_scale=bench.COND_SCALE)      # warm-up (packing, attributes)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    mg.generate([""] * b, timesteps=bench.TIMESTEPS, cond_scale=bench.COND_SCALE)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
