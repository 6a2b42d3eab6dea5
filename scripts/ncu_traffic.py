"""Aggregate an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list of one eager generate()
by kernel (count, time, DRAM bytes read / written) and write the per-launch DRAM traffic of the dominant kernel (the tcgen05 GEMM family) to
profiles/r2_traffic.json, which bench.py reports as roofline.traffic.
usage: python scripts/ncu_traffic.py launches.csv [out.json] > summary.txt"""
import collections
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]
ki, mi, vi, ui, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}
per = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", "")) * SCALE.get(r[ui], 1)
    except ValueError:
        continue
    d = per.setdefault(r[idi], {"name": r[ki]})
    d[r[mi]] = v
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d in per.values():
    name = d["name"].split("(")[0].replace("void ", "").replace("mmg::", "")[:64]
    a = agg[name]
    a[0] += 1; a[1] += d.get("gpu__time_duration.sum", 0.0); a[2] += d.get("dram__bytes_read.sum", 0.0); a[3] += d.get("dram__bytes_write.sum", 0.0)
tot_t = sum(a[1] for a in agg.values())
print(f"total {tot_t / 1e6:.3f} ms over {sum(a[0] for a in agg.values())} launches; DRAM read {sum(a[2] for a in agg.values()) / 1e9:.2f} GB, written {sum(a[3] for a in agg.values()) / 1e9:.2f} GB")
print(f"{'kernel':66s} {'n':>5s} {'ms':>9s} {'share':>7s} {'avg us':>8s} {'rd GB':>8s} {'wr GB':>8s} {'GB/s':>8s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
    print(f"{k:66s} {a[0]:5d} {a[1] / 1e6:9.3f} {a[1] / tot_t:7.1%} {a[1] / a[0] / 1e3:8.1f} {a[2] / 1e9:8.2f} {a[3] / 1e9:8.2f} {(a[2] + a[3]) / max(a[1], 1):8.0f}")
gemm = [a for k, a in agg.items() if k.startswith("tc_gemm_kernel") or k.startswith("tc_logits_kernel")]
n = sum(a[0] for a in gemm)
out = {"tc_gemm_dram_bytes_per_launch": round(sum(a[2] + a[3] for a in gemm) / max(n, 1)), "tc_gemm_launches": n,
       "tc_gemm_dram_read_gb": round(sum(a[2] for a in gemm) / 1e9, 3), "tc_gemm_dram_write_gb": round(sum(a[3] for a in gemm) / 1e9, 3),
       "source": f"{sys.argv[1].split('/')[-1]}: ncu dram__bytes_read.sum + dram__bytes_write.sum over the tcgen05 GEMM launches of one eager generate() (batch 64), per launch"}
# share of the fused logits + sampling entry point (mmg_logits_fused) that is tcgen05 GEMM time: its logits GEMM and the 4096-column sample GEMM
# (the only fp32 TMA-store GEMM of a bf16 generate()) against the threshold / finisher / fallback-sampler kernels of the same call
tsum = lambda pred: sum(a[1] for k, a in agg.items() if pred(k))
fused_gemm = tsum(lambda k: k.startswith("tc_logits_kernel") or k.startswith("tc_gemm_kernel<256, 0, 1, 3>") or k.startswith("tc_gemm_kernel<256, 0, 0, 3>"))
fused_other = tsum(lambda k: k.startswith("logits_finish_kernel") or k.startswith("logits_threshold") or k.startswith("logits_sample"))
if fused_gemm > 0:
    out["logits_fused_gemm_share"] = round(fused_gemm / (fused_gemm + fused_other), 4)
    out["logits_fused_note"] = "ncu time of tc_logits_kernel + the sample GEMM / (those + logits_threshold + logits_finish + fallback sampler) in the same launch list"
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out))
