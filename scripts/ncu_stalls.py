"""Warp-stall and pipe breakdown of every launch in an .ncu-rep (ncu --set full): the top issue-stall reasons (warps per issue-active
cycle), pipe utilisations, issue-slot use, achieved occupancy, DRAM / L2 traffic.  One block per launch.
usage: python scripts/ncu_stalls.py file.ncu-rep [max_launches]"""
import csv, subprocess, sys
rep = sys.argv[1]
maxn = int(sys.argv[2]) if len(sys.argv) > 2 else 8
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
col = {k: i for i, k in enumerate(hdr)}


def val(r, k):
    try:
        return float(r[col[k]].replace(",", ""))
    except (KeyError, ValueError):
        return None


KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.avg.per_cycle_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]
for li, r in enumerate(rows[2:2 + maxn]):
    name = r[col["Kernel Name"]].split("(")[0].replace("void mmg::", "")[:60]
    print(f"=== [{li}] {name}  grid {r[col.get('launch__grid_size', 0)]} block {r[col.get('launch__block_size', 0)]}")
    for k in KEYS:
        if k in col and r[col[k]] != "":
            print(f"    {k:72s} {r[col[k]]} {units[col[k]]}")
    stalls = []
    for k, i in col.items():
        if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
            v = val(r, k)
            if v is not None:
                stalls.append((v, k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
    stalls.sort(reverse=True)
    print("    stall reasons (warps stalled per issue-active cycle): " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:8]))
