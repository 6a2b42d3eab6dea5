"""Condense an .ncu-rep (ncu --set full) into the handful of metrics the roofline discussion uses, one line per launch.
usage: python scripts/ncu_summary.py file.ncu-rep [stride offset]   (stride/offset pick e.g. every 3rd launch = the timed one)"""
import csv, subprocess, sys
rep = sys.argv[1]
stride = int(sys.argv[2]) if len(sys.argv) > 2 else 1
offset = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = [("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
        ("gpu__time_duration.sum", "time"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_active_%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"), ("sm__inst_executed.sum.pct_of_peak_sustained_elapsed", "issue_%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
        ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"), ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2_to_sm_bytes"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_%")]
idx = [(hdr.index(k), n) for k, n in want if k in hdr]
for li, r in enumerate(rows[2:]):
    if li % stride != offset:
        continue
    parts = []
    for i, n in idx:
        v = r[i]
        if n == "kernel":
            v = v.split("(")[0].replace("void mmg::", "")[:48]
        elif units[i]:
            v = f"{v} {units[i]}"
        parts.append(f"{n}={v}")
    print(f"[{li}] " + "  ".join(parts))
