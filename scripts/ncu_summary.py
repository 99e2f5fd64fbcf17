"""Summarise an .ncu-rep (ncu --set full) as text: one block per launch with the metrics DESIGN.md quotes."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio"]
ki = hdr.index("Kernel Name")
idx = [(w, hdr.index(w)) for w in want if w in hdr]
def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None
print(f"# {rep}: {len(rows) - 2} launches (ncu --set full --clock-control none)")
agg = {}
for n, r in enumerate(rows[2:]):
    name = r[ki].split("(")[0]
    print(f"\n[{n}] {name}")
    for w, i in idx:
        print(f"    {w:85s} {r[i]:>14s} {units[i]}")
    t = num(r[hdr.index('gpu__time_duration.sum')])
    rd, wr = num(r[hdr.index('dram__bytes_read.sum')]), num(r[hdr.index('dram__bytes_write.sum')])
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    b = (rd or 0) * scale.get(units[hdr.index('dram__bytes_read.sum')], 1) + (wr or 0) * scale.get(units[hdr.index('dram__bytes_write.sum')], 1)
    tu = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}.get(units[hdr.index('gpu__time_duration.sum')], 1e-6)
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1; a[1] += (t or 0) * tu; a[2] += b
print("\n# per kernel: launches, total time (ms), DRAM bytes read+written (GB), DRAM GB/s over the kernel time")
for name, (c, t, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:40s} n={c:4d} time={t * 1e3:9.3f} ms  dram={b * 1e-9:8.3f} GB  {b / t * 1e-9 if t else 0:8.1f} GB/s")
