"""Turn the ncu captures brought back in gpurun_out/ into the tracked summaries under
profiles/ (run here, no GPU needed):  python scripts/summarize_profiles.py <tag>"""
import collections
import csv
import json
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
rep = "gpurun_out/prof_%s_final.ncu-rep" % tag
launches = "gpurun_out/launches_%sb.csv" % tag

out = ["# %s kernel summary (bench.py workload, B200, ncu --clock-control none)\n" % tag]

# ---- launch shares
rows = [r for r in csv.reader(open(launches)) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    agg.setdefault(r[ki][:100], []).append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
out.append("## Launch list (`ncu --metrics gpu__time_duration.sum -c 80 python bench.py --steps 2 --warmup 3 --no-cpu`)\n")
out.append("cold-cache, serialised per-launch times: compare shares.\n")
out.append("| kernel | launches | avg us | share |\n|---|---|---|---|")
for k, v in agg.items():
    out.append("| `%s` | %d | %.1f | %.1f %% |" % (k.replace("|", "/"), len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))

# ---- full metrics of the two interpolation kernels
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "lts__t_requests_srcunit_tex_op_red.sum",
    "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]
traffic = {}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    out.append("\n## `%s` (`ncu --set full`, one launch)\n" % name[:90])
    out.append("| metric | value | unit |\n|---|---|---|")
    vals = {}
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            out.append("| %s | %s | %s |" % (w, r[i], units[i]))
            vals[w] = (r[i], units[i])

    def to_bytes(v, u):
        v = float(v.replace(",", ""))
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]
    if "dram__bytes_read.sum" in vals:
        t = to_bytes(*vals["dram__bytes_read.sum"]) + to_bytes(*vals["dram__bytes_write.sum"])
        key = "adj_kernel_dram_bytes_per_launch" if "spread" in name else "fwd_kernel_dram_bytes_per_launch"
        traffic[key] = t
        out.append("| dram read+write per launch | %.3f | GB |" % (t / 1e9))
traffic["source"] = "profiles/%s_kernels_summary.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)" % tag
open("profiles/%s_kernels_summary.md" % tag, "w").write("\n".join(out) + "\n")
json.dump(traffic, open("profiles/traffic.json", "w"), indent=1)
print("\n".join(out[-62:]))
