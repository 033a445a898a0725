"""Summarise one-kernel `ncu --set full --import-source on` reports (read here, no GPU):
key counters plus the SASS grouped into regions of equal execution count.
usage: python scripts/ncu_regions.py gpurun_out/prof.ncu-rep [out.md]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = []


def page(name):
    txt = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(txt.splitlines()))


raw = page("raw")
hdr, units, vals = raw[0], raw[1], raw[2]
WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "lts__t_requests_srcunit_tex_op_red.sum",
    "lts__t_sector_hit_rate.pct",
]
out.append("kernel: `%s`\n" % vals[hdr.index("Kernel Name")][:120])
out.append("| metric | value | unit |\n|---|---|---|")
for i, h in enumerate(hdr):
    if h in WANT or ("issue_stalled" in h and "per_issue_active" in h and float(vals[i] or 0) > 0.15):
        out.append("| %s | %s | %s |" % (h, vals[i], units[i]))

src = page("source")
shdr, data = src[1], []
for r in src[2:]:           # first kernel of the report only
    if len(r) < len(shdr):
        break
    data.append(r)
ia, isrc = shdr.index("Address"), shdr.index("Source")
ist, iex = shdr.index("Warp Stall Sampling (All Samples)"), shdr.index("Instructions Executed")
tot = sum(int(r[iex]) for r in data)
tots = sum(int(r[ist]) for r in data) or 1
base = int(data[0][ia], 16)
out.append("\n| SASS range | executions | instructions | % of executed | % of stall samples | first instruction |\n|---|---|---|---|---|---|")
cur, acc = None, []


def flush():
    if acc:
        out.append("| 0x%04x-0x%04x | %d | %d | %.2f | %.2f | `%s` |" % (
            int(acc[0][ia], 16) - base, int(acc[-1][ia], 16) - base, int(acc[0][iex]), len(acc),
            100 * sum(int(r[iex]) for r in acc) / tot, 100 * sum(int(r[ist]) for r in acc) / tots,
            acc[0][isrc].strip()[:44]))


for r in data:
    n = int(r[iex])
    if cur is None or abs(n - cur) > 0.02 * max(cur, 1):
        flush()
        acc, cur = [], n
    acc.append(r)
flush()
txt = "\n".join(out) + "\n"
if len(sys.argv) > 2:
    open(sys.argv[2], "a").write(txt)
print(txt)
