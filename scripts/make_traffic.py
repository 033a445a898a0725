"""profiles/traffic.json from `ncu --set full` captures of the two interpolation kernels (read
here, no GPU): dram__bytes_read.sum + dram__bytes_write.sum per launch, stamped with the hash of
the kernel sources the capture was taken on (bench.py quotes the numbers only for that hash).
usage: python scripts/make_traffic.py <adjoint.ncu-rep> <forward.ncu-rep>"""
import csv
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def dram_bytes(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    tot = 0.0
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(name)
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
        tot += float(vals[i].replace(",", "")) * scale
    return tot, vals[hdr.index("Kernel Name")][:80]


adj, adj_name = dram_bytes(sys.argv[1])
fwd, fwd_name = dram_bytes(sys.argv[2])
out = {"adj_kernel_dram_bytes_per_launch": adj, "fwd_kernel_dram_bytes_per_launch": fwd,
       "adj_kernel": adj_name, "fwd_kernel": fwd_name, "csrc_sha": bench.csrc_sha(),
       "source": "ncu --set full --clock-control none, one launch each on the bench workload "
                 "(scripts/ab_adj.py): dram__bytes_read.sum + dram__bytes_write.sum; reports "
                 + os.path.basename(sys.argv[1]) + ", " + os.path.basename(sys.argv[2])}
json.dump(out, open(os.path.join(bench.ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(out)
