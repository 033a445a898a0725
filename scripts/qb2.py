"""Quick device timing of the C5 interpolation stages with a few option sets."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "scripts")
from quick_bench import radial3d, run
import numpy as np, torch
which = sys.argv[1] if len(sys.argv) > 1 else "big"
precision = "single"
if which.endswith("_double"):
    which, precision = which[:-7], "double"
sets = [eval(a) for a in sys.argv[2:]] or [{}]
if which == "big":
    om = radial3d(102944, 512); Nd=(256,)*3; Kd=(384,)*3
else:
    om = radial3d(12868, 256); Nd=(128,)*3; Kd=(192,)*3
for opts in sets:
    if precision == "double":
        om = om.astype(np.float64)
    run(which + " " + precision + " " + str(opts), Nd, Kd, om, 6, opts, precision)
