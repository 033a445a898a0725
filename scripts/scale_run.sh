#!/bin/bash
# usage: scripts/scale_run.sh N   (run under gpurun --gpus N)
N=$1
if [ "$N" = "1" ]; then
  python bench.py --no-cpu --steps 10 2>&1 | tail -1 > gpurun_out/scale_c5_n1.json
  python bench.py --workload coils --steps 50 2>&1 | tail -1 > gpurun_out/scale_coils_n1.json
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --no-cpu --steps 10 2>&1 | tail -1 > gpurun_out/scale_c5_n$N.json
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --no-cpu --steps 10 --sharding samples 2>&1 | tail -1 > gpurun_out/scale_c5samples_n$N.json
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --workload coils --steps 50 2>&1 | tail -1 > gpurun_out/scale_coils_n$N.json
fi
python - <<PY
import json
for w in ("c5", "coils"):
    d = json.loads(open("gpurun_out/scale_%s_n$N.json" % w).read())
    print(w, "N=$N", "ms/step %.3f" % d["ms_per_step"], "value %.4g" % d["value"], "e2e", d.get("e2e", {}).get("ms_per_step"))
PY
