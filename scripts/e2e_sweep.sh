#!/bin/bash
# e2e time of the bench workload for several host_chunks settings
for c in "$@"; do
  timeout 300 python bench.py --no-cpu --steps 6 --host-chunks $c 2>&1 | tail -1 > /tmp/e2e_$c.json
  python - "$c" <<'PY'
import sys, json
c = sys.argv[1]
d = json.loads(open("/tmp/e2e_%s.json" % c).read())
print("host_chunks", c, "device ms/step %.2f" % d["ms_per_step"], "e2e ms/step %.2f" % d["e2e"]["ms_per_step"])
PY
done
