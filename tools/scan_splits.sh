#!/bin/bash
# usage: scratch/scan_splits.sh "6 10 14" [workload]
for NS in $1; do
  SMOE_SPLITS=$NS timeout 300 python bench.py --workload ${2:-c3} --steps 60 --no-cpu --no-dense 2>/dev/null > /tmp/line.json
  python - "$NS" <<'PY'
import sys, json
j = json.load(open('/tmp/line.json')); r = j["roofline"]
print("NS", sys.argv[1], round(j["ms_per_step"], 4), round(r["forward_ms"], 4), round(r["backward_ms"], 4), round(r["other_ms"], 4))
PY
done
