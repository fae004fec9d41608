#!/bin/bash
# On the GPU box: time the c3 step with the in-tree library and with every variant under tools/_alt/.
#   tools/ab_run.sh [workload] ; output: one JSON object per variant in gpurun_out/ab_<variant>.json
cd "$(dirname "$0")/.."
wl=${1:-c3}
lib=steered-mixture-of-experts_b200/libsmoe_b200.so
cp $lib /tmp/base.so
for v in base $(ls tools/_alt/*.so 2>/dev/null); do
  if [ "$v" = base ]; then cp /tmp/base.so $lib; n=base; else cp $v $lib; n=$(basename $v .so); fi
  python tools/step_breakdown.py $wl > gpurun_out/ab_${wl}_$n.json 2> gpurun_out/ab_${wl}_$n.err || tail -3 gpurun_out/ab_${wl}_$n.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_${wl}_$n.json"))
print("$n", {k: d[k] for k in ("graph_step_ms","smoe_forward","smoe_loss","smoe_backward") if k in d})
PY
done
cp /tmp/base.so $lib
