#!/bin/bash
# usage: tools/run_scale.sh N WORKLOAD RAYS STEPS WARMUP [extra bench.py flags]   -> gpurun_out/r2c_scale_<workload>_n<N>.json
N=$1; WL=$2; RAYS=$3; STEPS=$4; WARM=$5; shift 5
OUT=gpurun_out/r2c_scale_${WL}_n${N}
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --workload $WL --rays $RAYS --steps $STEPS --warmup $WARM "$@" > $OUT.json 2> $OUT.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload $WL --rays $RAYS --steps $STEPS --warmup $WARM "$@" > $OUT.json 2> $OUT.err
fi
echo "$WL N=$N rc=$?"; python - <<PY
import json
try:
    d = json.load(open("$OUT.json"))
    e = d.get("e2e") or {}
    print("  value %.3e tests/s  ms/step %.1f  e2e %.3e (%.1f ms)  scaling %s  balance %s" % (d["value"], d["ms_per_step"], e.get("value", 0), e.get("ms_per_step", 0), d["scaling"], d.get("rays_per_iteration_max_over_mean")))
    print("  e2e breakdown", (e.get("breakdown_per_step_rank0") or [None])[-1], " weak", (d.get("weak_scaling") or {}).get("value"))
except Exception as ex:
    print("  no line:", ex); import subprocess; print(open("$OUT.err").read()[-1500:])
PY
