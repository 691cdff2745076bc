#!/bin/bash
# Sweep of the overlapped-projection knobs: one short bench run per setting, one summary line each.
OUT=gpurun_out/r2
mkdir -p $OUT
run() { # label env...
    local label=$1; shift
    local extra=""
    case "$label" in *USEPIPE*) extra="--pipeline";; esac
    env "$@" timeout -s KILL 200 python bench.py --no-cpu-baseline --no-side-configs --steps ${STEPS:-5} $extra > $OUT/exp_$label.json 2> $OUT/exp_$label.err
    python - "$label" $OUT/exp_$label.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    k = d["kernels"]
    print(f"{sys.argv[1]:28s} ms/step {d['ms_per_step']:7.3f}  e2e {d['e2e']['value']/1e6:6.1f} M/s  " +
          "  ".join(f"{n}={k[n]['ms_per_step']:.2f}" for n in sorted(k) if n.startswith("tc_")), "loss", round(d["metrics_of_timed_steps"]["loss"], 6))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
for spec in "$@"; do
    label=$(echo "$spec" | sed -e 's#[^ ]*/##g' | tr ' =' '__')
    run "$label" $spec
done
