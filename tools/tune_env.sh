#!/bin/bash
# development aid: time the fused step for several values of one environment knob
#   tools/tune_env.sh VAR v1 v2 ...
var=$1; shift
for v in "$@"; do
  echo "== $var=$v"
  env $var=$v python - <<'PY'
import sys, os, time, json
sys.path.insert(0, os.getcwd())
import hyperelasticsolver_b200 as H
def run(model, logn, steps, flux="hll"):
    n = 1 << logn
    if model == H.MPH30:
        eos = (H.Barton2009(), H.Barton2009()); Ql, Qr = H.initial_states(eos, 6)
    else:
        eos = H.Barton2009(); Ql, Qr = H.hyperelasticity.initial_states(eos, 1)
    Q0 = H.initial_condition(Ql, Qr, n)
    with H.Solver(eos, n, model=model) as sol:
        sol.upload(Q0)
        sol.advance(1e9, flux, 0.6, 1.0 / n, max_steps=3)
        t0 = time.perf_counter()
        sol.advance(1e9, flux, 0.6, 1.0 / n, max_steps=steps)
        dt = time.perf_counter() - t0
    print(json.dumps(dict(model="mph30" if model else "sp13", n=n, gcups=round(n * steps / dt / 1e9, 3))))
run(H.MPH30, 22, 10); run(H.SP13, 23, 30)
PY
done
