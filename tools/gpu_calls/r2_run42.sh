#!/bin/bash
# L2 window on the benchmark graph (rows fit L2: does pinning them still pay?) -- variants of GVC_L2_WINDOW_MIN_MB / GVC_L2_WINDOW_MB
mkdir -p gpurun_out
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_win_$name.json 2>gpurun_out/r2_win_$name.err
}
run base_a GVC_L2_WINDOW_MB=48
run w48 GVC_L2_WINDOW_MIN_MB=0 GVC_L2_WINDOW_MB=48
run w68 GVC_L2_WINDOW_MIN_MB=0 GVC_L2_WINDOW_MB=68
run w32 GVC_L2_WINDOW_MIN_MB=0 GVC_L2_WINDOW_MB=32
run base_b GVC_L2_WINDOW_MB=48
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_win_*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1]); r=d["roofline"]
        print(f, "ms/step %.4f"%d["ms_per_step"], [round(x,4) for x in r["stage_ms"]], "frac %.3f fwd %.3f"%(r["frac"], r["forward_frac"]), "other", d.get("other_mode",{}).get("ms_per_step"))
    except Exception as e: print(f, "ERR", e)
PY
