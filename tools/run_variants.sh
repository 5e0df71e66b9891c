#!/bin/bash
# On the GPU box: bench every variants/*.so (device leg only) and print the per-stage times of the C2 step and, unless
# --headline-only is passed, of the north-star (C3_full) step.
# usage: tools/run_variants.sh [extra bench.py args]
set -u
LIB=orb-slam-birdview_b200/liborbb200.so
cp $LIB /tmp/liborbb200.orig.so
for v in variants/*.so; do
  cp "$v" $LIB
  out=$(python bench.py --no-cpu-baseline --no-e2e --no-side-configs --steps 10 --warmup 3 "$@" 2>/dev/null | tail -1)
  echo "$(basename $v .so) $(echo "$out" | python -c '
import sys,json
d=json.loads(sys.stdin.read())
print("ms_per_step=%.4f"%d["ms_per_step"], {k: round(v,4) for k,v in d["stage_ms_per_step"].items()})
f=d.get("configs",{}).get("C3_full")
if f: print("   C3_full ms_per_step=%.4f"%f["ms_per_step"], {k: round(v,4) for k,v in f["stage_ms_per_step"].items()})
')"
done
cp /tmp/liborbb200.orig.so $LIB
