#!/bin/bash
# Developer sweep: pipeline operating point at the driver's 20 steps and at 100 steps
for cfg in "4 6" "3 6" "3 5" "5 8" "4 8"; do set -- $cfg; for k in 20 100; do
  echo -n "streams $1 depth $2 steps $k: "; python bench.py --steps $k --warmup 5 --no-other-configs --no-cpu-baseline --decode-streams $1 --depth $2 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(round(d['value']), round(d['e2e']['value']), d['ms_per_step'])"; done; done
