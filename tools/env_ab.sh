#!/bin/bash
# bench.py under a list of environment settings (one per argument, "-" = defaults); prints ms/step
for cfg in "$@"; do
  if [ "$cfg" = "-" ]; then e=""; else e="$cfg"; fi
  env $e timeout 400 python bench.py --no-cpu-baseline --no-gpu-baseline --steps 30 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%-40s %.3f ms/step  %.1f samples/s  e2e %.1f' % ('$cfg', d['ms_per_step'], d['value'], d['e2e']['value']))"
done
