#!/bin/bash
# development: time the seeded randomization batch under HS_PERM_* knobs ("FUSED=0" "TILE_PARTS=4,TILE_THREADS=512" ...)
for cfg in "$@"; do
  envs="HS_TILE_DEBUG=0"; IFS=',' read -ra kv <<< "$cfg"; for x in "${kv[@]}"; do envs="$envs HS_PERM_$x"; done
  echo "== $cfg"
  env $envs python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-dense --no-extra-configs > /tmp/perm_exp.out 2> /tmp/perm_exp.err
  python -c "
import json
try:
    d=json.loads(open('/tmp/perm_exp.out').readline()); print(d['ms_per_step'], d['config']['rank0_ms'], d.get('parity'))
except Exception as e: print('bench failed', e)"
  grep "tile dbg] prepare" /tmp/perm_exp.err | tail -n 2
  grep -v "tile dbg" /tmp/perm_exp.err | tail -n 4
done
