#!/bin/bash
# development: time the tile kernels under the HS_TILE_* knobs given as arguments ("FILL=2,EXP=0" ...)
for cfg in "$@"; do
  envs=""; IFS=',' read -ra kv <<< "$cfg"; for x in "${kv[@]}"; do envs="$envs HS_TILE_$x"; done
  echo "== $cfg"; env $envs HS_TILE_DEBUG=1 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-dense 2>&1 >/dev/null | grep "tile dbg" | head -n 3 | sed 's/producer1.*//'
done
