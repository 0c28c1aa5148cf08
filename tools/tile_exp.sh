#!/bin/bash
# development: time the tile kernels under the HS_TILE_* knobs given as arguments ("FORM=2,DEBUG=40" ...)
for cfg in "$@"; do
  envs="HS_TILE_DEBUG=0"; IFS=',' read -ra kv <<< "$cfg"; for x in "${kv[@]}"; do envs="$envs HS_TILE_$x"; done
  echo "== $cfg"; env $envs python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-dense --no-extra-configs 2>&1 >/dev/null | grep "tile dbg" | grep -E "prepare|CTAs:" | tail -n 4 | cut -c1-150
done
