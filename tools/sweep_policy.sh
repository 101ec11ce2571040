#!/bin/bash
# tile-width / branch-count sweep of the sampling graph (one bench line each): ms per 10-step call
for br in 2 4 8; do for bd in 64 128 192; do for bw in 128 192; do
  r=$(MDTB200_BRANCHES=$br MDTB200_BN_D=$bd MDTB200_BN_WIDE=$bw python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(round(j['ms_per_step'],3), j.get('gpu_launches'))")
  echo "branches=$br BN_D=$bd BN_WIDE=$bw -> $r"
done; done; done
