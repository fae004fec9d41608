#!/bin/bash
# per-rank work of config 4 on 8 ranks for several block grids, emulated on one GPU (rank list: corner + interior)
for g in 8,1,1 1,4,2 2,2,2 1,2,4 2,4,1 1,1,8; do
  echo "grid $g"
  SMOE_BLOCK_GRID=$g timeout 400 python tools/emulate_shards.py c4 8 0,3,5 2>&1 | grep -v "^$" | cut -c1-330
done
