#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for u in 1 0; do
  echo "## OC_ROW_UNDO=$u"
  export OC_ROW_UNDO=$u
  python tools/time_rollout.py open-divider_tomato 2 10 500 2 65536            # cfg2: 92 floats
  python tools/time_rollout.py open-divider_tomato 2 9 500 2 65536             # 88 floats: grouped rows
  python tools/time_rollout.py open-divider_tomato 2 10 500 2 65536 blind1     # cfg2 with a blind partner
  python tools/time_rollout.py random-open-divider_salad_small_cramped 2 8 900 10 65536          # cfg4 level, nobody blind (96 floats, grouped)
  python tools/time_rollout.py random-open-divider_salad_small_cramped 2 8 900 10 65536 blind1
  python tools/time_rollout.py random-open-divider_salad_small_cramped 2 9 900 10 65536          # 100 floats: not grouped
done 2>&1 | grep -v Warning | tee gpurun_out/r2_undo_diag.txt
