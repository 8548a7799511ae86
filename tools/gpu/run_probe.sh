#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
export OC_B200_LIB=$PWD/gym_comm_b200/variants/liboc_b200_probe.so
for w in cfg4 cfg2; do for c in "" "--chain"; do echo "## $w $c"; python tools/probe_step.py $w $c 2>&1 | tail -16; done; done > gpurun_out/r2_probe_cfg4.txt
cat gpurun_out/r2_probe_cfg4.txt
