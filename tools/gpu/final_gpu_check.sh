#!/bin/bash
# Round-end GPU check: gpurun -- 'bash tools/gpu/final_gpu_check.sh [notests]'
# parity suite, smoke, the bench lines kept under profiles/, and the ncu launch list of the bench command.
mkdir -p gpurun_out
if [ "$1" != "notests" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
fi
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo bench rc=$?
for w in cfg3 cfg4 cfg5; do python bench.py --workload $w --no-cpu-baseline --steps 5000 > gpurun_out/bench_$w.json 2>/dev/null; done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null
B="python bench.py --steps 256 --warmup 4 --no-cpu-baseline --no-e2e"
$B > /dev/null 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_ll.log 2>&1; echo ncu rc=$?
