#!/bin/bash
# ncu --set full of one workload's kernel for two builds of the library + burst timings (no profiler) of both.
# usage: profiles/ncu_ab.sh WORKLOAD KERNEL_REGEX base.so TAG
WL=$1; KR=$2; BASE=$3; TAG=$4
for which in base new; do
  if [ $which = base ]; then export ZOS_CUDA_LIB=$PWD/$BASE; else unset ZOS_CUDA_LIB; fi
  python bench.py --workload $WL --no-cpu --no-e2e --min-seconds 0.01 --steps 20 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$which burst', d['roofline']['frac'], d['clocks']['sm_mhz'], d['roofline']['kernel_ms_avg'])"
  ncu --set full --clock-control none --import-source on -k regex:$KR -s 6 -c 1 -f -o gpurun_out/${TAG}_$which python bench.py --workload $WL --no-cpu --no-e2e --min-seconds 0.001 --steps 3 --warmup 3 > /dev/null 2>&1
done
ls -la gpurun_out/${TAG}_*.ncu-rep
