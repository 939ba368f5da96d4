#!/bin/bash
# profiles/sweep.sh TAG -- every bench workload once (1 GPU), JSON lines into gpurun_out/sweep_TAG.jsonl
# usage: gpurun --timeout 900 -- 'bash profiles/sweep.sh r01'
TAG=${1:-r01}
mkdir -p gpurun_out
out=gpurun_out/sweep_$TAG.jsonl
: > $out
timeout 300 python bench.py >> $out 2> gpurun_out/sweep_$TAG.err
for w in c2_inscribe c1_oklab c5_rgba8 c5_rgba16f c5_rgb10a2 c5_yuv420_yuv420 c5_yuv420_rgba8; do
  timeout 200 python bench.py --workload $w --no-cpu --steps 10 --warmup 3 >> $out 2>> gpurun_out/sweep_$TAG.err
done
for w in c3_affine_bilinear c3_affine_nearest; do
  timeout 200 python bench.py --workload $w --frames 2 --no-cpu --steps 10 --warmup 3 >> $out 2>> gpurun_out/sweep_$TAG.err
done
timeout 200 python bench.py --workload c4_fused --frames 64 --no-cpu --steps 10 --warmup 3 >> $out 2>> gpurun_out/sweep_$TAG.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 >> $out 2>> gpurun_out/sweep_$TAG.err
cat $out
