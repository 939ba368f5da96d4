#!/bin/bash
# profiles/sweep_c5_sizes.sh TAG -- BASELINE config 5: the colour-transform chain at 1-64 MP per image for every texel
# format, ~256 MP per launch (a batch of frames in one allocation, one kernel launch), JSON lines into
# gpurun_out/c5_sizes_TAG.jsonl.   usage: gpurun --timeout 900 -- 'bash profiles/sweep_c5_sizes.sh r01'
TAG=${1:-r01}
mkdir -p gpurun_out
out=gpurun_out/c5_sizes_$TAG.jsonl
: > $out
for size in 1024x1024 1920x1080 2048x2048 3840x2160 4096x4096 7680x4320 8192x8192; do
  w=${size%x*}; h=${size#*x}
  frames=$(( (268435456 + w * h - 1) / (w * h) ))
  for wl in c5_rgba8 c5_rgba16f c5_rgb10a2 c5_yuv420_rgba8 c5_yuv420_yuv420; do
    timeout 200 python bench.py --workload $wl --size $size --frames $frames --no-cpu --steps 10 --warmup 3 >> $out 2>> gpurun_out/c5_sizes_$TAG.err
  done
done
python - "$out" <<'PY'
import json, sys
rows = [json.loads(l) for l in open(sys.argv[1]) if l.strip().startswith("{")]
print("| image | frames/launch | " + " | ".join(sorted({r["config"]["workload"] for r in rows})) + " |")
sizes = []
for r in rows:
    k = (r["config"]["out_px_per_frame"], r["config"]["frames_per_step_per_gpu"])
    if k not in sizes: sizes.append(k)
for px, fr in sizes:
    cells = []
    for wl in sorted({r["config"]["workload"] for r in rows}):
        m = [r for r in rows if r["config"]["workload"] == wl and r["config"]["out_px_per_frame"] == px]
        cells.append("%.0f MP/s (%.2f)" % (m[0]["value"], m[0]["roofline"]["frac"]) if m else "-")
    print("| %.1f MP | %d | %s |" % (px / 1e6, fr, " | ".join(cells)))
PY
