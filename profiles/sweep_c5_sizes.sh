#!/bin/bash
# profiles/sweep_c5_sizes.sh TAG -- BASELINE config 5: the colour-transform chain at 1-64 MP per image for every texel
# format, ~256 MP per launch (a batch of frames in one allocation, one kernel launch), one bench.py invocation per size,
# sustained (0.5 s) and burst figures; JSON lines into gpurun_out/c5_sizes_TAG.jsonl, table on stdout.
TAG=${1:-r02}
mkdir -p gpurun_out
out=gpurun_out/c5_sizes_$TAG.jsonl
: > $out
for size in 1024x1024 1920x1080 2048x2048 3840x2160 4096x4096 7680x4320 8192x8192; do
  w=${size%x*}; h=${size#*x}
  frames=$(( (268435456 + w * h - 1) / (w * h) ))
  timeout 300 python bench.py --workload c5_rgba8,c5_rgba16f,c5_rgb10a2,c5_yuv420_rgba8,c5_yuv420_yuv420 --size $size --frames $frames --no-cpu --no-e2e --min-seconds 0.5 >> $out 2>> gpurun_out/c5_sizes_$TAG.err
done
python - "$out" <<'PY'
import json, sys
order = ["c5_rgba8", "c5_rgba16f", "c5_rgb10a2", "c5_yuv420_rgba8", "c5_yuv420_yuv420"]
print("| image | frames / launch | " + " | ".join(order) + " |")
print("|---|---|" + "---|" * len(order))
for l in open(sys.argv[1]):
    if not l.strip().startswith("{"): continue
    d = json.loads(l)
    rows = {d["config"]["workload"]: d}
    rows.update(d.get("workloads", {}))
    c = d["config"]
    cells = ["%.0f (%.2f / %.2f)" % (rows[w]["value"], rows[w]["roofline"]["frac"], rows[w]["roofline"]["burst"]["frac"]) if w in rows else "-" for w in order]
    print("| %dx%d (%.1f MP) | %d | %s |" % (c["width"], c["height"], c["width"] * c["height"] / 1e6, d["frames_per_launch_per_gpu"], " | ".join(cells)))
PY
