#!/bin/bash
# profiles/profile_all.sh TAG -- one `ncu --set full` capture of the dominant kernel of every bench workload, and the launch list
# of the default bench command.  Reports land in gpurun_out/TAG_<workload>.ncu-rep; summarise here with profiles/facts.py.
TAG=${1:-r02}
mkdir -p gpurun_out
for w in c2_blend c2_inscribe c1_oklab c3_affine_nearest c3_affine_bilinear c4_fused c5_rgba8 c5_rgba16f c5_rgb10a2 c5_yuv420_yuv420 c5_yuv420_rgba8; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ -s 8 -c 1 -f -o gpurun_out/${TAG}_$w \
    python bench.py --workload $w --no-cpu --no-e2e --min-seconds 0.001 --steps 3 --warmup 3 > gpurun_out/${TAG}_$w.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches_default_bench.csv \
  python bench.py --steps 2 --warmup 3 --min-seconds 0.01 --no-cpu > gpurun_out/${TAG}_launches_default_bench.log 2>&1
ls -la gpurun_out/${TAG}_*.ncu-rep | wc -l
