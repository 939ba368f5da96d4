#!/bin/bash
# profiles/sanitize.sh TAG -- compute-sanitizer over the GPU tests: memcheck on the whole suite, racecheck (shared-memory hazards: the
# mbarrier pipelines of k_affine_f16 / k_frame_spec / k_gather_tma, the table fills) and synccheck on the kernels that stage through shared memory.
TAG=${1:-r02}
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 1200 $CS --tool memcheck --error-exitcode 1 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/${TAG}_compute_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
timeout 1200 $CS --tool racecheck --racecheck-report all --error-exitcode 1 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "frame or c4 or affine or fused or blend or porter or lut or rgb10 or lab or yuv or multi" > gpurun_out/${TAG}_compute_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
timeout 900 $CS --tool synccheck --error-exitcode 1 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "frame or c4 or affine or fused" > gpurun_out/${TAG}_compute_sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?"
tail -n 4 gpurun_out/${TAG}_compute_sanitizer_memcheck.log gpurun_out/${TAG}_compute_sanitizer_racecheck.log gpurun_out/${TAG}_compute_sanitizer_synccheck.log
