#!/bin/bash
# profiles/sanitize.sh TAG -- compute-sanitizer over the GPU tests: memcheck on the whole suite; racecheck (shared-memory hazards: table
# fills, the footprint hand-over of k_frame_spec / k_frame_fast, the staged tiles of the gather kernels) and synccheck on the kernels
# that stage through shared memory.  racecheck runs on the small-image tests only: it is ~100x slower, and on the full-size affine test
# it (a) reports the geometry ring of k_affine_f16 as "potential WAR" -- the ring is ordered by the empty / full mbarriers (consumer
# read -> arrive(empty) -> producer try_wait(empty) -> producer write, RING = 8 > AHEAD + STAGES), which racecheck does not model --
# and (b) outlasts the kernels' run-away guards (profiles/r02_compute_sanitizer_racecheck_fullsize_note.txt).
TAG=${1:-r02}
CS=/usr/local/cuda/bin/compute-sanitizer
if [ "$2" != "race-only" ]; then
timeout 1200 $CS --tool memcheck --error-exitcode 1 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/${TAG}_compute_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
timeout 900 $CS --tool synccheck --error-exitcode 1 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "frame or c4 or affine or fused" > gpurun_out/${TAG}_compute_sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?"
fi
timeout 900 $CS --tool racecheck --racecheck-report all --error-exitcode 1 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_multi.py tests/test_gpu_program.py -m gpu -q -x -p no:cacheprovider \
  -k "frame or fused or yuv_fast or lab_kernel or fast_u8 or rgb10 or porter or multi or inscribe or affine or async" > gpurun_out/${TAG}_compute_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -n 4 gpurun_out/${TAG}_compute_sanitizer_racecheck.log
