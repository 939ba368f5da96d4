#!/bin/bash
# profiles/sass_evidence.sh > profiles/r02_sass_evidence.txt -- what the built objects contain (cuobjdump -sass of zosimos_b200/csrc/_build/*.o):
# TMA bulk tensor loads (UTMALDG), mbarrier traffic (SYNCS), packed f32x2 arithmetic (FFMA2 / FMUL2 / FADD2), f16 pair arithmetic (HADD2 / HFMA2),
# SFU calls (MUFU), and that there is no tensor-core instruction (nothing on this path is a contraction).
cd "$(dirname "$0")/../zosimos_b200/csrc/_build" || exit 1
echo "nvcc: $(nvcc --version | tail -2 | head -1);  flags: $(grep '^NVFLAGS' ../Makefile)"
printf "%-22s %8s %8s %8s %8s %8s %8s %8s %8s %8s %10s\n" object UTMALDG SYNCS FFMA2 FMUL2 FADD2 HADD2 MUFU LDS LDG "UTCMMA/HMMA"
for o in *.o; do
  [ "$o" = host.o ] && continue
  cuobjdump -sass "$o" > /tmp/sass_all.txt 2>/dev/null
  c() { grep -c -E "^\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P[0-9T]+\s+)?$1" /tmp/sass_all.txt; }
  printf "%-22s %8d %8d %8d %8d %8d %8d %8d %8d %8d %10d\n" "$o" $(c UTMALDG) $(c SYNCS) $(c FFMA2) $(c FMUL2) $(c FADD2) $(c HADD2) $(c MUFU) $(c "LDS") $(c "LDG") $(c "(UTCMMA|HMMA|IMMA|UTCHMMA)")
done
echo
echo "kernels (entry points) per object:"
for o in *.o; do [ "$o" = host.o ] && continue; printf "%-22s %d\n" "$o" $(cuobjdump -sass "$o" 2>/dev/null | grep -c "Function :"); done
