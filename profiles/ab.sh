#!/bin/bash
# A/B of builds of libzosimos_cuda.so on ONE box (sustained numbers depend on the individual GPU's power behaviour, so
# runs from different gpurun calls do not compare): usage  profiles/ab.sh "w1,w2,.." "base.so other.so .." [tag]
# prints, per build ("new" = the in-tree library), burst (10 ms timed region) and sustained (1 s) fractions of the HBM figure.
WL=$1; LIBS=$2; TAG=${3:-ab}
for round in ${ROUNDS:-1 2}; do
  for which in $LIBS new; do
    if [ $which = new ]; then unset ZOS_CUDA_LIB; else export ZOS_CUDA_LIB=$PWD/$which; fi
    for mode in burst sustained; do
      if [ $mode = burst ]; then MS=0.01; else MS=1.0; fi
      f=gpurun_out/${TAG}_$(basename $which .so)_${mode}_$round.json
      python bench.py --workload $WL --no-cpu --no-e2e --min-seconds $MS > $f 2>${f%.json}.err
      python - <<PY
import json
d=json.loads(open("$f").read().strip().splitlines()[-1])
rows=list({d["config"]["workload"]: d, **d.get("workloads",{})}.items())
print("%-22s %-9s %d"%("$(basename $which .so)","$mode",$round), "  ".join("%s %.4f (%s MHz, %s W)"%(n,r["roofline"]["frac"],r["clocks"]["sm_mhz"],r["clocks"].get("power_w_max")) for n,r in rows if "roofline" in r))
PY
    done
  done
done
