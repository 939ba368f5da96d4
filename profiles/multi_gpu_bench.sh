#!/bin/bash
# usage (box with N GPUs): bash profiles/multi_gpu_bench.sh N TAG -> gpurun_out/TAG_*.json(l)
# (i) headline with e2e + gather at every n <= N; (ii) BASELINE config 3: c4_fused, 256 frames, STRONG scaling; (iii) one 8192^2 image in row bands
N=$1; TAG=$2; OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { n=$1; shift; if [ $n = 1 ]; then python bench.py --gpus 1 "$@"; else $TR --nproc-per-node $n --master-port $((29600+n)) bench.py --gpus $n "$@"; fi 2>>$OUT/${TAG}_err.txt | grep '^{'; }
for n in ${NS:-1 2 4 8}; do
  [ $n -le $N ] || continue
  run $n --workload c2_blend --no-cpu >> $OUT/${TAG}_c2_blend.jsonl
  if [ $n = 1 ] || [ $n = $N ] || [ -n "$ALL_N" ]; then
    run $n --workload c4_fused --total-frames 256 --no-cpu >> $OUT/${TAG}_c4_strong.jsonl
    run $n --workload band_affine --no-cpu >> $OUT/${TAG}_bands.jsonl
  fi
done
python - <<PY
import json
for f,keys in (("c2_blend",("value","e2e","gather")),("c4_strong",("value",)),("bands",("value","gather"))):
    for l in open("$OUT/${TAG}_%s.jsonl"%f):
        d=json.loads(l); e=d.get("e2e") or {}; g=d.get("gather") or {}
        print(f, d["n_gpus"], d["value"], d["roofline"]["frac"], "e2e", e.get("value"), (e.get("pcie_gbs") or {}), "gather", {k:(v or {}).get("value") if isinstance(v,dict) and "value" in (v or {}) else v for k,v in g.items() if k in ("ms","gbs_received_per_gpu","off","all_ranks","root_only")})
PY
