#!/bin/bash
# usage (on a box with N GPUs): bash profiles/multi_gpu_checks.sh N TAG   -> gpurun_out/TAG_*.txt
N=$1; TAG=$2; OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
mkdir -p $OUT
for n in 1 2 4 8; do
  [ $n -le $N ] || continue
  $TR --nproc-per-node $n --master-port 2950$n profiles/pcie_ceiling.py 2>/dev/null | grep '^{' >> $OUT/${TAG}_pcie.jsonl
done
if [ -n "$VARIANTS" ]; then
PCIE_NUMA=0 $TR --nproc-per-node $N --master-port 29511 profiles/pcie_ceiling.py 2>/dev/null | grep '^{' >> $OUT/${TAG}_pcie.jsonl
PCIE_WC=1 $TR --nproc-per-node $N --master-port 29512 profiles/pcie_ceiling.py 2>/dev/null | grep '^{' >> $OUT/${TAG}_pcie.jsonl
fi
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
lscpu | head -30 >> $OUT/${TAG}_topo.txt; numactl -H >> $OUT/${TAG}_topo.txt 2>&1
$TR --nproc-per-node $N --master-port 29513 tests/multi_gpu/gather_c_abi.py > $OUT/${TAG}_gather_nccl.txt 2>&1
python tests/multi_gpu/gather_c_abi.py --peer $N > $OUT/${TAG}_gather_peer.txt 2>&1
tail -n 6 $OUT/${TAG}_gather_nccl.txt; tail -n 2 $OUT/${TAG}_gather_peer.txt; cat $OUT/${TAG}_pcie.jsonl
