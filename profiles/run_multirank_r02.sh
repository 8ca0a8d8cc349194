#!/bin/bash
# Multi-rank parity logs of round 2: tests/multirank_check.py at W ranks x {NCCL, NVLink peer memory}, kept under
# gpurun_out/$OUT (copied to profiles/ by hand).  usage: run_multirank_r02.sh "<world sizes>" <outdir>
set -u
WORLDS=${1:-"2"}
OUT=${2:-gpurun_out/mr_r02}
mkdir -p $OUT
rc=0
for W in $WORLDS; do
  for P2P in 0 1; do
    tag=$([ $P2P = 1 ] && echo p2p || echo nccl)
    port=$((29700 + W * 4 + P2P))
    NSB_TEST_P2P=$P2P timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$W --master-addr 127.0.0.1 \
      --master-port $port tests/multirank_check.py > $OUT/multirank_check_w${W}_${tag}.log 2>&1
    r=$?
    echo "world=$W transport=$tag rc=$r" | tee -a $OUT/summary.txt
    grep -h "^\[rank" $OUT/multirank_check_w${W}_${tag}.log | tee -a $OUT/summary.txt
    [ $r -ne 0 ] && rc=1 && tail -30 $OUT/multirank_check_w${W}_${tag}.log
  done
done
exit $rc
