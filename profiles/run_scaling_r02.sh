#!/bin/bash
# Round-2 multi-GPU evidence on one 8-GPU box: parity logs at 4 and 8 ranks (both transports), then the bench
# at 8 and 4 GPUs in three launch structures: default (kernel tails + one CUDA graph per Arnoldi step), NSB_GRAPH=0
# (tails, plain launches) and NSB_GRAPH=0 NSB_TAIL=0 (round 1: separate reduce / all-reduce / add launches).
OUT=${1:-gpurun_out/scale_r02}
mkdir -p $OUT
bash profiles/run_multirank_r02.sh "4 8" $OUT
run_bench() {  # gpus tag env...
  local n=$1 tag=$2; shift 2
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
    --master-port $((29800 + n)) bench.py --gpus $n --steps 3 --warmup 3 --no-e2e --no-dgks --no-cpu --no-c0 --no-single-rank-check \
    > $OUT/bench_${n}gpu_$tag.json 2> $OUT/bench_${n}gpu_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_${n}gpu_$tag.json").read().strip().splitlines()[-1])
    print("$n GPUs $tag: %.1f steps/s  %.4f ms/step  launches %d  parity %s" % (d["value"], d["arnoldi_ms_per_step"], d["gpu_launches"], d["parity"]["ok"]))
except Exception as e:
    print("$n GPUs $tag: FAILED", e)
PY
}
run_bench 8 graph NSB_DUMMY=1
run_bench 8 plain NSB_GRAPH=0
run_bench 8 legacy NSB_GRAPH=0 NSB_TAIL=0
run_bench 8 graph_b NSB_DUMMY=1
run_bench 4 graph NSB_DUMMY=1
run_bench 4 legacy NSB_GRAPH=0 NSB_TAIL=0
# the full line (parity block incl. the single-rank comparison, e2e, DGKS) at 8 GPUs
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29877 \
  bench.py --gpus 8 --steps 3 --warmup 3 > $OUT/bench_8gpu_full.json 2> $OUT/bench_8gpu_full.err
python -c "
import json; d=json.loads(open('$OUT/bench_8gpu_full.json').read().strip().splitlines()[-1]); print('8 GPUs full:', d['value'], d['e2e']['value'], d['parity'])"
