#!/bin/bash
# 8-GPU evidence of the final round-2 code in one call: multi-rank parity logs (both transports) and the bench line
# at N = 8 with the folded normalisation on (default) and off.   gpurun --gpus 8 -- 'bash profiles/run_scale8_r02.sh'
set -u
OUT=gpurun_out/scale8_r02b
mkdir -p $OUT
bash profiles/run_multirank_r02.sh "8" $OUT > $OUT/mr.out 2>&1; echo "multirank rc=$?"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29811 bench.py --gpus 8 > $OUT/bench_8gpu.json 2> $OUT/bench_8gpu.err; echo "bench8 rc=$?"
NSB_FOLD_NORM=0 timeout 300 $TR --master-port 29812 bench.py --gpus 8 --no-e2e --no-c0 --no-dgks > $OUT/bench_8gpu_nofold.json 2> $OUT/bench_8gpu_nofold.err; echo "bench8 nofold rc=$?"
python - <<'PY'
import json
for f in ('bench_8gpu', 'bench_8gpu_nofold'):
    try:
        d = json.load(open(f'gpurun_out/scale8_r02b/{f}.json'))
        print(f, round(d['value'], 1), round(d['arnoldi_ms_per_step'], 4), d.get('value_c0'), d.get('e2e', {}).get('value'), d['parity']['ok'],
              {k: round(v['ms'] / v['launches'], 4) for k, v in d['roofline']['kernels'].items()})
    except Exception as e:
        print(f, 'failed', e)
PY
