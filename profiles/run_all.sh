#!/usr/bin/env bash
# Every CUDA-event timing driver of this directory in one go (one gpurun call, ~3 GPU-minutes on a B200):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash profiles/run_all.sh r02'
# Outputs land in gpurun_out/<tag>_*.txt; copy the ones worth keeping into profiles/.
set -u
tag=${1:-run}
out=gpurun_out
mkdir -p "$out"
run() { name=$1; shift; echo "== $name: $*"; "$@" 2>&1 | grep -v '^\[build\]' | tee "$out/${tag}_${name}.txt"; }
run axhelm_gs   python profiles/tune_axhelm.py --check
run orth        python profiles/tune_orth.py --ks 25,50,75,100
run conv        python profiles/run_conv.py
run hmholtz     python profiles/run_hmholtz.py
run stepper     python profiles/run_stepper.py
run kschur      python profiles/run_krylov_schur.py
python bench.py --steps 3 --warmup 3 > "$out/${tag}_bench_1gpu.json" 2> "$out/${tag}_bench_1gpu.err"
python - "$out/${tag}_bench_1gpu.json" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
k = d['roofline']['kernels']
print('== bench: %.1f steps/s, %.2f ms/step, matvec %.1f GDOF/s, e2e %.1f steps/s, SM %d MHz %s' % (
    d['value'], d['arnoldi_ms_per_step'], d['matvec_gdof_per_s'], d['e2e']['value'], d['clocks']['sm_mhz'],
    d['clocks']['reasons']))
print('   ' + '  '.join('%s %.3f ms %.0f GB/s' % (n, v['ms'] / v['launches'], v['achieved_gbs']) for n, v in k.items()))
PY
