#!/bin/bash
# Round-2 evidence in one gpurun call (1 GPU): the full GPU test suite, both bench arms, the ncu launch list of the bench
# command and one `ncu --set full` capture of the hot kernels at k = 100 (source of profiles/ncu_traffic.json).
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash profiles/run_r02_evidence.sh r2ev'
set -u
OUT=gpurun_out/${1:-r2ev}
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee $OUT/summary.txt
tail -3 $OUT/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/summary.txt
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/summary.txt
timeout 600 python bench.py --impl reference > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?" | tee -a $OUT/summary.txt
# launch list: per-launch gpu time of the same command (serialised, cold cache: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-dgks --no-c0 --no-single-rank-check > $OUT/ncu_bench.log 2>&1
echo "ncu list rc=$?" | tee -a $OUT/summary.txt
gzip -f $OUT/launches.csv
# full capture of the last repetition of the hot kernels at k = 100
python profiles/run_hot_kernels.py --k 100 > $OUT/plain_hot.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o $OUT/prof_hot_k100 -f \
  python profiles/run_hot_kernels.py --k 100 > $OUT/ncu_hot.log 2>&1
echo "ncu full rc=$?" | tee -a $OUT/summary.txt
ncu -i $OUT/prof_hot_k100.ncu-rep --page raw --csv \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active \
  > $OUT/hot_k100_raw.csv 2>/dev/null
timeout 300 python profiles/run_ns_stepper.py --nelx 32 --nsteps 3 2>&1 | grep -v "^\[build" > $OUT/run_ns_stepper_32.txt
timeout 300 python profiles/run_ns_stepper.py --nelx 16 --nsteps 5 2>&1 | grep -v "^\[build" > $OUT/run_ns_stepper_16.txt
cat $OUT/summary.txt
