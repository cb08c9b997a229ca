#!/bin/bash
# Round-2 evidence run (1 GPU): full GPU suite, smoke, bench lines, CPU arm, call-site probe, phase shares, tail timeline,
# ncu launch list + full capture.  Everything lands in gpurun_out/ (copied to profiles/ by hand, see profiles/README.md).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -v Warning gpurun_out/pytest_gpu.log | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_c2_n1.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2_bench_c2_n1.json
TL_NO_FUSED_GRAD=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_c2_n1_nofusedgrad.json 2> gpurun_out/bench_nf.err; echo "bench (gradient in its own launch) rc=$?"; cut -c1-200 gpurun_out/r2_bench_c2_n1_nofusedgrad.json
timeout 600 python bench.py --config c5 --steps 10 --warmup 3 > gpurun_out/r2_bench_c5_n1.json 2> gpurun_out/bench_c5.err; echo "bench c5 rc=$?"; cut -c1-300 gpurun_out/r2_bench_c5_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_n1.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/r2_bench_ref_n1.json
timeout 300 python scripts/probe_callsite.py > gpurun_out/r2_callsite.json 2> gpurun_out/callsite.err; cut -c1-200 gpurun_out/r2_callsite.json
timeout 200 python scripts/r2_probe.py > gpurun_out/r2_phase_shares.log 2>&1; cut -c1-300 gpurun_out/r2_phase_shares.log
timeout 200 python scripts/tail_probe.py > gpurun_out/r2_tail_timeline.log 2>&1; cat gpurun_out/r2_tail_timeline.log
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'ph_|seg_sort|match_|loss_kernel|grad_kernel|scale_kernel|unpack' -c 30 --csv --log-file gpurun_out/r2_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r2_launches_c2.csv
timeout 120 python scripts/ncu_ph.py pred > gpurun_out/ncu_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ph_small -c 1 -o gpurun_out/ph_small_r2 -f python scripts/ncu_ph.py pred > gpurun_out/ncu_run.log 2>&1
echo "full rc=$?"; ls -la gpurun_out/ph_small_r2.ncu-rep
