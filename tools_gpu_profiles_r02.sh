#!/bin/bash
# Round-2 measurement run (under gpurun, 1 GPU): GPU tests, smoke, every bench line, ncu launch lists (shares of the step)
# and one --set full capture of the dominant kernels.  Everything lands in gpurun_out/; tools/make_profiles.py turns it
# into the tracked summaries under profiles/r02_*.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -n 2 gpurun_out/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -n 1 gpurun_out/r02_smoke.log
python bench.py --steps 200 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench.err; echo "bench exit $?"
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r02_bench_reference.json 2>> gpurun_out/r02_bench.err
python bench.py --steps 50 --precision fp32 --no-cpu-baseline --no-stress > gpurun_out/r02_bench_n1_fp32.json 2>> gpurun_out/r02_bench.err
for c in obb assign mask train; do timeout 400 python bench.py --config $c --steps 50 > gpurun_out/r02_bench_$c.json 2> gpurun_out/r02_bench_$c.err; echo "$c exit $?"; done
for t in mb_roi_sweep.py "mb_roi_sweep.py rot" "mb_roi_bwd.py 5000" "mb_roi_bwd.py 96000" mb_gemm_waves.py mb_heads.py trace_train.py "mb_train_dist.py 0"; do echo "== $t"; python tools/$t; done > gpurun_out/r02_microbench.log 2>&1
# launch lists (each only after the same command exited 0 without ncu)
B="bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-stress --no-train"
python $B > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv python $B > gpurun_out/ncu_launches.log 2>&1
python tools/prof_obb.py 3 > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/obb_launches.csv python tools/prof_obb.py 3 > gpurun_out/ncu_obb.log 2>&1
python tools/prof_train.py 3 > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/train_launches.csv python tools/prof_train.py 3 > gpurun_out/ncu_train.log 2>&1
# full captures
ncu --set full --clock-control none --import-source on -k regex:"fc_gemm_kernel|roi_align_mma_kernel" -s 8 -c 6 \
    -o gpurun_out/prof_step -f python $B > gpurun_out/ncu_full.log 2>&1
python tools/prof_roi.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:roi_align_mma -s 1 -c 1 -o gpurun_out/prof_roi_stress -f python tools/prof_roi.py > gpurun_out/ncu_roi_stress.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"roi_align_mma_kernel" -s 2 -c 2 -o gpurun_out/prof_obb_roi -f python tools/prof_obb.py 2 > gpurun_out/ncu_obb_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"roi_align_bwd_mma_kernel" -s 2 -c 1 -o gpurun_out/prof_roi_bwd -f python tools/prof_train.py 2 > gpurun_out/ncu_bwd_full.log 2>&1
ls gpurun_out | wc -l
