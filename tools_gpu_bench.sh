#!/bin/bash
# bench + smoke + ncu launch list + one full capture of the top kernels (run under gpurun, 1 GPU)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
python bench.py --steps 50 --precision fp32 --no-cpu-baseline --no-stress > gpurun_out/bench_fp32.json 2>> gpurun_out/bench.err
# launch list (only after the same command exited 0 without ncu)
python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-stress --no-train > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-stress --no-train > gpurun_out/ncu_launches.log 2>&1
# full capture of the two dominant kernels inside the step
ncu --set full --clock-control none --import-source on -k regex:"fc_gemm_kernel|roi_align_mma_kernel" -s 8 -c 6 \
    -o gpurun_out/prof_step -f python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-stress --no-train > gpurun_out/ncu_full.log 2>&1
tail -n 3 gpurun_out/smoke.log; cat gpurun_out/bench.json | cut -c1-4000; cat gpurun_out/bench_ref.json; tail -n 5 gpurun_out/bench.err; tail -n 3 gpurun_out/ncu_full.log
python tools/prof_roi.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:roi_align_mma -s 1 -c 1 -o gpurun_out/prof_roi_stress -f python tools/prof_roi.py > gpurun_out/ncu_roi_stress.log 2>&1
# secondary configurations (one JSON line each) and the OBB launch list / rotated RoIAlign capture
for c in obb assign mask; do timeout 280 python bench.py --config $c --steps 50 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "$c exit $?"; done
python tools/prof_obb.py 3 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/obb_launches.csv python tools/prof_obb.py 3 > gpurun_out/ncu_obb.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"roi_align_mma_kernel" -s 2 -c 2 -o gpurun_out/prof_obb_roi -f python tools/prof_obb.py 2 > gpurun_out/ncu_obb_full.log 2>&1
