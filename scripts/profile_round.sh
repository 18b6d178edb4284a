#!/bin/bash
# ncu evidence for profiles/ (run on the GPU box through gpurun; outputs into gpurun_out/).
#   launch list of bench.py, full capture of one TrackFrame batch (296 streams), full capture of the BA
#   kernels of one LM step at C4 (per-measurement passes, Schur, early and late LDL^T steps, back substitution)
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 700 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-ba --no-cpu-baseline > gpurun_out/launch_bench.log 2>&1
$NCU --set full --import-source on --profile-from-start off -f -o gpurun_out/tracker \
    python bench.py --steps 2 --warmup 3 --no-ba --no-cpu-baseline --ncu-step > gpurun_out/ncu_tracker.log 2>&1
$NCU --set full --import-source on -f -o gpurun_out/ba_passes \
    -k regex:'k_ba_project|k_ba_jacobian|k_ba_acc|k_ba_vinv|k_ba_schur|k_ba_point_update|k_ba_new_error|k_ldlt_back|k_ldlt_scale' -c 11 \
    python scripts/ba_c4_once.py C4 1 > gpurun_out/ncu_ba_passes.log 2>&1
$NCU --set full --import-source on -f -o gpurun_out/ba_solve -k regex:'k_ldlt_panel|k_ldlt_step|k_ldlt_update' -c 4 \
    python scripts/ba_c4_once.py C4 1 > gpurun_out/ncu_ba_solve.log 2>&1
$NCU --set full --import-source on -f -o gpurun_out/ba_solve_late -k regex:'k_ldlt_step' -s 38 -c 2 \
    python scripts/ba_c4_once.py C4 1 > gpurun_out/ncu_ba_solve_late.log 2>&1
ls -la gpurun_out/*.ncu-rep; wc -l gpurun_out/launches.csv; tail -2 gpurun_out/ncu_tracker.log
