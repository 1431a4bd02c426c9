#!/bin/bash
# Round-1 measurement pass on one B200 (run under gpurun from the repo root): GPU test-suite, default bench line, ncu launch
# list of the bench command, one ncu --set full capture of the tile kernels + the preconditioner solve.
mkdir -p gpurun_out
(time timeout 400 python -m pytest tests -m gpu -x -q) > gpurun_out/tests_gpu.log 2>&1; tail -4 gpurun_out/tests_gpu.log
timeout 300 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err
VDO_BENCH_FRAMES=6 timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
timeout 240 ncu --set full --import-source on --clock-control none -k regex:"k_tile_lin|k_tile_schur|k_pcg_step_a|k_tile_precond" -c 14 -f -o gpurun_out/r1_final python profiles/ncu_lm.py --iters 2 > gpurun_out/ncu_final.log 2>&1; tail -2 gpurun_out/ncu_final.log
