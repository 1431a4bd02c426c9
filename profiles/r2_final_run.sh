#!/bin/bash
# Round-2 measurement pass on one B200 (run under gpurun from the repo root): GPU test-suite, smoke(), default bench line, ncu launch
# list of the bench command, one ncu --set full capture of the per-frame kernels (image side, flow LM, RANSAC) and of the PCG's kernels.
mkdir -p gpurun_out
(time timeout 500 python -m pytest tests -m gpu -x -q) > gpurun_out/tests_gpu.log 2>&1; tail -4 gpurun_out/tests_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err
VDO_BENCH_FRAMES=6 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_bench_r2.csv python bench.py --steps 1 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
timeout 240 ncu --set full --import-source on --clock-control none -k regex:"k_resize_u8|k_fast_score|k_fast_cells|k_flow2_lm|k_pnp|k_blur7|k_orb_descriptors|k_ic_angle|k_sample|k_filter" -c 24 -f -o gpurun_out/r2_frame python profiles/time_pipeline.py 5 > gpurun_out/ncu_frame.log 2>&1; tail -2 gpurun_out/ncu_frame.log
timeout 240 ncu --set full --import-source on --clock-control none -k regex:"k_tile_schur2|k_band_mul|k_pcg_step_a|k_pcg_p_hpp|k_tile_finalize_schur2" -c 12 --launch-skip 40 -f -o gpurun_out/r2_pcg python profiles/ncu_lm.py --iters 1 > gpurun_out/ncu_pcg.log 2>&1; tail -2 gpurun_out/ncu_pcg.log
