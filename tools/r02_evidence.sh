# round-2 evidence run (one GPU): tests, bench (both arms), launch list, ncu captures of the kernels, timeline
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r02_bench_reference_n1.json 2> gpurun_out/r02_bench_reference_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 1 --warmup 3 --calls-per-step 16 --no-cpu > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:walk_kernel -c 1 -o gpurun_out/r02_walk_steady_1e7 -f python tools/profile_walk.py 1e7 1 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:walk_kernel -c 1 -o gpurun_out/r02_walk_bench_1e6 -f python tools/profile_walk.py 1e6 1 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:init_kernel -c 1 -o gpurun_out/r02_init_1e7 -f python tools/profile_walk.py 1e7 1 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:finalize_kernel -c 1 -o gpurun_out/r02_finalize_1e7 -f python tools/profile_walk.py 1e7 1 > /dev/null 2>&1
MC3D_WALK_PATH=fused ncu --set full --import-source on --clock-control none -k regex:fused_kernel -c 1 -o gpurun_out/r02_fused_2p1um_1e7 -f python tools/short_walk_probe.py 1e7 4 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4 --csv --log-file gpurun_out/r02_persistent_short_walk_launches.csv python tools/short_walk_probe.py 1e7 4 > /dev/null 2>&1
MC3D_TAIL=0 MC3D_LIB=monte_carlompi_b200/libmc3d_timeline.so python tools/timeline.py 1e6 > gpurun_out/r02_lone_launch_timeline.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:tail_kernel -c 1 -o gpurun_out/r02_tail_1e6 -f python tools/profile_walk.py 1e6 1 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/r02_lone_call_launches.csv python tools/profile_walk.py 1e6 3 > /dev/null 2>&1
(for t in 0 1; do echo "== MC3D_TAIL=$t"; MC3D_TAIL=$t python tools/tail_latency.py; done) > gpurun_out/r02_tail_latency.log 2>&1
(for t in 0 1; do echo "== MC3D_TAIL=$t (one isolated 1e6-photon call, C2; then 1e7 visible photons, C4)"; MC3D_TAIL=$t python tools/profile_walk.py 1e6 6 | tail -5; MC3D_TAIL=$t python tools/profile_walk.py 1e7 3 const-vis | tail -2; done) > gpurun_out/r02_tail_kernel_on_off.log 2>&1
python tools/fused_vs_persistent.py > gpurun_out/r02_fused_vs_persistent.log 2>&1
python tools/run_configs.py c1 c2 c4 > gpurun_out/r02_configs_c1_c2_c4.log 2>&1
python tools/run_c3_grid.py 1e7 1 3 > gpurun_out/r02_c3_grid.log 2>&1
python tools/driver_demo.py > gpurun_out/r02_driver_demo.log 2>&1
ls -la gpurun_out | tail -30
