python -m pytest tests/test_gpu_production.py -x -q 2>&1 | tail -5
python tools/short_walk_probe.py 2>&1 | tee gpurun_out/r02_short_walk_probe.log
MC3D_WALK_PATH=fused ncu --set full --import-source on --clock-control none -k regex:fused_kernel -c 1 -o gpurun_out/r02_fused_2p1um_1e7 -f python tools/short_walk_probe.py 1e7 4 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
