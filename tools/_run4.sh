python -m pytest tests/test_gpu_production.py -x -q 2>&1 | tail -3
python tools/short_walk_probe.py 2>&1 | tee gpurun_out/r02_short_walk_probe_v2.log
for g in 0 8 16 24 31; do for l in "1 256 4" "2 256 4" "4 256 4"; do
  echo "give $g launch $l: $(MC3D_DRAIN_GIVE=$g python tools/profile_walk.py 1e6 5 spectral $l | awk '{print $4}' | tail -4 | tr '\n' ' ')"
done; done
for g in 0 16; do echo "give $g 1e7: $(MC3D_DRAIN_GIVE=$g python tools/profile_walk.py 1e7 3 | awk '{print $4}' | tr '\n' ' ')"; done
for g in 0 16; do echo "give $g vis 1e7: $(MC3D_DRAIN_GIVE=$g python tools/profile_walk.py 1e7 2 const-vis | awk '{print $4}' | tr '\n' ' ')"; done
