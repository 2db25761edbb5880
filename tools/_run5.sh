for lat in 0 1; do for g in 0 8 16 24; do
  echo "latency $lat give $g: $(MC3D_DRAIN_LATENCY=$lat MC3D_DRAIN_GIVE=$g python tools/profile_walk.py 1e6 6 spectral | awk '{print $4}' | tail -5 | tr '\n' ' ')"
done; done
for lat in 0 1; do for g in 0 16; do echo "latency $lat give $g vis 1e7: $(MC3D_DRAIN_LATENCY=$lat MC3D_DRAIN_GIVE=$g python tools/profile_walk.py 1e7 2 const-vis | awk '{print $4}' | tr '\n' ' ')"; done; done
