# one 10^6-photon call at a time (what a plain MonteCarlo.run issues): drain-phase variants x grid size
for bps in 1 2 4; do for lat in 0 1 2; do for g in 0 8 16 24; do
  echo "blocks/SM $bps latency $lat give $g: $(MC3D_DRAIN_LATENCY=$lat MC3D_DRAIN_GIVE=$g python tools/profile_walk.py 1e6 7 spectral $bps 256 4 | awk '{print $4}' | tail -5 | tr '\n' ' ')"
done; done; done
