# steady-state walk kernel rate (1e7 photons per launch) for several launch shapes
for cfg in "4 256 4" "5 256 4" "6 256 4" "5 256 3" "5 256 6" "10 128 4" "12 128 4" "9 128 4" "3 512 4"; do set -- $cfg
echo "bps $1 bt $2 thr $3: $(python tools/profile_walk.py 1e7 3 spectral $1 $2 $3 | sort -t' ' -k4 -n | head -1)"; done
