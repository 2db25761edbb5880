set -x
for l in "" "2 256 4" "4 256 4" "3 256 4"; do
  echo "== launch $l"; python tools/profile_walk.py 1e6 4 spectral $l
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_iso_launches_auto.csv python tools/profile_walk.py 1e6 3 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_iso_launches_2bps.csv python tools/profile_walk.py 1e6 3 spectral 2 256 4 > /dev/null 2>&1
