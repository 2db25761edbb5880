for v in "$@"; do
  echo "== variant $v"
  for thr in 3 4 6; do MC3D_LIB=$PWD/variants/libmc3d_$v.so python tools/profile_walk.py 1e7 3 spectral 4 256 $thr | sort -k4 -n | head -1; done
  MC3D_LIB=$PWD/variants/libmc3d_$v.so python tools/profile_walk.py 1e8 2 spectral 4 256 4 | sort -k4 -n | head -1
done
