for v in "$@"; do
  echo "== variant $v"
  MC3D_LIB=$PWD/variants/libmc3d_$v.so python tools/profile_walk.py 1e7 3 spectral 4 256 4 | sort -k4 -n | head -1
  MC3D_LIB=$PWD/variants/libmc3d_$v.so python tools/profile_walk.py 1e7 3 spectral 5 256 4 | sort -k4 -n | head -1
done
