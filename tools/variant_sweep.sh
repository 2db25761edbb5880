for v in "$@"; do
  echo "== variant $v"
  MC3D_LIB=$PWD/variants/libmc3d_$v.so python tools/profile_walk.py 1e7 3 spectral 4 256 4 | sort -k4 -n | head -1
  MC3D_LIB=$PWD/variants/libmc3d_$v.so python tools/profile_walk.py 1e6 5 spectral 255 256 4 | sort -k4 -n | head -1
  MC3D_LIB=$PWD/variants/libmc3d_$v.so python -m pytest tests/test_gpu_production.py -x -q -k "same_philox or ragged or split" 2>&1 | tail -1
done
