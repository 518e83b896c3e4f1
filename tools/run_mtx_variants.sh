# usage: bash tools/run_mtx_variants.sh v1 v2 ...   (libraries under build/variants/, made by tools/build_variants.py)
for v in "$@"; do
  echo "== $v"
  P3D_CORE_LIB=$PWD/build/variants/$v.so timeout 60 python tools/prof_mt.py 128 30 capi
  P3D_CORE_LIB=$PWD/build/variants/$v.so timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_mtx --launch-skip 5 -c 5 --csv python tools/prof_mt.py 128 0 capi 2>/dev/null | grep k_mtx | awk -F'","' '{split($5,a,"("); print "   ", a[1], $NF}' | tr -d '"'
done
