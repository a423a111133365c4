for v in base probe16; do
  if [ $v = base ]; then unset PANTEA_B200_LIB; else export PANTEA_B200_LIB=$PWD/pantea_b200/variants/lib_$v.so; fi
  echo "== $v"
  timeout 300 python tools/v2_check.py 99999 0 2>&1 | grep "fast\]" | grep "eval2\|filter2"
done
