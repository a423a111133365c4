for v in default apw8 apw16 apw2f2; do
  if [ $v = default ]; then unset PANTEA_B200_LIB; else export PANTEA_B200_LIB=$PWD/pantea_b200/variants/lib_$v.so; fi
  echo "== $v"; timeout 300 python tools/brick_profile.py 99999 20 2>&1 | grep -v "Warn\|warn" | head -3
done
