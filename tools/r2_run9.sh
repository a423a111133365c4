for v in default ru4 wrapdiag; do
  if [ $v = default ]; then unset PANTEA_B200_LIB; else export PANTEA_B200_LIB=$PWD/pantea_b200/variants/lib_$v.so; fi
  echo "== $v"; timeout 300 python tools/brick_profile.py 99999 20 2>&1 | grep -v "Warn\|warn" | head -4
done
unset PANTEA_B200_LIB
timeout 300 python tests/mgpu_check.py 99999 10 brick oracle 2>&1 | tail -1
