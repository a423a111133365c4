out=gpurun_out/r2_v2c
mkdir -p $out
for v in default ch4 nu2b3 nu1; do
  if [ $v = default ]; then unset PANTEA_B200_LIB; else export PANTEA_B200_LIB=$PWD/pantea_b200/variants/lib_$v.so; fi
  timeout 300 python tools/v2_check.py 99999 2048 > $out/v2_check_$v.txt 2>&1; echo "== $v"; grep -E "fast vs|force evaluation|vs oracle|\[fast\]" $out/v2_check_$v.txt | head -6
done
unset PANTEA_B200_LIB
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"hdnnp_eval2_kernel|pair_filter2_kernel" \
    -s 2 -c 2 -o $out/prof_v2 -f python tools/v2_check.py 99999 0 > $out/ncu_full.log 2>&1
