out=gpurun_out/r2_refresh
mkdir -p $out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"hdnnp_eval2f_kernel|neighbor_rows_kernel" -s 6 -c 4 -o $out/prof_full_b -f \
    python tools/v2_check.py 99999 0 > $out/ncu_full_b.log 2>&1
ls -la $out/prof_full_b.ncu-rep
