for v in p128 p256 p128t128; do
  for o in "" "bucket_width=0.025"; do
    echo "$v $o: $(SIMIAN_B200_LIB=$PWD/simian_b200/libsimian_b200_$v.so timeout 300 python tools/quick_run.py config2_phold_1m 1 $o 2>&1 | tail -1)"
  done
done
echo "default bw.025: $(timeout 300 python tools/quick_run.py config2_phold_1m 1 bucket_width=0.025 2>&1 | tail -1)"
