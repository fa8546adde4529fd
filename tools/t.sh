for v in e768 e640 e1024; do
  echo "$v: $(SIMIAN_B200_LIB=$PWD/simian_b200/libsimian_b200_$v.so timeout 300 python tools/quick_run.py config2_phold_1m 1 2>&1 | tail -1)"
done
