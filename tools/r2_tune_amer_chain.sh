set -x
timeout 600 python -m pytest tests -x -q -m gpu -k "amer" 2>&1 | tail -4
( for n in 100000000 12500000; do echo "== N=$n"; PCF_LIB=parcompfin_b200/libpcf_tuning.so timeout 300 python tools/tune_amer_chain.py $n; done ) > gpurun_out/r2_tune_amer_chain.log 2>&1
grep -v "rep [01]" gpurun_out/r2_tune_amer_chain.log
