# round-2 measurement set on one B200 (run through gpurun from the repo root)
set -x
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
( for n in 100000000 12500000; do echo "== N=$n"; PCF_LIB=parcompfin_b200/libpcf_tuning.so timeout 300 python tools/tune_amer_persistent.py $n; done ) > gpurun_out/r2_tune_amer_persistent.log 2>&1
tail -12 gpurun_out/r2_tune_amer_persistent.log
timeout 900 python bench.py > gpurun_out/r2e_bench_1gpu.json 2> gpurun_out/r2e_bench_1gpu.err; tail -c 300 gpurun_out/r2e_bench_1gpu.json
timeout 600 python bench.py --impl reference > gpurun_out/r2e_bench_reference_arm.json 2>/dev/null; head -c 600 gpurun_out/r2e_bench_reference_arm.json
python __graft_entry__.py smoke 2>&1 | tail -2
