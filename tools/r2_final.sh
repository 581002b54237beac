# final measurement set of round 2 on one B200: tests, sanitizer, both bench arms, smoke, the two binomial captures,
# the launch list of the bench command
set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_target.py > gpurun_out/r2_compute_sanitizer_memcheck.log 2>&1; tail -1 gpurun_out/r2_compute_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_target.py > gpurun_out/r2_compute_sanitizer_racecheck.log 2>&1; tail -1 gpurun_out/r2_compute_sanitizer_racecheck.log
timeout 900 python bench.py > gpurun_out/r2g_bench_1gpu.json 2> gpurun_out/r2g_bench_1gpu.err; tail -c 200 gpurun_out/r2g_bench_1gpu.json
timeout 600 python bench.py --impl reference > gpurun_out/r2g_bench_reference_arm.json 2>/dev/null; head -c 300 gpurun_out/r2g_bench_reference_arm.json
python __graft_entry__.py smoke 2>&1 | tail -1
PM=sm__inst_executed_pipe_fp64,sm__inst_executed_pipe_fmaheavy,sm__inst_executed_pipe_fmalite,sm__inst_executed_pipe_fma,sm__inst_executed_pipe_alu,sm__inst_executed_pipe_xu,sm__inst_executed_pipe_lsu,sm__pipe_fp64_cycles_active,sm__pipe_shared_cycles_active,sm__pipe_fmaheavy_cycles_active,sm__pipe_fmalite_cycles_active,sm__pipe_alu_cycles_active,sm__inst_executed,sm__cycles_active,smsp__issue_active
mkdir -p /tmp/ncu
cap() {
  timeout 600 ncu --set full --metrics $PM --clock-control none --import-source on -k regex:"$2" -s "$3" -c 1 -f \
    -o /tmp/ncu/$1 python tools/ncu_target.py "$4" "$5" 2 2>&1 | tail -1
  { python tools/ncu_summary.py /tmp/ncu/$1.ncu-rep "ncu --set full --clock-control none, shipped libpcf.so, one launch: python tools/ncu_target.py $4 $5 (kernel regex $2, launch skip $3)"
    python tools/ncu_pipes.py /tmp/ncu/$1.ncu-rep
    python tools/ncu_source_top.py /tmp/ncu/$1.ncu-rep 20; } > gpurun_out/r2g_ncu_$1.txt 2>&1
  rm -f /tmp/ncu/$1.ncu-rep
}
cap binom_screen       binom_terms_kernel          1 binom_embar 2147483647
cap binom_noscreen     binom_terms_kernel          1 binom_embar_noscreen 100000000
cap amer_sweep         amer_sweep_kernel           26 mc_amer 100000000
bash tools/r2_launch_list.sh
