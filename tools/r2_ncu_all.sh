# One `ncu --set full` capture per dominant kernel AS SHIPPED (libpcf.so defaults), plus the pipe microbenchmarks.
# Run through gpurun from the repo root. Every report is digested ON THE BOX (tools/ncu_summary.py: raw page;
# tools/ncu_source_top.py: source page) into gpurun_out/r2g_ncu_<name>.txt and deleted: gpurun brings back at most 64 MiB.
PM=sm__inst_executed_pipe_fp64,sm__inst_executed_pipe_fmaheavy,sm__inst_executed_pipe_fmalite,sm__inst_executed_pipe_fma,sm__inst_executed_pipe_alu,sm__inst_executed_pipe_xu,sm__inst_executed_pipe_lsu,sm__pipe_fp64_cycles_active,sm__pipe_shared_cycles_active,sm__pipe_fmaheavy_cycles_active,sm__pipe_fmalite_cycles_active,sm__pipe_alu_cycles_active,sm__inst_executed,sm__cycles_active,smsp__issue_active
mkdir -p /tmp/ncu
cap() {  # name kernel-regex skip workload N header
  timeout 600 ncu --set full --metrics $PM --clock-control none --import-source on -k regex:"$2" -s "$3" -c 1 -f \
    -o /tmp/ncu/$1 python tools/ncu_target.py "$4" "$5" 2 2>&1 | tail -1
  { python tools/ncu_summary.py /tmp/ncu/$1.ncu-rep "ncu --set full --clock-control none, shipped libpcf.so, one launch: python tools/ncu_target.py $4 $5 (kernel regex $2, launch skip $3)"
    python tools/ncu_pipes.py /tmp/ncu/$1.ncu-rep
    python tools/ncu_source_top.py /tmp/ncu/$1.ncu-rep 20; } > gpurun_out/r2g_ncu_$1.txt 2>&1
  [ "$1" = amer_sweep ] && cp /tmp/ncu/$1.ncu-rep gpurun_out/r2g_ncu_$1.ncu-rep
  rm -f /tmp/ncu/$1.ncu-rep
  head -4 gpurun_out/r2g_ncu_$1.txt | tail -2
}
cap mc_asia            mc_asia_kernel              1 mc_asia 100000000
cap mc_eur             mc_eur_kernel               1 mc_eur 2000000000
cap basket_equi        mc_basket_equi_kernel       1 mc_eur_multi 100000000
cap basket_general     "mc_basket_kernel"          1 mc_basket_general 100000000
cap amer_paths         amer_paths_kernel           1 mc_amer 25000000
cap amer_sweep         amer_sweep_kernel           26 mc_amer 100000000
cap binom_screen       binom_terms_kernel          1 binom_embar 2147483647
cap binom_noscreen     binom_terms_kernel          1 binom_embar_noscreen 100000000
cap tree_amer          tree_cta_kernel             900 tree_amer 100000
cap tree_eur           tree_cta_kernel             900 tree_eur 100000
# pipe microbenchmarks: which pipe counter moves for IMAD.WIDE alone, DFMA alone and the mixed loop
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o /tmp/ubench tools/ubench/ubench.cu
timeout 600 ncu --set full --metrics $PM --clock-control none -k regex:"k_imadwide|k_mixed|k_dfma_const|k_mix2|k_lop3" -f -o /tmp/ncu/ubench /tmp/ubench > gpurun_out/r2g_ubench_under_ncu.log 2>&1
{ python tools/ncu_summary.py /tmp/ncu/ubench.ncu-rep "ncu --set full --clock-control none of tools/ubench/ubench.cu: which pipe counters move for IMAD.WIDE alone (k_imadwide), DFMA alone (k_dfma_const), LOP3 alone (k_lop3) and the mixed loops (k_mixed, k_mix2)"
  python tools/ncu_pipes.py /tmp/ncu/ubench.ncu-rep; } > gpurun_out/r2g_ncu_ubench_pipes.txt 2>&1
tail -30 gpurun_out/r2g_ncu_ubench_pipes.txt
du -sh gpurun_out
