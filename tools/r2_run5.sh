set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
PM=sm__inst_executed_pipe_fp64,sm__inst_executed_pipe_fmaheavy,sm__inst_executed_pipe_fmalite,sm__inst_executed_pipe_alu,sm__inst_executed_pipe_xu,sm__inst_executed_pipe_lsu,sm__pipe_fp64_cycles_active,sm__inst_executed,sm__cycles_active,smsp__issue_active
timeout 600 ncu --set full --metrics $PM --clock-control none --import-source on -k regex:amer_sweep_kernel -s 26 -c 1 -f -o gpurun_out/r2f_ncu_amer_sweep python tools/ncu_target.py mc_amer 100000000 1 2>&1 | tail -2
