# ncu launch list of the bench command itself (per-kernel shares and DRAM bytes per launch -> profiles/ncu_traffic.json)
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2g_bench_under_ncu.log 2>&1
tail -c 300 gpurun_out/r2g_bench_under_ncu.log
python tools/summarize_launches.py gpurun_out/r2g_launches.csv gpurun_out/r2g_launch_list_summary.txt gpurun_out/ncu_traffic.json | head -30
du -sh gpurun_out
