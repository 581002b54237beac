set -x
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 900 python bench.py > gpurun_out/r1s_bench_1gpu.json 2> gpurun_out/r1s_bench_1gpu.err; tail -c 600 gpurun_out/r1s_bench_1gpu.json
timeout 600 python bench.py --impl reference > gpurun_out/r1s_bench_reference_arm.json 2>/dev/null; cat gpurun_out/r1s_bench_reference_arm.json | head -c 800
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r1s_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r1s_bench_under_ncu.log 2>&1
for k in eur amer; do timeout 200 ncu --set full --clock-control none --import-source on -k regex:"tree_cta" -s 160 -c 1 -o gpurun_out/prof_tree_cta_$k python tools/tree_target.py $k 100000 2>&1 | tail -1; done
python __graft_entry__.py smoke 2>&1 | tail -2
