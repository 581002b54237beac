# round-2 validation set on one B200 (run through gpurun from the repo root): tests, sanitizer, bench, reference arm, smoke
set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_target.py > gpurun_out/r2_compute_sanitizer_memcheck.log 2>&1; tail -3 gpurun_out/r2_compute_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_target.py > gpurun_out/r2_compute_sanitizer_racecheck.log 2>&1; tail -3 gpurun_out/r2_compute_sanitizer_racecheck.log
timeout 900 python bench.py > gpurun_out/r2g_bench_1gpu.json 2> gpurun_out/r2g_bench_1gpu.err; tail -c 300 gpurun_out/r2g_bench_1gpu.json
timeout 600 python bench.py --impl reference > gpurun_out/r2g_bench_reference_arm.json 2>/dev/null; head -c 400 gpurun_out/r2g_bench_reference_arm.json
python __graft_entry__.py smoke 2>&1 | tail -2
