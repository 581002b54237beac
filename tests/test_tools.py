"""tools/summarize_launches.py turns an ncu launch list into profiles/ncu_traffic.json (bench.py's roofline.traffic)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CSV = '''==PROF== Connected to process 1
"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size","Device","CC","Section Name","Metric Name","Metric Unit","Metric Value"
"0","1","python","h","void pcf::mc_asia_kernel<(bool)1, (int)6, (int)1>(pcf::AsiaArgs)","1","7","(256, 1, 1)","(148, 1, 1)","0","10.0","Command line profiler metrics","dram__bytes_read.sum","Kbyte","800.5"
"0","1","python","h","void pcf::mc_asia_kernel<(bool)1, (int)6, (int)1>(pcf::AsiaArgs)","1","7","(256, 1, 1)","(148, 1, 1)","0","10.0","Command line profiler metrics","dram__bytes_write.sum","byte","1,024"
"0","1","python","h","void pcf::mc_asia_kernel<(bool)1, (int)6, (int)1>(pcf::AsiaArgs)","1","7","(256, 1, 1)","(148, 1, 1)","0","10.0","Command line profiler metrics","gpu__time_duration.sum","ms","700.0"
"1","1","python","h","void pcf::amer_sweep_kernel<unsigned char, (bool)1>(pcf::SweepArgs)","1","7","(288, 1, 1)","(444, 1, 1)","0","10.0","Command line profiler metrics","dram__bytes_read.sum","Gbyte","2.0"
"1","1","python","h","void pcf::amer_sweep_kernel<unsigned char, (bool)1>(pcf::SweepArgs)","1","7","(288, 1, 1)","(444, 1, 1)","0","10.0","Command line profiler metrics","dram__bytes_write.sum","Mbyte","100"
"1","1","python","h","void pcf::amer_sweep_kernel<unsigned char, (bool)1>(pcf::SweepArgs)","1","7","(288, 1, 1)","(444, 1, 1)","0","10.0","Command line profiler metrics","gpu__time_duration.sum","us","400"
"2","1","python","h","void pcf::amer_sweep_kernel<unsigned char, (bool)1>(pcf::SweepArgs)","1","7","(288, 1, 1)","(444, 1, 1)","0","10.0","Command line profiler metrics","dram__bytes_read.sum","Gbyte","2.2"
"2","1","python","h","void pcf::amer_sweep_kernel<unsigned char, (bool)1>(pcf::SweepArgs)","1","7","(288, 1, 1)","(444, 1, 1)","0","10.0","Command line profiler metrics","dram__bytes_write.sum","Mbyte","100"
"2","1","python","h","void pcf::amer_sweep_kernel<unsigned char, (bool)1>(pcf::SweepArgs)","1","7","(288, 1, 1)","(444, 1, 1)","0","10.0","Command line profiler metrics","gpu__time_duration.sum","us","420"
'''


def test_summarize_launches_units_and_aggregation(tmp_path):
    src = tmp_path / "launches.csv"
    src.write_text(CSV)
    txt, js = tmp_path / "summary.txt", tmp_path / "traffic.json"
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "summarize_launches.py"), str(src), str(txt), str(js)],
                   check=True, stdout=subprocess.DEVNULL)
    k = json.load(open(js))["kernels"]
    asia = next(v for n, v in k.items() if n.startswith("mc_asia_kernel"))
    sweep = next(v for n, v in k.items() if n.startswith("amer_sweep_kernel"))
    assert asia["launches"] == 1 and abs(asia["dram_bytes_per_launch"] - (800.5e3 + 1024)) < 1e-6
    assert abs(asia["avg_ms_under_ncu"] - 700.0) < 1e-9
    assert sweep["launches"] == 2 and abs(sweep["dram_bytes_per_launch"] - (2.1e9 + 100e6)) < 1.0
    assert abs(sweep["avg_ms_under_ncu"] - 0.41) < 1e-9
    lines = txt.read_text().splitlines()
    assert lines[0].startswith("mc_asia_kernel") and "share=0.999" in lines[0]


def test_bench_reads_the_committed_traffic_file():
    sys.path.insert(0, ROOT)
    import bench
    t = bench.ncu_traffic("amer_sweep_kernel+amer_paths_kernel")
    assert t is not None and t > 1e9          # the path kernel writes ~40 GB per launch
    assert bench.ncu_traffic("no_such_kernel") is None


def test_bench_reads_the_committed_ncu_digests():
    """`roofline.ncu` of every bench line comes from profiles/r2g_ncu_<kernel>.txt (tools/ncu_summary.py output)."""
    sys.path.insert(0, ROOT)
    import bench
    for name in bench._NCU_DIGEST:
        assert name in bench.WORKLOADS, name
        c = bench.ncu_counters(name)
        assert c is not None and os.path.exists(os.path.join(ROOT, c["source"])), name
        assert 0.0 < c["fp64_pipe_pct"] <= 100.0 and 0.0 < c["issue_slots_pct"] <= 100.0, (name, c)
    assert bench.ncu_counters("mc_amer")["dram_throughput_pct"] > 50.0   # the one HBM-bound line
    assert bench.ncu_counters("no_such_workload") is None
    # the step traffic of mc_amer: four kernels, ~157 GB per step at 1e8 x 50
    t = bench.ncu_traffic(bench.WORKLOADS["mc_amer"]["kernel"], all_kernels=True)
    assert 1.4e11 < t < 1.8e11
