# multi-GPU validation (run with gpurun --gpus N): the multi-GPU tests on N >= 2 GPUs, then both bench arms under torchrun
N=${1:-2}
set -x
timeout 1200 python -m pytest tests -x -q -m gpu -k "multi_gpu or in_process" 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2g_bench_${N}gpu_torchrun.json 2> gpurun_out/r2g_bench_${N}gpu.err
tail -c 400 gpurun_out/r2g_bench_${N}gpu_torchrun.json
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2g_bench_${N}gpu_torchrun.json") if l.startswith("{")][-1])
print("headline", d["value"], d["ms_per_step"], "ms")
for o in d["others"]:
    print(o["workload"][:60], round(o["ms_per_step"],3), "ms")
p=d.get("multi_gpu_parity")
print(json.dumps(p)[:1500] if p else "no parity block")
PY
