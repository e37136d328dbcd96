#!/bin/bash
# Concurrent host->device ceilings and bench.py at N = 4, 8 on one 8-GPU box (gpurun --gpus 8).
# Output: gpurun_out/r02_scale_probe.jsonl (one JSON line per run), gpurun_out/r02_bench_n{4,8}.json
out=gpurun_out/r02_scale_probe.jsonl
: > $out
python tools/h2d_probe.py topo > gpurun_out/r02_topo_n8.json 2>/dev/null
port=29600
run_probe() {  # $1 = CUDA_VISIBLE_DEVICES list
  n=$(echo $1 | tr ',' '\n' | wc -l)
  port=$((port+1))
  CUDA_VISIBLE_DEVICES=$1 timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
      --master-port $port tools/h2d_probe.py 162 20 2>/dev/null | tail -1 >> $out
}
for set in 0 4 0,1 0,4 2,3 6,7 0,1,2,3 4,5,6,7 0,2,4,6 0,1,2,3,4,5,6,7; do run_probe $set; done
for n in 4 8; do
  port=$((port+1))
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --steps 10 --warmup 3 2>gpurun_out/r02_bench_n$n.err | tail -1 > gpurun_out/r02_bench_n$n.json
done
cat $out | cut -c1-330
python - <<'PY'
import json
for n in (4, 8):
    try:
        d = json.load(open(f"gpurun_out/r02_bench_n{n}.json"))
        print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e_fp32_inputs"]["value"], d["clocks"])
    except Exception as e:
        print(n, "failed", e)
PY
