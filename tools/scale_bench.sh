#!/bin/bash
# bench.py at N = 2, 4, 8 on one 8-GPU box (gpurun --gpus 8), with the host-link placement probe.
port=29700
for n in 2 4 8; do
  port=$((port+1))
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --steps 10 --warmup 3 2>gpurun_out/r02_bench_n$n.err | tail -1 > gpurun_out/r02_bench_n$n.json
done
python - <<'PY'
import json
for n in (2, 4, 8):
    try:
        d = json.load(open(f"gpurun_out/r02_bench_n{n}.json"))
        print(n, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e_fp32_inputs"]["value"], d["config"]["device_placement"], d["clocks"])
    except Exception as e:
        print(n, "failed", e)
        print(open(f"gpurun_out/r02_bench_n{n}.err").read()[-1500:])
PY
