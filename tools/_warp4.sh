timeout 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -3
python bench.py > gpurun_out/bench_warp2.json 2> gpurun_out/bench_warp2.err; tail -c 1500 gpurun_out/bench_warp2.json
timeout 300 python - <<'PY' 2>&1 | grep kernel
import sys, os; sys.path.insert(0, "tools"); sys.path.insert(0, "torch-optical-flow_b200")
import microbench as mb
for r in mb.bench_warp_c2(variants=(0, 4), flows=("white5px", "smooth5px", "smooth5px64"), masks=(0, 1)): print(r)
PY
