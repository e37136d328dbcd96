import ctypes, sys, json
sys.path.insert(0, "tools"); sys.path.insert(0, "torch-optical-flow_b200")
import torch
torch.zeros(1, device="cuda")
rt = ctypes.CDLL("libcudart.so.12")
gran = int(sys.argv[1])
val = ctypes.c_size_t()
rt.cudaDeviceGetLimit(ctypes.byref(val), 5); before = val.value
rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(gran))
rt.cudaDeviceGetLimit(ctypes.byref(val), 5)
print("limit before", before, "set rc", rc, "after", val.value)
import microbench
for shp in [("C5 B4", (4, 256, 136, 240))]:
    recs, blk, gen = microbench.bench_corr(shp[0], *shp[1], cta_groups=(1,), prep=False)
    for r in recs: print(json.dumps(r)[:200])
    print(json.dumps(microbench.bench_lookup(shp[0], blk, gen, shp[1][0], shp[1][2], shp[1][3]))[:200])
for r in microbench.bench_warp_c2(variants=(3,), masks=(1,)): print(json.dumps(r)[:160])
