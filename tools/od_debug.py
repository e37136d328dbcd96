import sys, os
sys.path.insert(0, "torch-optical-flow_b200")
import torch
from model.corr import CorrBlock
from model.utils import coords_grid
gen = torch.Generator(device="cuda").manual_seed(71)
shape=(1,64,19,37)
b,c,h,w=shape
f1 = torch.randn(shape, device="cuda", generator=gen).bfloat16().float()
f2 = torch.randn(shape, device="cuda", generator=gen).bfloat16().float()
od = CorrBlock(f1, f2, on_demand=True)
ref = CorrBlock(f1, f2, pyramid_dtype=torch.float32, builder="simt")
base = coords_grid(b,h,w).cuda()
for name, coords in (("int", base), ("half", base+0.5), ("quarter", base+0.25), ("noise", base + 4*torch.randn((b,2,h,w), device="cuda", generator=gen))):
    g = od(coords); r = ref(coords)
    for l in range(4):
        gl, rl = g[:, l*81:(l+1)*81], r[:, l*81:(l+1)*81]
        print(name, "level", l, "rel", float((gl-rl).norm()/rl.norm()))
    d = (g-r).abs()[0]     # (324,h,w)
    l0 = d[:81].reshape(9,9,h,w)
    print(name, "level0 err by i (x tap):", [round(float(l0[i].max()),3) for i in range(9)])
    print(name, "level0 err by j (y tap):", [round(float(l0[:,j].max()),3) for j in range(9)])
    print(name, "level0 err by query x:", [round(float(l0[:,:,:,x].max()),2) for x in range(0,w,3)])
