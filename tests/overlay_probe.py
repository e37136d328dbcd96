"""Subprocess body of tests/test_reference_overlay.py (a fresh interpreter, so `optical_flow` / `model` can be bound to
the reference, to the drop-in packages, or to both in a chosen order).  Prints one JSON line.

    python tests/overlay_probe.py names  <mode> <ref_root>     # import-surface checks, no GPU work
    python tests/overlay_probe.py raft   <mode> <ref_root> <out.pt> [H W iters]   # live RAFT.forward on cuda:0

mode: "stock"  -- the reference alone (its own ATen ops);
      "path"   -- torch-optical-flow_b200/ ahead of the reference on sys.path (overlay packages);
      "patch"  -- the reference imported as usual, then ofb200.patch_reference().
pytorch_lightning / torchmetrics are not installed in this image: they are stubbed with the minimum surface the model
touches -- the reference's files themselves run unmodified.
"""
import json
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "torch-optical-flow_b200")


def stub_third_party():
    import torch  # noqa: F401
    from torch import nn

    if "pytorch_lightning" not in sys.modules:
        try:
            import pytorch_lightning  # noqa: F401
        except Exception:
            pl = types.ModuleType("pytorch_lightning")

            class LightningModule(nn.Module):
                def save_hyperparameters(self):
                    import inspect

                    frame = inspect.currentframe().f_back
                    args = {k: v for k, v in frame.f_locals.items() if k not in ("self", "__class__")}
                    self.hparams = types.SimpleNamespace(**args)

            pl.LightningModule = LightningModule
            pl.LightningDataModule = object
            loggers = types.ModuleType("pytorch_lightning.loggers")
            loggers.WandbLogger = object
            pl.loggers = loggers
            sys.modules["pytorch_lightning"] = pl
            sys.modules["pytorch_lightning.loggers"] = loggers
    if "torchmetrics" not in sys.modules:
        try:
            import torchmetrics  # noqa: F401
        except Exception:
            tm = types.ModuleType("torchmetrics")

            class Metric(nn.Module):
                def add_state(self, name, default, dist_reduce_fx=None):
                    self.register_buffer(name, default.clone())

                def forward(self, *a, **k):
                    return self.update(*a, **k)

            tm.Metric = Metric
            sys.modules["torchmetrics"] = tm
    try:
        import wandb  # noqa: F401
    except Exception:
        wb = types.ModuleType("wandb")
        wb.Image = object
        sys.modules["wandb"] = wb


def set_path(mode, ref):
    ref_paths = [ref, os.path.join(ref, "methods", "raft")]
    for p in (PKG, ROOT):
        while p in sys.path:
            sys.path.remove(p)
    if mode == "path":
        sys.path[:0] = [PKG] + ref_paths
    else:
        sys.path[:0] = ref_paths
        sys.path.append(PKG)          # ofb200 importable; the reference's optical_flow / model stay first


def origin(obj):
    import inspect

    f = inspect.getsourcefile(obj) or ""
    return "ours" if f.startswith(PKG) else ("reference" if f else "?")


def cmd_names(mode, ref):
    stub_third_party()
    set_path(mode, ref)
    import optical_flow
    from optical_flow import colorwheel, flow2rgb, read, write  # noqa: F401  (reference optical_flow/__init__.py:1,3)
    from optical_flow import warp  # noqa: F401
    from optical_flow.metrics import AverageEndPointError, OutlierRatio  # noqa: F401  (reference metrics/__init__.py:1-2)
    import optical_flow.metrics.epe  # noqa: F401
    import model
    import model.raft as raft_mod
    import model.update
    import model.extractor

    patched = []
    if mode == "patch":
        import ofb200

        patched = ofb200.patch_reference()
    import optical_flow.operator.operator as opmod
    import model.corr as corr_mod
    import model.utils as utils_mod

    res = {
        "mode": mode, "patched": patched,
        "optical_flow.warp": origin(optical_flow.warp), "operator.warp": origin(opmod.warp),
        "operator.warp_grid": origin(opmod.warp_grid), "optical_flow.resize": origin(optical_flow.resize),
        "optical_flow.flow2rgb": origin(flow2rgb), "optical_flow.read": origin(read), "optical_flow.write": origin(write),
        "optical_flow.colorwheel": origin(colorwheel),
        "model.RAFT": origin(model.RAFT), "model.RAFT.forward": origin(model.RAFT.forward),
        "RAFT.upsample_flow": origin(model.RAFT.upsample_flow),
        "raft.CorrBlock": origin(raft_mod.__dict__.get("CorrBlock", sys.modules.get("model._reference_raft", raft_mod).CorrBlock)),
        "model.corr.CorrBlock": origin(corr_mod.CorrBlock), "model.utils.bilinear_sampler": origin(utils_mod.bilinear_sampler),
        "model.utils.upflow8": origin(utils_mod.upflow8),
        "model.update.BasicUpdateBlock": origin(model.update.BasicUpdateBlock),
        "model.extractor.BasicEncoder": origin(model.extractor.BasicEncoder),
        "raft.sequence_loss": origin(sys.modules.get("model._reference_raft", raft_mod).sequence_loss),
        "metric_is_torchmetrics": any(c.__name__ == "Metric" for c in AverageEndPointError.__mro__),
        "epe.update": origin(sys.modules["optical_flow.metrics.epe"].AverageEndPointError.update),
    }
    # flow2rgb really runs from the fall-through package (host-side visualisation, CPU only)
    import torch

    rgb = flow2rgb(torch.zeros(2, 4, 5))
    res["flow2rgb_shape"] = list(rgb.shape)
    if mode == "patch":
        import ofb200

        res["unpatched"] = ofb200.unpatch_reference()
        res["after_unpatch.warp"] = origin(optical_flow.warp)
    print(json.dumps(res))


def cmd_raft(mode, ref, out_path, H=436, W=1024, iters=12):
    import torch

    stub_third_party()
    set_path(mode, ref)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    import model
    from model.utils import InputPadder

    if mode == "patch":
        import ofb200

        ofb200.patch_reference()
    if mode != "stock":
        import ofb200

        launches0 = ofb200.launch_count()
    torch.manual_seed(1234)
    net = model.RAFT().cuda().eval()
    g = torch.Generator(device="cuda").manual_seed(99)
    img0 = torch.rand((1, 3, H, W), device="cuda", generator=g) * 255.0
    img1 = torch.roll(img0, shifts=(2, -3), dims=(2, 3)) + 2.0 * torch.randn((1, 3, H, W), device="cuda", generator=g)
    padder = InputPadder(img0.shape)
    img0, img1 = padder.pad(img0, img1)
    times = []
    with torch.no_grad():
        for rep in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            flow_lo, flow_up = net(img0, img1, iters=iters, test_mode=True)
            torch.cuda.synchronize()
            times.append((time.perf_counter() - t0) * 1e3)
        # the hot-path part alone, at the same shapes: one pyramid build + `iters` lookups + one convex upsample
        import model.raft as raft_mod

        ns = sys.modules.get("model._reference_raft", raft_mod)
        fmap1, fmap2 = net.fnet([2 * (img0 / 255.0) - 1.0, 2 * (img1 / 255.0) - 1.0])
        coords0, _ = net.initialize_flow(img0)
        up_mask = torch.randn((1, 576) + tuple(coords0.shape[-2:]), device="cuda", generator=g)
        hot = []
        for rep in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn = ns.CorrBlock(fmap1.float(), fmap2.float(), radius=4)
            for _ in range(iters):
                fn(coords0)
                net.upsample_flow(coords0, up_mask)
            torch.cuda.synchronize()
            hot.append((time.perf_counter() - t0) * 1e3)
    torch.save({"flow_lo": flow_lo.cpu(), "flow_up": padder.unpad(flow_up).cpu()}, out_path)
    res = {"mode": mode, "forward_ms": min(times), "hot_path_ms": min(hot), "iters": iters, "shape": [H, W],
           "pyramid_dtype": os.environ.get("OFB200_PYRAMID_DTYPE", "bf16") if mode != "stock" else "fp32 (torch.matmul)",
           "flow_absmax": float(flow_up.abs().max())}
    if mode != "stock":
        res["ofb_launches"] = ofb200.launch_count() - launches0
    print(json.dumps(res))


if __name__ == "__main__":
    cmd, mode, ref = sys.argv[1:4]
    if cmd == "names":
        cmd_names(mode, ref)
    else:
        extra = [int(x) for x in sys.argv[5:8]]
        cmd_raft(mode, ref, sys.argv[4], *extra)
