"""One pass of the hot path over a batch of image pairs, and a host-staged variant that overlaps the
host->device copies of the next micro-batch with the kernels of the current one.

The pass is the sequence the reference's RAFT.forward drives per pair (reference
methods/raft/model/raft.py:112-142) followed by the library operators and metric it feeds
(optical_flow/operator/operator.py:8-33, optical_flow/metrics/epe.py:25-38):

    CorrBlock(fmap1, fmap2)                    K2  prep x3 + tcgen05 pyramid x2 (levels 0-1, levels 2-3)
    corr_fn(coords) x iters                    K3  one launch per refinement iteration
    RAFT.upsample_flow(flow_lo, up_mask)       K4b convex 8x upsampling
    warp(frame, normalize(flow_up)) + mask     K1 (normalize fused: pixel_flow=True)
    AverageEndPointError.update(flow_up, gt)   K4c masked sum / count

The GRU update block between the lookups is out of scope, so its products
(`coords` per iteration, `flow_lo`, `up_mask`) are inputs of the pass.
"""
from typing import Dict, List, Optional

import torch
from torch import Tensor

from ofb200.ops.corr import CorrBlock
from ofb200.ops.epe import AverageEndPointError
from ofb200.ops.operator import warp
from ofb200.ops.raft_ops import upsample_flow

FIELDS = ("fmap1", "fmap2", "coords", "flow_lo", "up_mask", "frame", "target", "valid")


class KernelTimers:
    """CUDA-event brackets on the launching stream, one list per kernel class."""

    def __init__(self) -> None:
        self.spans: Dict[str, List] = {}

    def span(self, name: str, launches: int):
        return _Span(self, name, launches)

    def summary(self) -> Dict[str, Dict[str, float]]:
        torch.cuda.synchronize()
        out = {}
        for name, items in self.spans.items():
            ms = sum(a.elapsed_time(b) for a, b, _ in items)
            n = sum(k for _, _, k in items)
            out[name] = {"launches": n, "ms_total": ms, "ms_per_launch": ms / max(n, 1)}
        return out


class _Span:
    def __init__(self, timers: KernelTimers, name: str, launches: int) -> None:
        self.t, self.name, self.launches = timers, name, launches

    def __enter__(self):
        self.a = torch.cuda.Event(enable_timing=True)
        self.b = torch.cuda.Event(enable_timing=True)
        self.a.record()
        return self

    def __exit__(self, *exc):
        self.b.record()
        self.t.spans.setdefault(self.name, []).append((self.a, self.b, self.launches))
        return False


class _NoSpan:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def hot_path(batch: Dict[str, Tensor], metric: AverageEndPointError, timers: Optional[KernelTimers] = None,
             lookup_out: Optional[Tensor] = None, cta_group: int = 0) -> Dict[str, Tensor]:
    """Run the pass on device tensors.  `batch["coords"]` is (iters, B, 2, h, w)."""
    import ofb200.ops.corr as corr_mod

    sp = (lambda n, k: timers.span(n, k)) if timers is not None else (lambda n, k: _NoSpan())
    corr_mod.TIMERS = timers                       # CorrBlock brackets its prep launches and the pyramid kernel itself
    try:
        blk = CorrBlock(batch["fmap1"], batch["fmap2"], num_levels=4, radius=4, cta_group=cta_group)
    finally:
        corr_mod.TIMERS = None
    iters = batch["coords"].shape[0]
    corr = None
    with sp("lookup", iters):
        for it in range(iters):
            corr = blk(batch["coords"][it], out=lookup_out)
    with sp("convex_upsample", 1):
        flow_up = upsample_flow(batch["flow_lo"], batch["up_mask"])
    with sp("warp", 1):
        warped, vmask = warp(batch["frame"], flow_up, return_mask=True, pixel_flow=True)   # normalize fused
    with sp("epe", 1):
        metric.update(flow_up, batch["target"], batch["valid"])
    return {"corr": corr, "flow_up": flow_up, "warped": warped, "mask": vmask}


LAUNCHES_PER_PASS = lambda iters: 5 + iters + 1 + 1 + 1  # noqa: E731  (prep x3, pyramid x2, lookups, upsample, warp, epe)


class PairArena:
    """Pair-major staging memory: one byte buffer of shape (pairs, bytes_per_pair) in which every field of a pair
    sits at a fixed, 256-byte aligned offset; `arena[name]` is a strided VIEW with the field's usual shape and dtype
    ((pairs, ...), and (iters, pairs, 2, h, w) for `coords`).  A producer (data loader, upstream network) fills the
    views of a pinned host arena; HostStagedRunner then moves a whole micro-batch -- all eight fields -- with ONE
    host->device DMA into a device arena of the same layout instead of one copy per field and per iteration
    (19 copies per pair at the bench shape, each paying its launch + DMA start-up).

    `dtypes` (default fp32 everywhere) lets a field travel in the precision its producer has: the reference's shipped
    configs run `precision: 16` (methods/raft/config/train/default.yaml:20), so the feature maps leave the encoder in
    half precision; CorrBlock reads bf16 / fp16 maps directly (ofb_corr_prep_from), which halves their share of the
    host->device bytes (67 -> 33 MB per 1080p pair)."""

    ALIGN = 256   # bytes

    def __init__(self, pairs: int, shapes: Dict[str, tuple], device=None, pin: bool = False,
                 dtypes: Optional[Dict[str, torch.dtype]] = None) -> None:
        """shapes[name] = the field's per-pair shape; `coords` as (iters, 2, h, w)."""
        self.pairs, self.shapes, self.offsets = pairs, dict(shapes), {}
        self.dtypes = {k: (dtypes or {}).get(k, torch.float32) for k in FIELDS}
        off = 0
        payload = 0
        for k in FIELDS:
            self.offsets[k] = off
            nbytes = int(torch.Size(shapes[k]).numel()) * self.dtypes[k].itemsize
            payload += nbytes
            off += (nbytes + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.pair_bytes = off
        self.payload_bytes_per_pair = payload
        if device is None:
            self.buf = torch.empty((pairs, off), dtype=torch.uint8, pin_memory=pin)
        else:
            self.buf = torch.empty((pairs, off), dtype=torch.uint8, device=device)

    @staticmethod
    def shapes_of(batch: Dict[str, Tensor]) -> Dict[str, tuple]:
        """Per-pair shapes of an ordinary batch dict (coords (iters, pairs, 2, h, w) -> (iters, 2, h, w))."""
        out = {}
        for k in FIELDS:
            sh = tuple(batch[k].shape)
            out[k] = (sh[0],) + sh[2:] if k == "coords" else sh[1:]
        return out

    def same_layout(self, other: "PairArena") -> bool:
        return self.shapes == other.shapes and self.dtypes == other.dtypes

    def view(self, name: str, lo: int = 0, hi: Optional[int] = None) -> Tensor:
        hi = self.pairs if hi is None else hi
        sh = self.shapes[name]
        nbytes = int(torch.Size(sh).numel()) * self.dtypes[name].itemsize
        flat = self.buf[lo:hi, self.offsets[name]:self.offsets[name] + nbytes]
        v = flat.view(self.dtypes[name]).view(hi - lo, *sh)
        return v.transpose(0, 1) if name == "coords" else v

    def __getitem__(self, name: str) -> Tensor:
        return self.view(name)

    def fill(self, batch: Dict[str, Tensor]) -> "PairArena":
        for k in FIELDS:
            self[k].copy_(batch[k])
        return self


class HostStagedRunner:
    """Runs the pass on PINNED HOST batches: micro-batches are copied on a side stream into two
    alternating device slots while the previous micro-batch computes; the EPE state is read back at the
    end.  This is the end-to-end entry point bench.py times (`e2e`).  `run` takes either a dict of pinned
    tensors (one copy per field) or a pinned PairArena (one DMA per micro-batch)."""

    def __init__(self, device: torch.device, micro_pairs: int) -> None:
        self.device = device
        self.micro = micro_pairs
        self.copy_stream = torch.cuda.Stream(device=device)
        self.slots: List[Optional[Dict[str, Tensor]]] = [None, None]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.freed = [torch.cuda.Event(), torch.cuda.Event()]
        self.lookup_out: Optional[Tensor] = None
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self._slot = 0                  # slot the next step's first micro-batch goes to
        self._prefetched = None         # (arena, (lo, hi)) already staged into that slot

    def _slice(self, host: Dict[str, Tensor], lo: int, hi: int) -> Dict[str, Tensor]:
        return {k: (host[k][:, lo:hi] if k == "coords" else host[k][lo:hi]) for k in FIELDS}

    def _stage(self, slot: int, src: Dict[str, Tensor]) -> None:
        if self.slots[slot] is None or any(self.slots[slot][k].shape != src[k].shape for k in FIELDS):
            self.slots[slot] = {k: torch.empty(src[k].shape, dtype=src[k].dtype, device=self.device) for k in FIELDS}
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.freed[slot])       # previous user of this slot has finished
            for k in FIELDS:
                if k == "coords":
                    # (iters, pairs, 2, h, w): a pair range is contiguous per iteration, not as a whole --
                    # copy iteration by iteration so every transfer is one pinned, asynchronous DMA
                    for it in range(src[k].shape[0]):
                        self.slots[slot][k][it].copy_(src[k][it], non_blocking=True)
                else:
                    self.slots[slot][k].copy_(src[k], non_blocking=True)
                self.h2d_bytes += src[k].numel() * src[k].element_size()
            self.ready[slot].record(self.copy_stream)

    def _stage_arena(self, slot: int, host: PairArena, lo: int, hi: int) -> None:
        cur = self.slots[slot]
        if not isinstance(cur, PairArena) or cur.pairs != hi - lo or not cur.same_layout(host):
            self.slots[slot] = PairArena(hi - lo, host.shapes, device=self.device, dtypes=host.dtypes)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.freed[slot])       # previous user of this slot has finished
            self.slots[slot].buf.copy_(host.buf[lo:hi], non_blocking=True)   # whole rows: one contiguous DMA
            self.h2d_bytes += (hi - lo) * host.pair_bytes           # what the DMA moves (payload + <= 2 KB of alignment padding per pair)
            self.ready[slot].record(self.copy_stream)

    def _run_arena(self, host: PairArena, metric: AverageEndPointError, prefetch: Optional[PairArena]) -> float:
        if host.buf.is_cuda or not host.buf.is_pinned():
            raise RuntimeError("HostStagedRunner: the arena must live in pinned host memory")
        pairs = host.pairs
        chunks = [(lo, min(lo + self.micro, pairs)) for lo in range(0, pairs, self.micro)]
        cur = torch.cuda.current_stream(self.device)
        first = self._slot
        if self._prefetched is not None and self._prefetched[0] is host and self._prefetched[1] == chunks[0]:
            pass                                                # the previous call already staged this micro-batch
        else:
            self._stage_arena(first, host, *chunks[0])
        self._prefetched = None
        for i, _ in enumerate(chunks):
            slot = (first + i) & 1
            if i + 1 < len(chunks):
                self._stage_arena(slot ^ 1, host, *chunks[i + 1])
            elif prefetch is not None:
                # the next step's first micro-batch goes out while this step's last one computes, so the copy
                # engine never idles across the per-step read-back of the metric
                nxt = (0, min(self.micro, prefetch.pairs))
                self._stage_arena(slot ^ 1, prefetch, *nxt)
                self._prefetched = (prefetch, nxt)
            cur.wait_event(self.ready[slot])
            arena = self.slots[slot]
            dev = {k: arena[k] for k in FIELDS}                 # views; contiguous per field when the micro-batch is 1 pair
            b, _, h, w = dev["fmap1"].shape
            if self.lookup_out is None or self.lookup_out.shape[0] != b or self.lookup_out.shape[2:] != (h, w):
                self.lookup_out = torch.empty((b, 324, h, w), dtype=torch.float32, device=self.device)
            hot_path(dev, metric, lookup_out=self.lookup_out)
            self.freed[slot].record(cur)
        self._slot = (first + len(chunks)) & 1
        state = metric.sync().cpu()                             # all-reduced COPY of (sum, count), read back to the host
        self.d2h_bytes += state.numel() * state.element_size()
        return float(state[0] / state[1])

    def run(self, host, metric: AverageEndPointError, prefetch: Optional[PairArena] = None) -> float:
        """One step over `host`.  `prefetch` (arena path): the arena of the NEXT step; its first micro-batch is
        staged while this step finishes and the next `run(prefetch, ...)` picks it up."""
        if isinstance(host, PairArena):
            return self._run_arena(host, metric, prefetch)
        self._prefetched = None
        self._slot = 0
        for k in FIELDS:
            if host[k].is_cuda or not host[k].is_pinned():
                raise RuntimeError(f"HostStagedRunner: {k} must be a pinned host tensor")
        pairs = host["fmap1"].shape[0]
        chunks = [(lo, min(lo + self.micro, pairs)) for lo in range(0, pairs, self.micro)]
        cur = torch.cuda.current_stream(self.device)
        self._stage(0, self._slice(host, *chunks[0]))
        for i, _ in enumerate(chunks):
            slot = i & 1
            if i + 1 < len(chunks):
                self._stage(slot ^ 1, self._slice(host, *chunks[i + 1]))
            cur.wait_event(self.ready[slot])
            dev = self.slots[slot]
            b, _, h, w = dev["fmap1"].shape
            if self.lookup_out is None or self.lookup_out.shape[0] != b or self.lookup_out.shape[2:] != (h, w):
                self.lookup_out = torch.empty((b, 324, h, w), dtype=torch.float32, device=self.device)
            hot_path(dev, metric, lookup_out=self.lookup_out)
            self.freed[slot].record(cur)
        state = metric.sync().cpu()                      # all-reduced COPY of (sum, count), read back to the host
        self.d2h_bytes += state.numel() * state.element_size()
        return float(state[0] / state[1])


def choose_device_set(rank: int, world: int, local_rank: int, probe_mb: int = 64, reps: int = 6, min_gain: float = 1.10):
    """Topology-aware rank -> GPU placement for host-staged runs (one process per GPU, `torch.distributed` initialised
    with a CPU-capable backend such as "cpu:gloo,cuda:nccl").

    On an 8-GPU B200 host the GPUs do not all reach pinned host memory at the same rate: measured on this pool
    (tools/h2d_probe.py, profiles/r02_scale_probe.jsonl) four ranks on GPUs 0-3 share ~115 GB/s while GPUs 4-7 take
    55 GB/s EACH (220 GB/s) -- and nothing inside the guest (one NUMA node, no PCIe topology) says which half is
    which.  So when more GPUs are visible than ranks, the two candidate sets {0..N-1} and {V-N..V-1} are probed:
    every rank copies `probe_mb` from pinned memory to its GPU of the set at the same time, the slowest rank's time
    is agreed on through the CPU backend, and the faster set is taken if it is at least `min_gain` times better.
    Returns (device index for this rank, a dict describing the decision)."""
    import torch.distributed as dist

    visible = torch.cuda.device_count()
    info = {"visible": visible, "sets": {}, "chosen": "default"}
    if world < 2 or visible < 2 * world or not dist.is_initialized():
        return local_rank, info
    sets = {"low": list(range(world)), "high": list(range(visible - world, visible))}
    nbytes = probe_mb << 20
    host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    rates = {}
    for name, devs in sets.items():
        dev = torch.device("cuda", devs[local_rank])
        with torch.cuda.device(dev):
            buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            st = torch.cuda.Stream(device=dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(st):
                buf.copy_(host, non_blocking=True)                      # warm-up: context, link power state
            st.synchronize()
            dist.all_reduce(torch.zeros(1))                             # CPU tensor -> CPU backend: all ranks start together
            with torch.cuda.stream(st):
                e0.record(st)
                for _ in range(reps):
                    buf.copy_(host, non_blocking=True)
                e1.record(st)
            st.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)                    # CPU tensor -> CPU backend
            rates[name] = world * nbytes * reps / (float(t.item()) * 1e-3) / 1e9
            del buf
        info["sets"][name] = {"devices": devs, "aggregate_h2d_gbs": round(rates[name], 1)}
    pick = "high" if rates["high"] >= min_gain * rates["low"] else "low"
    info["chosen"] = pick
    return sets[pick][local_rank], info


def shard_range(total: int, rank: int, world: int):
    """Contiguous batch slice [lo, hi) of `total` pairs owned by `rank` (pairs never cross GPUs)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch: Dict[str, Tensor], rank: int, world: int) -> Dict[str, Tensor]:
    lo, hi = shard_range(batch["fmap1"].shape[0], rank, world)
    return {k: (v[:, lo:hi] if k == "coords" else v[lo:hi]) for k, v in batch.items()}
