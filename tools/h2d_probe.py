"""Concurrent host->device copy ceiling of one box: every rank copies pinned host pieces of the bench's size to its own
GPU at the same time (barrier start, CUDA events, max over ranks), for plain pinned memory and for write-combined pinned
memory (cudaHostAllocWriteCombined: the DMA engine does not snoop CPU caches).  Explains the end-to-end scaling curve of
bench.py: `e2e` at N GPUs cannot exceed aggregate_GBps / bytes_per_pair.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/h2d_probe.py [piece_mb] [reps]
    python tools/h2d_probe.py topo        # PCI bus id / NUMA node / CPU list of every GPU + nvidia-smi topo -m
One JSON line per run (rank 0)."""
import ctypes
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist


def topo():
    out = {"gpus": []}
    for i in range(torch.cuda.device_count()):
        pr = torch.cuda.get_device_properties(i)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        rec = {"index": i, "pci": bus}
        for f in ("numa_node", "local_cpulist", "current_link_speed", "current_link_width"):
            try:
                rec[f] = open(f"/sys/bus/pci/devices/{bus}/{f}").read().strip()
            except OSError:
                rec[f] = None
        out["gpus"].append(rec)
    try:
        out["topo_m"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout
    except Exception as e:
        out["topo_m"] = repr(e)
    try:
        out["host_cpus"] = len(os.sched_getaffinity(0))
        out["numa_nodes"] = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
    except OSError:
        pass
    return out


def host_buffer(nbytes, write_combined):
    """Pinned host memory as a torch uint8 tensor; write-combined through cudaHostAlloc (flag 0x04)."""
    if not write_combined:
        return torch.empty(nbytes, dtype=torch.uint8, pin_memory=True), None
    rt = ctypes.CDLL("libcudart.so.12")
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(0x04))
    if rc != 0:
        raise RuntimeError(f"cudaHostAlloc(write-combined) failed: {rc}")
    arr = (ctypes.c_uint8 * nbytes).from_address(p.value)
    return torch.frombuffer(arr, dtype=torch.uint8), (rt, p)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "topo":
        print(json.dumps(topo()))
        return
    piece_mb = float(sys.argv[1]) if len(sys.argv) > 1 else 162.0
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nbytes = int(piece_mb * 1e6) // 256 * 256
    res = {"world": world, "piece_mb": round(nbytes / 1e6, 1), "reps": reps, "visible": os.environ.get("CUDA_VISIBLE_DEVICES", "all")}
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    for name, wc in (("pinned", False), ("pinned_write_combined", True)):
        try:
            h, keep = host_buffer(2 * nbytes, wc)
        except Exception as e:
            res[name] = {"error": repr(e)[:120]}
            continue
        h[:nbytes].fill_(1)
        h[nbytes:].fill_(2)
        st = torch.cuda.Stream(device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        rt = ctypes.CDLL("libcudart.so.12")

        def copy(k):
            # one cudaMemcpyAsync per piece, as ofb200.runner does for a pair
            rc = rt.cudaMemcpyAsync(ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(h.data_ptr() + (k & 1) * nbytes),
                                    ctypes.c_size_t(nbytes), ctypes.c_int(1), ctypes.c_void_p(st.cuda_stream))
            assert rc == 0, rc
        for k in range(3):
            copy(k)
        st.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record(st)
        for k in range(reps):
            copy(k)
        e1.record(st)
        st.synchronize()
        ms = e0.elapsed_time(e1)
        mine = nbytes * reps / ms / 1e6
        t = torch.tensor([ms, mine], device=dev, dtype=torch.float64)
        per_rank = [mine]
        if world > 1:
            tl = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(tl, t)
            ms = max(float(x[0]) for x in tl)
            per_rank = [round(float(x[1]), 1) for x in tl]
        res[name] = {"aggregate_gbs": round(world * nbytes * reps / ms / 1e6, 1), "per_rank_gbs": per_rank,
                     "pairs_per_s_ceiling_at_this_piece": round(world * reps / (ms * 1e-3), 1)}
        ok = bool((d[:16] == (1 if (reps - 1) & 1 == 0 else 2)).all())
        res[name]["data_ok"] = ok
        del h
        if keep is not None:
            keep[0].cudaFreeHost(keep[1])
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
