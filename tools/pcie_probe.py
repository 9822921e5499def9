#!/usr/bin/env python3
"""Host <-> device copy bandwidth of the box, per GPU and in aggregate (the limiter of bench.py's e2e figure).

    python tools/pcie_probe.py                                              # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_probe.py

Every rank copies the same 32 MiB (one end-to-end call of bench.py moves 32 MiB each way at 2^20 sessions x 32 bytes)
between pinned host memory and its GPU: host->device only, device->host only, both directions at once; all ranks
start together (barrier), so the aggregate is what the host side sustains when every GPU pulls on it.  Also
prints where each GPU and each rank's memory sit (NUMA node of the GPU's PCI device, CPUs the rank may run on) and
repeats the measurement with write-combined pinned input buffers.  One JSON line per rank 0."""
import ctypes
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def numa_of_gpu(index):
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        path = "/sys/bus/pci/devices/%s/numa_node" % bus.lower()[-12:]
        return int(open(path).read().strip())
    except Exception:
        return None


def wc_pinned(nbytes):
    """Write-combined page-locked host memory (cudaHostAllocWriteCombined = 4) as a uint8 tensor, or None."""
    try:
        rt = ctypes.CDLL("libcudart.so")
    except OSError:
        try:
            import glob
            rt = ctypes.CDLL(sorted(glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*")) +
                                    glob.glob("/usr/local/cuda/lib64/libcudart.so*"))[0])
        except Exception:
            return None
    p = ctypes.c_void_p()
    rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
    if rt.cudaHostAlloc(ctypes.byref(p), nbytes, 4) != 0:
        return None
    buf = (ctypes.c_uint8 * nbytes).from_address(p.value)
    return torch.frombuffer(buf, dtype=torch.uint8)


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 32 << 20
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(h_in, h2d, d2h, reps=30):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return n * reps / dt / 1e9

    res = {}
    for kind in ("pinned", "write_combined"):
        h_in = torch.empty(n, dtype=torch.uint8).pin_memory() if kind == "pinned" else wc_pinned(n)
        if h_in is None:
            res[kind] = None
            continue
        run(h_in, True, True, reps=3)
        res[kind] = {"h2d_only": run(h_in, True, False), "d2h_only": run(h_in, False, True), "both_each_way": run(h_in, True, True)}
    mine = {"rank": rank, "gpu": local, "gpu_numa_node": numa_of_gpu(local), "cpus": len(os.sched_getaffinity(0)), **res}
    if world > 1:
        allr = [None] * world
        dist.all_gather_object(allr, mine)
    else:
        allr = [mine]
    if rank == 0:
        agg = {}
        for kind in ("pinned", "write_combined"):
            if all(r.get(kind) for r in allr):
                agg[kind] = {k: sum(r[kind][k] for r in allr) for k in ("h2d_only", "d2h_only", "both_each_way")}
        try:
            nodes = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
        except Exception:
            nodes = []
        print(json.dumps({"world": world, "host_numa_nodes": nodes, "host_cpus": os.cpu_count(), "per_rank": allr,
                          "aggregate_GBps": agg, "bytes_per_copy": n}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
