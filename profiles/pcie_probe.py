#!/usr/bin/env python
"""Host-link probe for the end-to-end leg at N > 1: device->host bandwidth into pinned memory from every
GPU alone and from all GPUs at once, with and without binding the rank to the GPU's NUMA node before the
pinned buffer is allocated.  Launch with torchrun (one rank per GPU); rank 0 prints one JSON line per case.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 profiles/pcie_probe.py
"""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nesti_net_b200 import dist as mdist  # noqa: E402


def d2h_gbps(dev_buf, host_buf, reps=6):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        host_buf.copy_(dev_buf, non_blocking=True)
    s.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        e0.record(s)
        for _ in range(reps):
            host_buf.copy_(dev_buf, non_blocking=True)
        e1.record(s)
    s.synchronize()
    return dev_buf.numel() * dev_buf.element_size() * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    info = mdist.gpu_numa_info(local)
    info["affinity_before"] = len(os.sched_getaffinity(0))
    n = 1 << 28                                            # 1 GiB of fp32
    dbuf = torch.empty(n, dtype=torch.float32, device=dev).normal_()

    def gather(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world == 1:
            return [float(x)]
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    def case(name, hbuf):
        alone = []
        for r in range(world):
            if world > 1:
                dist.barrier()
            alone.append(d2h_gbps(dbuf, hbuf) if r == rank else 0.0)
        alone = [max(v) for v in zip(*[gather(a) for a in alone])] if world > 1 else alone
        if world > 1:
            dist.barrier()
        together = gather(d2h_gbps(dbuf, hbuf))
        if rank == 0:
            print(json.dumps({"case": name, "alone_GBps": [round(a, 1) for a in alone],
                              "together_GBps": [round(a, 1) for a in together], "sum_together": round(sum(together), 1)}), flush=True)

    infos = [None] * world
    if world > 1:
        dist.all_gather_object(infos, info)
    else:
        infos = [info]
    if rank == 0:
        print(json.dumps({"gpus": infos}), flush=True)
        try:
            print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout, file=sys.stderr)
        except Exception:
            pass
    case("unbound", torch.empty(n, dtype=torch.float32, pin_memory=True))
    bound = mdist.bind_to_gpu_numa_node(local)
    hb = torch.empty(n, dtype=torch.float32, pin_memory=True)
    hb.zero_()
    case("bound_to_gpu_numa_node" if bound else "bind_unavailable", hb)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
