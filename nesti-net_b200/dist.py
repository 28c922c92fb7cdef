"""Multi-GPU partitioning of the MuPS path.

The reference is single-process; the path shards naturally: every query point is independent
given the read-only cloud.  One process per GPU (torch.distributed, NCCL on GPUs / gloo in the
CPU tests), the cloud and GMM are replicated on every rank (each rank builds its own index: the
build is far cheaper than shipping it), the query list is split into contiguous ranges and each
rank writes its own feature slab.  There is NO collective on the data path; ``gather_slabs`` is
the optional final all-gather for a single-rank consumer.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_items, world_size, weights=None):
    """Contiguous split of ``n_items`` into ``world_size`` ranges -> int64 [world_size + 1].

    Without weights the ranges differ by at most one item.  With per-item ``weights`` (estimated
    work, e.g. the neighbour count at the largest radius of a cheap count pass) the cut points
    balance the prefix sum of the weights -- needed for clouds of non-uniform density."""
    n_items, world_size = int(n_items), int(world_size)
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    if weights is None:
        base, extra = divmod(n_items, world_size)
        sizes = np.full(world_size, base, dtype=np.int64)
        sizes[:extra] += 1
        return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    w = np.asarray(weights, dtype=np.float64).reshape(-1)
    if len(w) != n_items:
        raise ValueError("weights must have one entry per item")
    if np.any(w < 0):
        raise ValueError("weights must be non-negative")
    csum = np.cumsum(w)
    total = csum[-1] if n_items else 0.0
    if total <= 0:
        return shard_bounds(n_items, world_size)
    targets = total * np.arange(1, world_size, dtype=np.float64) / world_size
    cuts = np.searchsorted(csum, targets, side="left") + 1
    bounds = np.concatenate([[0], np.minimum(cuts, n_items), [n_items]]).astype(np.int64)
    return np.maximum.accumulate(bounds)


def shard_queries(query_idx, rank, world_size, weights=None):
    """This rank's contiguous slice of the query list and its (start, end) in the full list."""
    b = shard_bounds(len(query_idx), world_size, weights)
    return query_idx[b[rank]:b[rank + 1]], int(b[rank]), int(b[rank + 1])


def gather_slabs(slab, group=None):
    """All-gather per-rank feature slabs [B_r, ...] of unequal B_r along dim 0 (the only
    collective of the path; only for a single-rank consumer).  Works with NCCL (CUDA slabs) and
    gloo (CPU slabs)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return slab
    world = dist.get_world_size(group)
    count = torch.tensor([slab.shape[0]], dtype=torch.int64, device=slab.device)
    counts = [torch.zeros_like(count) for _ in range(world)]
    dist.all_gather(counts, count, group=group)
    counts = [int(c.item()) for c in counts]
    bmax = max(counts)
    padded = slab
    if slab.shape[0] < bmax:
        pad = torch.zeros((bmax - slab.shape[0],) + tuple(slab.shape[1:]), dtype=slab.dtype, device=slab.device)
        padded = torch.cat([slab, pad], dim=0)
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded.contiguous(), group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


# --------------------------------------------------------------------------------------------------
# host placement: keep a rank's pinned staging buffers on the NUMA node its GPU hangs off
# --------------------------------------------------------------------------------------------------

def gpu_numa_info(device_index):
    """PCI address and NUMA node of a CUDA device from sysfs (node -1: the platform does not say)."""
    import os
    info = {"device": int(device_index), "pci": None, "numa_node": -1, "cpus": None}
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        info["pci"] = bdf
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            info["numa_node"] = int(f.read().strip())
        if info["numa_node"] >= 0:
            with open("/sys/devices/system/node/node%d/cpulist" % info["numa_node"]) as f:
                info["cpus"] = f.read().strip()
    except Exception:
        pass
    info["online_nodes"] = None
    try:
        with open("/sys/devices/system/node/online") as f:
            info["online_nodes"] = f.read().strip()
    except Exception:
        pass
    return info


def _parse_cpulist(text):
    cpus = set()
    for part in text.split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            lo, hi = part.split("-")
            cpus.update(range(int(lo), int(hi) + 1))
        else:
            cpus.add(int(part))
    return cpus


def bind_to_gpu_numa_node(device_index):
    """Restrict this process to the CPUs of the GPU's NUMA node (first-touch then places the pinned
    staging buffers there, so device->host copies do not cross the socket interconnect).  Returns True
    when a binding was applied; silently does nothing where sysfs gives no node."""
    import os
    info = gpu_numa_info(device_index)
    if info["numa_node"] < 0 or not info["cpus"]:
        return False
    try:
        allowed = os.sched_getaffinity(0)
        want = _parse_cpulist(info["cpus"]) & allowed
        if not want:
            return False
        os.sched_setaffinity(0, want)
        return True
    except Exception:
        return False


# --------------------------------------------------------------------------------------------------
# gather through peer memory: the statistics kernel's own stores are the collective
# --------------------------------------------------------------------------------------------------

class PeerSlabGather(object):
    """Single-rank consumer of the full MuPS tensor without a separate collective.

    Every rank allocates the same symmetric buffer [total_rows, *row_shape]; the rendezvous maps each
    rank's buffer into every process over NVLink / NVSwitch peer memory.  Rank r then passes
    ``target(lo, hi)`` -- rows [lo, hi) of the CONSUMER's buffer -- as the ``out`` of its statistics
    launch: the kernel epilogue's float4 stores travel to the consumer as they are issued, CTA by CTA,
    overlapped with the arithmetic of the CTAs still running, and nothing is staged locally or re-sent.
    ``finish()`` (stream sync + a barrier over the signal pads) makes the slabs visible; the consumer
    reads ``result()``.  Bit-identical to the single-GPU tensor (there is no reduction).

    Uses torch's symmetric-memory allocator (cuMem + peer mapping) for the plumbing; raises
    RuntimeError where the platform cannot map peer memory between processes."""

    def __init__(self, total_rows, row_shape, dst=0, group=None, device=None, dtype=torch.float32):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerSlabGather needs an initialised process group")
        try:
            import torch.distributed._symmetric_memory as symm_mem
        except Exception as e:  # pragma: no cover
            raise RuntimeError("symmetric memory is not available in this torch build: %s" % (e,))
        self.group = dist.group.WORLD if group is None else group
        self.dst = int(dst)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.shape = (int(total_rows),) + tuple(int(x) for x in row_shape)
        self.dtype = dtype
        try:
            self.local = symm_mem.empty(self.shape, dtype=dtype, device=self.device)
            self.handle = symm_mem.rendezvous(self.local, self.group)
            self.remote = self.handle.get_buffer(self.dst, self.shape, dtype)
        except Exception as e:
            raise RuntimeError("peer-memory rendezvous failed (%s: %s)" % (type(e).__name__, e))

    def target(self, lo, hi):
        """Rows [lo, hi) of the consumer's buffer, addressable from this rank's kernels."""
        return self.remote[int(lo):int(hi)]

    def finish(self):
        """All slabs written before this call on any rank are visible at the consumer after it."""
        torch.cuda.current_stream(self.device).synchronize()
        self.handle.barrier()

    def result(self):
        """The consumer's full tensor (valid after finish()); other ranks get their own, unused buffer."""
        return self.local
