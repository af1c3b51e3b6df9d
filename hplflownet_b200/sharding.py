"""Sample-level sharding across GPUs.

Every cloud pair owns its lattice and hash table, so the path shards by sample with no data-path
collective: one process per GPU, rank r works on its own clouds.  torch.distributed is used only for
the barrier and the max-over-ranks reduction of timings (NCCL on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def cloud_ids(rank, world_size, clouds_per_rank, first=0):
    """Global ids (= synthetic seeds) of the clouds rank ``rank`` processes in one step (weak scaling:
    every rank has ``clouds_per_rank`` of its own)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    start = first + rank * clouds_per_rank
    return list(range(start, start + clouds_per_rank))


def split_strong(n_items, rank, world_size):
    """Contiguous slice [lo, hi) of ``n_items`` for ``rank`` (strong scaling: fixed total work)."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(values, device="cpu"):
    """Element-wise max of a list of floats over all ranks (the job's time is its slowest rank's)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def job_throughput(units_per_rank, world_size, seconds_max_over_ranks):
    """Whole-job units/s: all ranks' units divided by the slowest rank's time."""
    return units_per_rank * world_size / seconds_max_over_ranks


def bind_to_gpu_numa_node(local_rank):
    """Pin this process (and therefore the pinned host buffers it allocates next: first-touch placement) to the CPUs
    of the NUMA node its GPU hangs off.  With one process per GPU on a multi-socket host this keeps every rank's
    host-to-device traffic on its own socket.  Best effort: returns the node id, or None when the topology cannot be
    read (no sysfs entry, single node, insufficient rights)."""
    import os
    try:
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev)
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:                                            # noqa: BLE001 -- topology is optional information
        return None
