"""Multi-GPU evaluation: one process per GPU, disjoint contiguous image shards, replicated weights,
and ONE collective for the whole job -- a sum all-reduce of the int64 count tensor
(3 x 6 x 2 x 2 = 72 elements = 576 bytes) over NCCL / NVLink.  Integer sums are
order independent, so the reduced tensor is bit-identical to a single-GPU run for any world size.
The same code runs on the gloo backend for the CPU tests of the host logic.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous, balanced partition: the first ``n_items % world_size`` ranks get one extra item."""
    if n_items < 0 or world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad shard arguments")
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batches(n_images: int, batch: int, rank: int, world_size: int) -> tuple[int, int, int]:
    """Sharding by WHOLE batches: the job's ceil(n_images / batch) global batches are dealt to the ranks in contiguous
    balanced ranges -> (first batch, one past the last batch, number of global batches).  Every image then sits in the
    same batch, at the same row, whatever the world size, so per-image results (and therefore the summed counts) are
    bit-identical between a sharded run and a single-GPU run; only the last global batch can be ragged."""
    if batch < 1:
        raise ValueError("bad batch size")
    n_batches = -(-n_images // batch)
    lo, hi = shard_range(n_batches, rank, world_size)
    return lo, hi, n_batches


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """(rank, world_size, local_rank) from torchrun's environment; initialises the process group when
    WORLD_SIZE > 1."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def allreduce_counts(counts: torch.Tensor) -> torch.Tensor:
    """In-place sum over all ranks of the per-group count tensor (int64).  No-op for one process."""
    if counts.dtype != torch.int64:
        raise TypeError("counts must be int64")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts


def max_over_ranks(value: float, device=None) -> float:
    """Max of a host scalar over ranks (used for device-timed durations)."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def bind_to_gpu_numa_node(local_rank: int) -> dict:
    """Pins this process (and therefore the pinned host buffers it allocates afterwards, by first touch) to the
    CPUs of the NUMA node the GPU hangs off, so that host->device copies of N ranks do not cross sockets.
    Best effort: returns {"numa_node": n or None, "cpus": count}; never raises."""
    info = {"numa_node": None, "cpus": 0}
    try:
        props = torch.cuda.get_device_properties(local_rank)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info = {"numa_node": node, "cpus": len(allowed)}
    except Exception:
        pass
    return info
