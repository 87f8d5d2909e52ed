"""Data-parallel sharding of independent work units (audio clips, embedding sets) across ranks — one process per GPU.

Neither hot path communicates inside a unit of work (``AudioMelSpectrogram`` holds only per-instance
scratch, ``OfflineDiarizerManager.cluster(_:)`` touches one ``PreparedDiarization``), so the multi-GPU plan is
"weak scaling, no data-path collective": every rank processes its own units with the single-GPU library and the
only collectives are the barrier / MAX-reduction of timings and an optional gather of small results (labels).
A single AHC problem does not shard (N-1 dependent merges): within one embedding set it is "replicas only".

``torch.distributed`` is used strictly as plumbing (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np


def contiguous_shard(count: int, rank: int, world: int) -> range:
    """Equal-cost units (BASELINE config 4: 512 clips -> 64 per GPU): contiguous blocks, remainder to low ranks."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(count, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def lpt_partition(costs, world: int) -> list[list[int]]:
    """Unequal units (meetings of different size, cost ~ N^2 d): longest-processing-time-first greedy.
    Deterministic: ties broken by unit index, then by rank."""
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    load = [0.0] * world
    out: list[list[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += float(costs[i])
    for lst in out:
        lst.sort()
    return out


def ahc_cost(n: int, dim: int = 256) -> float:
    """Algorithmic bytes of one centroid-linkage problem: 8 d N^2."""
    return 8.0 * dim * float(n) * float(n)


@dataclass
class Dist:
    rank: int
    world: int
    local_rank: int
    backend: str | None

    @property
    def is_root(self) -> bool:
        return self.rank == 0


def init_distributed(backend: str | None = None) -> Dist:
    """Reads RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun) and joins the process group when world > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world <= 1:
        return Dist(0, 1, 0, None)
    import torch
    import torch.distributed as dist
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    if not dist.is_initialized():
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return Dist(rank, world, local, backend)


def bind_to_gpu_numa(local_rank: int) -> str:
    """Pin this process to the CPUs nearest to its GPU (NVML's ideal CPU affinity) BEFORE it allocates pinned host
    buffers, so that first touch places them on the GPU's NUMA node: with one process per GPU sharing two sockets the
    host<->device copies otherwise cross the socket interconnect.  Returns a short description for the bench line;
    never fatal."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        cpus = sorted(os.sched_getaffinity(0))
        return f"nvml cpu affinity: {len(cpus)} cpus ({cpus[0]}..{cpus[-1]})"
    except Exception as e:      # no NVML, a container without the privilege, a CPU-only box
        return f"unbound ({type(e).__name__})"


def numa_node_of(address: int):
    """NUMA node holding the page at `address` (move_pages(2) query through libc; None when it cannot be told): lets a
    bench line state that its pinned buffers are local to the GPU's socket."""
    try:
        import ctypes as C
        libc = C.CDLL(None, use_errno=True)
        page = C.c_void_p(address & ~4095)
        status = C.c_int(-1)
        rc = libc.syscall(279, 0, 1, C.byref(page), None, C.byref(status), 0)   # __NR_move_pages on x86-64
        return int(status.value) if rc == 0 and status.value >= 0 else None
    except Exception:
        return None


def _tensor(values, d: Dist, dtype):
    import torch
    dev = torch.device("cuda", d.local_rank) if d.backend == "nccl" else torch.device("cpu")
    return torch.tensor(values, dtype=dtype, device=dev)


def barrier(d: Dist) -> None:
    if d.world > 1:
        import torch.distributed as dist
        if d.backend == "nccl":
            dist.barrier(device_ids=[d.local_rank])
        else:
            dist.barrier()


def all_reduce_max(d: Dist, value: float) -> float:
    if d.world == 1:
        return float(value)
    import torch
    import torch.distributed as dist
    t = _tensor([float(value)], d, torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_reduce_sum(d: Dist, value: float) -> float:
    if d.world == 1:
        return float(value)
    import torch
    import torch.distributed as dist
    t = _tensor([float(value)], d, torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_labels(d: Dist, local: np.ndarray, counts: list[int]) -> np.ndarray | None:
    """Gathers per-rank int32 label vectors on rank 0 (C5: 64 x 5 000 labels = 1.28 MB in total).
    ``counts[r]`` is the number of labels rank r contributes.  Returns the concatenation on rank 0, None elsewhere."""
    local = np.ascontiguousarray(local, np.int32)
    if d.world == 1:
        return local
    import torch
    import torch.distributed as dist
    width = max(counts) if counts else 0
    dev = torch.device("cuda", d.local_rank) if d.backend == "nccl" else torch.device("cpu")
    mine = torch.full((width,), -1, dtype=torch.int32, device=dev)
    if local.size:
        mine[: local.size] = torch.from_numpy(local).to(dev)
    bucket = [torch.empty_like(mine) for _ in range(d.world)]
    dist.all_gather(bucket, mine)
    if not d.is_root:
        return None
    return np.concatenate([bucket[r][: counts[r]].cpu().numpy() for r in range(d.world)]) if width else local


def gather_bytes(d: Dist, local: np.ndarray, counts: list[int]) -> np.ndarray | None:
    """Gathers per-rank uint8 rows (e.g. one 32-byte SHA-256 digest per clip) on rank 0, in rank order.
    ``local`` is [counts[rank] x width]; returns [sum(counts) x width] on rank 0, None elsewhere."""
    local = np.ascontiguousarray(local, np.uint8)
    if d.world == 1:
        return local
    import torch
    import torch.distributed as dist
    width = local.shape[1] if local.ndim == 2 else 0
    rows = max(counts) if counts else 0
    dev = torch.device("cuda", d.local_rank) if d.backend == "nccl" else torch.device("cpu")
    mine = torch.zeros((rows, width), dtype=torch.uint8, device=dev)
    if local.size:
        mine[: local.shape[0]] = torch.from_numpy(local).to(dev)
    bucket = [torch.empty_like(mine) for _ in range(d.world)]
    dist.all_gather(bucket, mine)
    if not d.is_root:
        return None
    return np.concatenate([bucket[r][: counts[r]].cpu().numpy() for r in range(d.world)])


def finalize(d: Dist) -> None:
    if d.world > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
