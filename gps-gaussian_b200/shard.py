"""Multi-GPU host logic for the path: independent stereo view-pairs are the sharding unit.

One process per GPU (`torch.distributed`, NCCL on GPUs / gloo in the CPU tests).  Inference needs NO data-path
collective: each rank renders its own view-pairs; the only communication is a barrier and a MAX-reduce of the
elapsed time (so throughput = all units / slowest rank), plus an optional gather of small per-rank results."""
import os
import sys

import torch
import torch.distributed as dist


def world():
    """(rank, world_size, local_rank) from the torchrun environment (1-process default)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend=None, device=None):
    rank, ws, _ = world()
    if ws > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        # NCCL prints its version banner to STDOUT at INFO/VERSION level; bench.py's stdout must be one JSON line
        os.environ["NCCL_DEBUG"] = os.environ.get("GPSG_NCCL_DEBUG", "WARN")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        # NCCL prints its version banner to STDOUT while the communicator is created; bench.py's stdout must carry
        # exactly one JSON line, so fd 1 is pointed at stderr for the duration of the (eager) initialisation.
        sys.stdout.flush()
        saved = os.dup(1)
        try:
            os.dup2(2, 1)
            dist.init_process_group(backend, rank=rank, world_size=ws, **kw)
            if backend == "nccl":
                t = torch.zeros(1, device=device if device is not None else "cuda")
                dist.all_reduce(t)                     # forces communicator creation now
                torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    return rank, ws


def bind_host_to_gpu(local_rank):
    """Pin this process to the CPUs that are NUMA-local to GPU `local_rank` (NVML's ideal affinity) so that the pinned
    host buffers it allocates afterwards, and the copies it drives, stay on the socket the GPU hangs off: with one process
    per GPU, host-link transfers of different ranks then do not cross the inter-socket fabric.  Best effort: returns a
    small dict describing what happened (never raises)."""
    info = {"bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = None
        try:                                                    # CUDA ordinal -> NVML handle through the PCI address
            import torch                                        # (CUDA_VISIBLE_DEVICES can renumber the CUDA side)
            pr = torch.cuda.get_device_properties(int(local_rank))
            bus = "%08X:%02X:%02X.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            info["pci"] = bus
        except Exception:
            h = None
        if h is None:
            h = pynvml.nvmlDeviceGetHandleByIndex(int(local_rank))
        before = len(os.sched_getaffinity(0))
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info.update({"bound": True, "cpus": len(cpus), "cpus_before": before})
    except Exception as exc:                                    # containers may forbid it; the run is still valid
        info = {"bound": False, "why": repr(exc)[:120]}
    return info


def unbind_host():
    """Undo bind_host_to_gpu (e.g. before timing a CPU baseline on all host cores)."""
    try:
        os.sched_setaffinity(0, range(os.cpu_count()))
    except Exception:
        pass


def shard_units(n_units, rank, world_size):
    """Contiguous block partition of unit ids 0..n_units-1; sizes differ by at most one; every unit exactly once."""
    base, rem = divmod(n_units, world_size)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


def unit_seeds(first_seed, units_per_rank, rank):
    """Weak scaling (bench.py): rank r renders units with seeds first_seed + r*units_per_rank + k (reference seed 1314)."""
    return [first_seed + rank * units_per_rank + k for k in range(units_per_rank)]


def barrier(device=None):
    if dist.is_initialized():
        dist.barrier()
    if device is not None and torch.cuda.is_available():
        torch.cuda.synchronize(device)


def max_over_ranks(value, device=None):
    """MAX-reduce of a python float (the timed region's elapsed ms): throughput is set by the slowest rank."""
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_throughput(units_this_rank, elapsed_ms, device=None):
    """views/s of the whole job: (sum of units over ranks) / (max elapsed over ranks)."""
    total = sum_over_ranks(units_this_rank, device)
    ms = max_over_ranks(elapsed_ms, device)
    return total / (ms * 1e-3), total, ms


def finalize():
    if dist.is_initialized():
        dist.destroy_process_group()
