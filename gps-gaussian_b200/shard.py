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
    """Process group for this rank.  A caller-set NCCL_DEBUG (e.g. the driver's NCCL_DEBUG=INFO to count communicator
    ranks) is honoured and NCCL's own log lines go wherever NCCL sends them (stdout by default, or NCCL_DEBUG_FILE); the
    bench's JSON line is a separate line of stdout.  Nothing is redirected or overridden here."""
    rank, ws, _ = world()
    if ws > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, rank=rank, world_size=ws, **kw)
        if backend == "nccl":
            t = torch.zeros(1, device=device if device is not None else "cuda")
            dist.all_reduce(t)                         # forces communicator creation now, outside any timed region
            torch.cuda.synchronize()
            sys.stdout.flush()
    return rank, ws


def allreduce_grads(params, bucket=None):
    """Data-parallel gradient averaging for stage-2 training (BASELINE config C5): ONE flat fp32 bucket, SUM all-reduce,
    divide by the world size, written back in place -- issued between `scaler.scale(loss).backward()` and
    `scaler.unscale_()` exactly where /root/reference/train_stage2.py:83-85 needs it (clip_grad_norm_ then sees the global
    gradient and GradScaler's found-inf is identical on every rank).  5 144 408 fp32 values = 20.6 MB for the reference's
    stage-2 model.  Parameters without a gradient (the reference's unused 1/16 and 1/32 GRUs) are skipped on every rank
    alike.  `bucket`: optional reusable flat tensor.  Returns the flat bucket (averaged)."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return None
    n = sum(g.numel() for g in grads)
    if bucket is None or bucket.numel() != n or bucket.device != grads[0].device:
        bucket = torch.empty(n, dtype=torch.float32, device=grads[0].device)
    off = 0
    views = []
    for g in grads:
        v = bucket[off:off + g.numel()].view_as(g)
        views.append(v)
        off += g.numel()
    torch._foreach_copy_(views, grads)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(bucket, op=dist.ReduceOp.SUM)
        bucket.div_(dist.get_world_size())
    torch._foreach_copy_(grads, views)
    return bucket


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
