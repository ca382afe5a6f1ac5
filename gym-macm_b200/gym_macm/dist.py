"""Sharding a batch of environments over the GPUs of one box (BASELINE config 5).

Environments never interact (nothing in mvmnt.py / combat.py couples two worlds), so the batch is
cut into contiguous env ranges, one per rank (one process per GPU, torchrun), and `step` needs
no communication at all.  The only exchange is optional and sits outside the simulator: the
learner-side collectives below -- observations / rewards / done flags gathered from every shard
(NCCL all-gather over NVLink when the tensors are CUDA tensors; gloo works for CPU tensors and is
what the CPU tests use) and actions scattered back.
"""
import torch
import torch.distributed as dist


def bind_to_gpu_numa(local_rank):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so that the pinned host buffers of the
    host-stepping path (first touched here) and the DMA engines sit on the same socket.  Returns the CPU list it
    bound to, or None when the box does not say (no NVML, a VM without NUMA topology): then nothing changes."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(local_rank))
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:      # NVML prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None


def shard_range(n_total, rank, world):
    """Contiguous env range [start, start+count) of `rank`; the first n_total % world ranks get one more."""
    base, rem = divmod(int(n_total), int(world))
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def _padded_all_gather(x, counts, group=None):
    """all-gather of per-rank tensors whose first dimension differs (counts[r] rows on rank r)."""
    world = dist.get_world_size(group)
    mx = max(counts)
    if x.shape[0] != mx:
        pad = torch.zeros((mx - x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        x = torch.cat([x, pad], 0)
    out = torch.empty((world * mx,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    if hasattr(dist, "all_gather_into_tensor") and dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    else:
        parts = list(out.view((world, mx) + tuple(x.shape[1:])).unbind(0))
        dist.all_gather(parts, x.contiguous(), group=group)
    if all(c == mx for c in counts):
        return out
    return torch.cat([out[r * mx:r * mx + counts[r]] for r in range(world)], 0)


def all_gather_envs(tensors, n_total, group=None):
    """{name: [E_local, ...]} on every rank -> {name: [n_total, ...]} on every rank, in global env order."""
    world = dist.get_world_size(group)
    counts = [shard_range(n_total, r, world)[1] for r in range(world)]
    return {k: _padded_all_gather(v, counts, group) for k, v in tensors.items()}


def scatter_actions(actions_full, n_total, src=0, group=None, device=None, like=None):
    """Learner rank `src` holds actions for all n_total envs; every rank receives its own shard.
    Implemented as a broadcast + slice (4 bytes per agent; NCCL has no native scatterv)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if rank == src:
        buf = actions_full.contiguous()
    else:
        buf = torch.empty(like.shape if like is not None else actions_full.shape,
                          dtype=(like if like is not None else actions_full).dtype,
                          device=device if device is not None else (like if like is not None else actions_full).device)
    dist.broadcast(buf, src=src, group=group)
    start, count = shard_range(n_total, rank, world)
    return buf[start:start + count]


class ShardedFlock(object):
    """`n_envs_total` Flock environments split over the ranks of the default process group; this
    rank owns envs [start, start+count).  Same surface as BatchedFlock for the local shard, plus
    `gather()` for the learner-side exchange."""

    def __init__(self, n_envs_total, n_agents=[10], targets=None, seed=0, device=None, **kwargs):
        from gym_macm.batched import BatchedFlock
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.n_envs_total = int(n_envs_total)
        self.start, self.count = shard_range(self.n_envs_total, self.rank, self.world)
        # the device sampler is keyed by the GLOBAL env index, so the union of the shards is the
        # same batch whatever the number of ranks
        self.local = BatchedFlock(self.count, n_agents=n_agents, targets=targets, device=device, seed=seed,
                                  env_index_base=self.start, **kwargs)

    def __getattr__(self, name):
        return getattr(self.local, name)

    def step(self, actions):
        return self.local.step(actions)

    def gather(self, names=("obs", "rewards", "done")):
        st = self.local.state
        if self.world == 1:
            return {k: st[k] for k in names}
        return all_gather_envs({k: st[k] for k in names}, self.n_envs_total)
