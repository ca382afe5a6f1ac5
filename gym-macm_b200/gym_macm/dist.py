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


class _SharedAlloc(object):
    """A cudaMalloc'ed block owned through the C ABI, visible to torch through __cuda_array_interface__."""

    def __init__(self, nbytes, device):
        import ctypes as C
        from gym_macm import _lib
        ptr, handle = C.c_void_p(), C.create_string_buffer(64)
        _lib.check(_lib.lib().macm_device_alloc(int(device), int(nbytes), C.byref(ptr), handle))
        self.ptr, self.nbytes, self.handle = ptr.value, int(nbytes), handle.raw
        self.__cuda_array_interface__ = {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr, False), "version": 2}

    def __del__(self):
        try:
            from gym_macm import _lib
            _lib.lib().macm_device_free(self.ptr)
        except Exception:
            pass


def alloc_shared(shape, dtype, device):
    """(tensor, meta): a zeroed device tensor in an allocation of its own, and the picklable description the other
    ranks of the box feed to open_peer_tensor() (send it with broadcast_object_list)."""
    n = 1
    for d in shape:
        n *= int(d)
    mem = _SharedAlloc(max(1, n) * torch.empty((), dtype=dtype).element_size(), device)
    t = torch.as_tensor(mem, device=torch.device("cuda", int(device))).view(dtype).view(tuple(shape))
    t._macm_owner = mem   # the allocation lives as long as the tensor
    return t, {"handle": mem.handle, "alloc_offset": 0, "offset": 0, "dtype": dtype, "shape": tuple(shape)}


class PeerBuffer(object):
    """A contiguous array in ANOTHER process's device memory, mapped for this rank's kernels (raw address + shape):
    enough for the C ABI, which only wants pointers.  Slicing the leading axis gives a view."""

    def __init__(self, ptr, shape, dtype, base=None):
        self.ptr, self.shape, self.dtype, self.base = int(ptr), tuple(shape), dtype, base

    def data_ptr(self):
        return self.ptr

    def __getitem__(self, sl):
        lo, hi, step = sl.indices(self.shape[0])
        assert step == 1
        row = torch.empty((), dtype=self.dtype).element_size()
        for d in self.shape[1:]:
            row *= d
        return PeerBuffer(self.ptr + lo * row, (hi - lo,) + self.shape[1:], self.dtype, self.base)


def open_peer_tensor(meta, device=None):
    """Map another rank's device tensor for kernels of `device` (default: the current one): cudaIpcOpenMemHandle on
    that device with lazy peer access, i.e. plain loads/stores over NVLink."""
    import ctypes as C
    from gym_macm import _lib
    device = torch.cuda.current_device() if device is None else int(device)
    key = (device, meta["handle"])
    if key not in _OPENED:     # an allocation is mapped once per process; tensors that share it share the mapping
        base = C.c_void_p()
        _lib.check(_lib.lib().macm_ipc_open(meta["handle"], device, C.byref(base)))
        _OPENED[key] = base.value
    base = _OPENED[key]
    return PeerBuffer(base + meta["alloc_offset"] + meta["offset"], meta["shape"], meta["dtype"], base)


_OPENED = {}


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


class PeerGather(object):
    """The learner-side gather fused into the step: every rank's kernel stores its observations, rewards and done
    flags straight into the LEARNER rank's buffers (NVLink peer stores through CUDA IPC mappings), next to its own
    bound buffers.  No collective, no staging copy: the transfer rides on the step kernel, agent by agent.

    The step runs through macm_rollout with one step per launch, whose per-step output pointers may point
    anywhere -- here at this shard's rows of the learner's [n_total, ...] arrays.  `n_buffers` sets of arrays
    rotate (the learner reads set k while the shards fill set k+1).  Every rank of the group must call the
    constructor; `fence()` makes a filled set visible to the learner (stream sync on every rank + barrier)."""

    NAMES = ("obs", "nn_idx", "rewards", "collided", "done")

    def __init__(self, env, n_total, learner=0, names=("obs", "rewards", "done"), n_buffers=2, group=None):
        self.env, self.group, self.learner = env, group, learner
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.start, self.count = shard_range(n_total, self.rank, self.world)
        eng = env.engine
        assert self.count == eng.E, "the env batch of this rank must be its shard of n_total"
        self.names = tuple(n for n in names if n in eng.t)
        self.full, self.mine = [], []
        for b in range(n_buffers):
            if self.rank == learner:
                made = {n: alloc_shared((n_total,) + tuple(eng.t[n].shape[1:]), eng.t[n].dtype, eng.device.index)
                        for n in self.names}
                full = {n: v[0] for n, v in made.items()}
                meta = [{n: v[1] for n, v in made.items()}]
            else:
                full, meta = None, [None]
            dist.broadcast_object_list(meta, src=learner, group=group)
            if self.rank != learner:
                full = {n: open_peer_tensor(m, eng.device.index) for n, m in meta[0].items()}
            self.full.append(full)
            self.mine.append({n: t[self.start:self.start + self.count] for n, t in full.items()})
        self.k = 0

    def step(self, actions, stream=None):
        """One step of this rank's shard; its outputs land in buffer set `self.k % n_buffers` on the learner."""
        out = self.mine[self.k % len(self.mine)]
        self.env.engine.rollout(actions, 1, None, 0, out, stream)
        self._last_stream = stream if stream is not None else self.env.engine.stream
        self.k += 1
        return out

    _last_stream = None

    def fence(self):
        """The stream the last step() ran on has finished on every rank (the learner may read the filled set)."""
        st = self._last_stream if self._last_stream is not None else torch.cuda.current_stream(self.env.engine.device)
        st.synchronize()
        dist.barrier(group=self.group)

    def close(self):
        """Collective: the shards unmap the learner's arrays (cudaIpcCloseMemHandle), then the learner frees them.
        Call it before building another PeerGather whose arrays could reuse the same device addresses: CUDA refuses
        to map an allocation that is still mapped from an earlier export ("resource already mapped")."""
        import ctypes as C
        from gym_macm import _lib
        torch.cuda.synchronize(self.env.engine.device)
        dist.barrier(group=self.group)
        if self.rank != self.learner:
            bases = set()
            for full in self.full:
                for t in full.values():
                    bases.add(t.base)
            for key, base in list(_OPENED.items()):
                if base in bases:
                    _lib.check(_lib.lib().macm_ipc_close(C.c_void_p(base)))
                    del _OPENED[key]
        dist.barrier(group=self.group)      # every mapping is gone before the learner's tensors are released
        self.full, self.mine = [], []

    def gathered(self, back=1):
        """On the learner: the buffer set filled `back` steps ago (call fence() first)."""
        return self.full[(self.k - back) % len(self.full)]


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
