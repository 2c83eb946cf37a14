"""Data-parallel runtime: one process per GPU, NCCL over NVLink.

Replaces the reference's single-process nn.DataParallel (main.py:730-734), which re-broadcasts
all parameters and gathers features to GPU 0 every step. Here parameters are never
re-broadcast after initialisation; per modality turn there are exactly two collectives:
  1. all-reduce(avg) of the ACTIVE encoder's gradients, as one flat bucket (views of one
     contiguous buffer: no pack/unpack copies), and
  2. one small all-reduce(sum) of the packed head buffer [dW | db | sum_b feat] so that every
     rank forms the same averaged head gradient and the same global-batch r; the GS kernel is
     deterministic, hence P stays bit-identical on all ranks without a broadcast.
With world_size == 1 (or torch.distributed not initialised) everything is a no-op.
The same code runs on CPU with the gloo backend (tests, world_size 2).
"""
import os

import torch
import torch.distributed as dist


def is_dist():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def world_size():
    return dist.get_world_size() if is_dist() else 1


def rank():
    return dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0


def bind_host_to_gpu(device_index=None):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (sysfs `local_cpulist` of the GPU's PCI function), so
    that the pinned host batches it allocates afterwards live in that node's memory: with one process per GPU on a
    two-socket box, unpinned processes put most input buffers on the wrong socket and the per-step H2D copies of all ranks
    (8 x 89 MB for configs[1]) funnel through the inter-socket link. Returns the CPU set, or None when the topology is not
    exposed (containers without sysfs NUMA data, single-node hosts): then nothing changes. MLA_NUMA_BIND=0 disables it."""
    if os.environ.get("MLA_NUMA_BIND", "1") == "0" or not torch.cuda.is_available() or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        idx = torch.cuda.current_device() if device_index is None else device_index
        pr = torch.cuda.get_device_properties(idx)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open("/sys/bus/pci/devices/%s/local_cpulist" % bdf) as fh:
            text = fh.read().strip()
        cpus = set()
        for part in text.split(","):
            if part:
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)                  # never widen what the launcher / cgroup allows
        if not cpus or cpus == os.sched_getaffinity(0):
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment (RANK/WORLD_SIZE/LOCAL_RANK/MASTER_*)."""
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if ws <= 1 or (dist.is_available() and dist.is_initialized()):
        return rank(), world_size()
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        bind_host_to_gpu()
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group(backend=backend)
    return rank(), world_size()


class FlatGrads:
    """One contiguous gradient buffer per parameter group (an encoder); `.views[i]` aliases it.

    Backward kernels write (or autograd accumulates) straight into the views, the all-reduce
    runs on the flat buffer, and the optimizer reads the same memory through p.grad."""

    def __init__(self, params):
        self.params = [p for p in params]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views = []
        off = 0
        for p in self.params:
            # same shape AND strides as the parameter (conv weights are channels_last): kernels write
            # gradients in the parameter's own memory order, optimizers see matching layouts
            self.views.append(torch.as_strided(self.flat, p.shape, p.stride(), storage_offset=off))
            off += p.numel()

    def attach(self, zero=False):
        """Point p.grad at the views. The native encoder backward overwrites every element, so no
        zero fill is needed (zero=True only if something accumulates into the views)."""
        if zero:
            self.flat.zero_()
        for p, v in zip(self.params, self.views):
            p.grad = v

    def detach(self):
        """Hide the gradients from the optimiser (p.grad = None) without touching the flat buffer."""
        for p in self.params:
            p.grad = None

    def allreduce_avg(self):
        if is_dist():
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.mul_(1.0 / world_size())


_aux_groups = {}


def aux_group(tag):
    """A second communicator (own NCCL stream) for collectives that must not queue behind the main group's — the deferred
    encoder's gradient buckets. Created collectively: every rank must ask for the same tags in the same order."""
    if not is_dist():
        return None
    g = _aux_groups.get(tag)
    if g is None:
        g = dist.new_group()
        _aux_groups[tag] = g
    return g


class BucketedAllReduce:
    """Sum-all-reduce of one flat gradient buffer as buckets issued asynchronously while the backward pass is still producing
    the others: `span(lo, hi)` = elements [lo, hi); `tail(off)` = [off, n) (the layers nearest the loss: complete first),
    `head(off)` = [0, off). `wait()` makes the CURRENT stream wait for every bucket issued so far. With world_size 1
    everything is a no-op."""

    def __init__(self, flat, group=None):
        self.flat, self.group, self.works = flat, group, []

    def _issue(self, t):
        if is_dist() and t.numel():
            self.works.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def span(self, lo, hi):
        self._issue(self.flat[lo:hi])

    def tail(self, off):
        self._issue(self.flat[off:])

    def head(self, off):
        self._issue(self.flat[:off])

    def wait(self):
        for w in self.works:
            w.wait()
        self.works = []


def allreduce_sum_(t):
    if is_dist():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def allreduce_sum_async(t):
    """Issue the sum-all-reduce of `t` (it starts once the work queued on the current stream so far is done) and return the
    handle; `.wait()` makes the then-current stream wait for the result. None when not distributed."""
    if is_dist():
        return dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=True)
    return None


def all_gather_rows(t, sizes=None):
    """Concatenate [B_local, ...] tensors from all ranks in rank order. The local batch sizes may differ (the last batch
    of a sharded loader without drop_last): they are exchanged first, shards are padded to the largest one for the
    collective and trimmed afterwards. `sizes` (the list returned by `gather_sizes`) skips the exchange when several
    tensors of the same batch are gathered."""
    if not is_dist():
        return t
    if sizes is None:
        sizes = gather_sizes(t.shape[0], t.device)
    big = max(sizes)
    t = t.contiguous()
    if t.shape[0] < big:
        pad = torch.zeros((big - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        t = torch.cat([t, pad], dim=0)
    out = [torch.empty_like(t) for _ in range(world_size())]
    dist.all_gather(out, t)
    if all(n == big for n in sizes):
        return torch.cat(out, dim=0)
    return torch.cat([o[:n] for o, n in zip(out, sizes)], dim=0)


def gather_sizes(n, device):
    """Local batch sizes of all ranks, in rank order (one tiny all-gather; a host read)."""
    if not is_dist():
        return [int(n)]
    mine = torch.tensor([int(n)], dtype=torch.int64, device=device)
    out = [torch.empty_like(mine) for _ in range(world_size())]
    dist.all_gather(out, mine)
    return [int(o.item()) for o in out]


def broadcast_buffers_(module, src=0):
    """BatchNorm running statistics are per-rank during training (as under nn.DataParallel, where only replica
    0's survive, main.py:732); evaluation must use ONE set on every rank: rank `src`'s."""
    if not is_dist():
        return
    for b in module.buffers():
        dist.broadcast(b, src=src)


def params_checksum(t):
    """Order-independent integer checksum of a tensor's bits (cross-rank bit-identity checks)."""
    return int(t.detach().contiguous().view(torch.int32).to(torch.int64).sum().item())
