"""Multi-GPU sharding of the Poisson likelihood (one process per GPU, torch.distributed).

The path shards without halo exchange: every (sample s, module t, energy e) response plane is
independent in the forward direction.  Two partitions are provided:

  * by SAMPLE  (VI sample axis; BASELINE config 5): rank r owns samples r::world.  Energies are per sample, the
    sky gradient of a sample is complete on its owner.  Collectives: one all-reduce of the (S,) energy vector
    (owners fill their entries, others contribute zeros) and -- only if the caller wants the sample-summed
    gradient (KL gradient w.r.t. the mean) -- one all-reduce of (E,N,N).
  * by MODULE  (telescope modules / observations; BASELINE config 3): rank r owns modules r::world of every
    sample.  Each rank holds only its modules' PSF spectra, exposures and data; partial energies and partial
    sky gradients are summed with one all-reduce each.

The collective is NCCL over NVLink when the tensors are CUDA tensors and gloo in the CPU tests; the compute
object is anything with energy_and_grad(x) -> (energy (S,), grad (S,E,N,N)) -- ResponsePlan in production.

`NcclShardedLikelihood` is the production form of both partitions: the fused evaluation AND the single
ncclAllReduce of one flat buffer are issued by libjubik_b200.so on the same stream
(jubik_poisson_energy_and_grad_sharded); torch.distributed only distributes the ncclUniqueId.
"""
from __future__ import annotations


def owned(n_items, rank, world):
    """Indices owned by `rank` under a round-robin partition of n_items over world ranks."""
    return list(range(rank, n_items, world))


def split_counts(n_items, world):
    return [len(range(r, n_items, world)) for r in range(world)]


class SampleShardedLikelihood:
    """Poisson likelihood over S samples, samples partitioned round-robin over ranks."""

    def __init__(self, plan, n_samples, rank=None, world=None, group=None):
        import torch.distributed as dist
        self.plan = plan
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.S = n_samples
        self.mine = owned(n_samples, self.rank, self.world)

    def local(self, x_all):
        """Select this rank's samples from a replicated (S,E,N,N) batch."""
        return x_all[self.mine]

    def energy_and_grad(self, x_local, reduce_grad=True):
        """x_local: this rank's samples (len(mine),E,N,N).
        Returns (energies of ALL samples (S,), sample-summed gradient (E,N,N) or the local gradients)."""
        import torch
        if x_local.shape[0] != len(self.mine):
            raise ValueError("x_local holds %d samples but this rank owns %d" % (x_local.shape[0], len(self.mine)))
        e_all = torch.zeros((self.S,), dtype=x_local.dtype, device=x_local.device)
        if len(self.mine):
            e_loc, g_loc = self.plan.energy_and_grad(x_local)
            e_all[self.mine] = e_loc
        else:
            # a rank without samples (n_samples < world) contributes zeros but still enters every collective
            g_loc = torch.zeros((0,) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
        self.dist.all_reduce(e_all, group=self.group)
        if not reduce_grad:
            return e_all, g_loc
        g_sum = g_loc.sum(0) if len(self.mine) else torch.zeros(tuple(x_local.shape[1:]), dtype=x_local.dtype,
                                                                 device=x_local.device)
        self.dist.all_reduce(g_sum, group=self.group)
        return e_all, g_sum


class ModuleShardedLikelihood:
    """Poisson likelihood with the telescope-module axis T partitioned round-robin over ranks.
    `plan` must have been built from this rank's modules only (kernel[mine], exposure[mine], data[mine])."""

    def __init__(self, plan, n_modules, rank=None, world=None, group=None):
        import torch.distributed as dist
        self.plan = plan
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.mine = owned(n_modules, self.rank, self.world)

    def energy_and_grad(self, x):
        """x: replicated (S,E,N,N).  Returns the full energy (S,) and full gradient (S,E,N,N) on every rank."""
        import torch
        if len(self.mine):
            e, g = self.plan.energy_and_grad(x)
        else:
            e = torch.zeros((x.shape[0],), dtype=x.dtype, device=x.device)
            g = torch.zeros_like(x)
        self.dist.all_reduce(e, group=self.group)
        self.dist.all_reduce(g, group=self.group)
        return e, g


class NcclShardedLikelihood:
    """Sample- or module-sharded Poisson likelihood with the collective inside the C ABI.

    mode "samples": `plan` is replicated, every rank passes its own (S_local,E,N,N) block of samples (equal S_local
                    on every rank, contiguous blocks: rank r owns samples [r*S_local, (r+1)*S_local)).
                    energy_and_grad -> (energies of all samples (world*S_local,), sample-summed gradient (E,N,N)).
    mode "modules": `plan` holds this rank's modules (kernel[mine], exposure[mine], data[mine]); x is replicated.
                    energy_and_grad -> (energies (S,), gradient (S,E,N,N)), both summed over ranks."""

    def __init__(self, plan, mode="samples", comm=None, group=None):
        from . import _lib
        from .plan import Communicator
        self.plan = plan
        self.mode = {"samples": _lib.SHARD_SAMPLES, "modules": _lib.SHARD_MODULES}[mode]
        self.comm = comm if comm is not None else Communicator.from_torch_distributed(group)
        self.flat = None
        self.group = group
        self._checked = None

    def _check_uniform(self, S):
        """Every rank must pass the same number of samples: the all-reduce count depends on it, and a mismatch would
        hang or corrupt the collective.  Checked once per distinct S through torch.distributed (plumbing only)."""
        if self._checked == S:
            return
        try:
            import torch.distributed as dist
            live = dist.is_available() and dist.is_initialized()
        except Exception:
            live = False
        if live and self.comm.size > 1:
            box = [None] * dist.get_world_size(self.group)
            dist.all_gather_object(box, int(S), group=self.group)
            if len(set(box)) != 1:
                raise ValueError("NcclShardedLikelihood: ranks passed different sample counts %s" % (box,))
        self._checked = S

    def energy_and_grad(self, x, grad_local=None):
        S = x.shape[0] if x.dim() == 4 else 1
        if S < 1:
            raise ValueError("NcclShardedLikelihood needs at least one sample per rank (use SampleShardedLikelihood "
                             "for ragged partitions)")
        self._check_uniform(S)
        n = self.plan.flat_size(self.comm, S, self.mode)
        if self.flat is None or self.flat.numel() != n:
            import torch
            self.flat = torch.empty((n,), dtype=torch.float64, device=self.plan.device)
        e, g = self.plan.energy_and_grad_sharded(x, self.comm, self.mode, flat=self.flat, grad_local=grad_local)
        return e.reshape(-1), g
