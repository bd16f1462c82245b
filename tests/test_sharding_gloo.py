"""world_size-2 gloo test of the multi-GPU host logic (jubik_b200/sharding.py): sample-sharded and
module-sharded likelihoods combine to the single-process result.  The compute object on this CPU box is the
oracle (test infrastructure); on the GPU box the same wrappers hold a ResponsePlan and NCCL does the reduce."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import jubik_oracle as orc
from tests import synth

N, n, m, T, E, S = 32, 4, 4, 2, 1, 4


class OraclePlan:
    def __init__(self, tsel):
        dom = orc.Domain((E, N, N), synth.DIST)
        kern = synth.psf_full(T, n, E, N, sig=(0.8, 2.0))
        expo = synth.exposure(T, E, N)
        full = orc.Response(dom, kern, n, m, expo, 500.0, synth.PIXEL_AREA)
        self.d = synth.counts_from(full.dense(synth.sky(1, E, N, seed=5)[0]))[tsel]
        self.R = orc.Response(dom, kern[tsel], n, m, expo[tsel], 500.0, synth.PIXEL_AREA)

    def energy_and_grad(self, x):
        x = x.numpy()
        out = [orc.energy_and_grad(self.R, self.d, x[s]) for s in range(x.shape[0])]
        e = torch.tensor([o[0] for o in out], dtype=torch.float64)
        g = torch.as_tensor(np.stack([o[1] for o in out])) if out else torch.zeros((0, E, N, N), dtype=torch.float64)
        return e, g


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from jubik_b200.sharding import ModuleShardedLikelihood, SampleShardedLikelihood, owned
    x = torch.as_tensor(synth.sky(S, E, N))
    lh = SampleShardedLikelihood(OraclePlan(list(range(T))), S)
    e_s, g_s = lh.energy_and_grad(lh.local(x))
    mine = owned(T, rank, world)
    lm = ModuleShardedLikelihood(OraclePlan(mine), T)
    e_m, g_m = lm.energy_and_grad(x)
    # fewer samples than ranks: the rank that owns nothing must still enter both collectives
    l1 = SampleShardedLikelihood(OraclePlan(list(range(T))), 1)
    e_1, g_1 = l1.energy_and_grad(l1.local(x[:1]))
    if rank == 0:
        q.put((e_s.numpy(), g_s.numpy(), e_m.numpy(), g_m.numpy(), e_1.numpy(), g_1.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_sample_and_module_sharding_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    e_s, g_s, e_m, g_m, e_1, g_1 = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = OraclePlan(list(range(T)))
    e, g = ref.energy_and_grad(torch.as_tensor(synth.sky(S, E, N)))
    assert np.allclose(e_s, e.numpy(), rtol=1e-13) and np.allclose(g_s, g.numpy().sum(0), rtol=1e-12, atol=1e-12)
    assert np.allclose(e_m, e.numpy(), rtol=1e-13) and np.allclose(g_m, g.numpy(), rtol=1e-12, atol=1e-12)
    assert np.allclose(e_1, e.numpy()[:1], rtol=1e-13) and np.allclose(g_1, g.numpy()[0], rtol=1e-12, atol=1e-12)


def test_partition_helpers():
    from jubik_b200.sharding import owned, split_counts
    assert split_counts(21, 8) == [3, 3, 3, 3, 3, 2, 2, 2]      # C3: 21 planes on 8 GPUs
    assert split_counts(16, 8) == [2] * 8                       # C5: 16 samples on 8 GPUs
    assert sorted(sum((owned(7, r, 3) for r in range(3)), [])) == list(range(7))
