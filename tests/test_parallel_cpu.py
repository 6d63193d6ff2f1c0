"""world_size-2 gloo tests (CPU) of the data-parallel path: the product's allreduce helper turns per-rank shard
gradients into exactly the full-batch gradient of the reference algorithm (oracle), because InstanceNorm is per
sample and every loss is a mean over equal shards."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import model as O
from tests.util_weights import random_d_sd, random_g_sd


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _flat(grads):
    return torch.cat([g.reshape(-1) for g in grads.values()])


def _grads(opt, g_sd, d_sd, vgg, batch):
    gp = {k: v.clone() for k, v in g_sd.items()}
    dp = {k: v.clone() for k, v in d_sd.items()}
    ls, _, gG, gD, _ = O.train_step(opt, gp, dp, vgg, batch, dtype=torch.float64)
    return torch.tensor(ls, dtype=torch.float64), torch.cat([_flat(gG), _flat(gD)])


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from neurips18_hierchical_image_manipulation_b200 import parallel
    assert parallel.world() == (rank, world)
    opt = O.Opt(ngf=4, n_downsample_global=1, n_blocks_global=1, ndf=4, num_D=2, label_nc=5, no_vgg_loss=True)
    g_sd, d_sd = random_g_sd(8, 3, 4, 1, 1), random_d_sd(11, 4, 3, 2)
    g_sd = {k: v.double() for k, v in g_sd.items()}
    d_sd = {k: v.double() for k, v in d_sd.items()}
    full = O.synthetic_batch(4, 32, 32, label_nc=5, seed=21)
    shard = {k: v[rank * 2:(rank + 1) * 2] for k, v in full.items()}
    losses, flat = _grads(opt, g_sd, d_sd, None, shard)
    scale = parallel.allreduce_sum_(flat)
    flat *= scale
    parallel.allreduce_losses_(losses)
    if rank == 0:
        ref_losses, ref_flat = _grads(opt, g_sd, d_sd, None, full)
        out["grad_err"] = float((flat - ref_flat).abs().max() / ref_flat.abs().max())
        out["loss_err"] = float((losses - ref_losses).abs().max())
        out["seeds"] = [parallel.shard_seed(1234, r) for r in range(world)]
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_full_batch():
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out["grad_err"] < 1e-9, dict(out)
    assert out["loss_err"] < 1e-12, dict(out)
    assert out["seeds"] == [1234, 1235]


def test_single_process_is_identity():
    from neurips18_hierchical_image_manipulation_b200 import parallel
    t = torch.arange(8, dtype=torch.float32)
    assert parallel.world() == (0, 1)
    assert parallel.allreduce_sum_(t) == 1.0 and torch.equal(t, torch.arange(8, dtype=torch.float32))
