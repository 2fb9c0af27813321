"""Multi-process host logic on CPU (gloo, world_size 2): sharding, the single flat all-reduce of the
gradient buffer + loss scalars, and that sharded-sum equals the global-mean gradient of the oracle."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import helpers as H
        from neuromancer_b200 import rollout
        from neuromancer_b200.distributed import GradientSync
        sync = GradientSync()
        assert rollout._WORLD["size"] == world and sync.shard(64) == (rank * 32, rank * 32 + 32)
        # every rank: same weights, its own shard of the global batch; the "kernel" here is the CPU
        # oracle scaled the way the adjoint kernel scales (mean over the GLOBAL batch)
        case = H.make_case("t_vdp_euler_elu", "cpu")
        full = case.make_batch(64, "cpu", 3)
        lo, hi = sync.shard(64)
        shard = {k: (v[lo:hi] if isinstance(v, torch.Tensor) else v) for k, v in full.items()}
        params = [p.detach().clone() for p in case.plan_params]
        local = H.oracle_run("t_vdp_euler_elu", params, shard)
        gbuf = torch.cat([g.reshape(-1) / world for g in local["grads"]] +
                         [torch.stack([local["loss"].detach() / world])])     # [grad | loss] in one message
        sync.allreduce_flat_(gbuf)
        whole = H.oracle_run("t_vdp_euler_elu", params, full)
        want = torch.cat([g.reshape(-1) for g in whole["grads"]] + [torch.stack([whole["loss"].detach()])])
        torch.testing.assert_close(gbuf, want, rtol=1e-5, atol=1e-7)
        # sync_gradients on module parameters (Trainer path)
        ps = [torch.nn.Parameter(torch.zeros(3)), torch.nn.Parameter(torch.zeros(2, 2))]
        for p in ps:
            p.grad = torch.full_like(p, float(rank + 1))
        sc = sync.sync_gradients(ps, scalars=torch.tensor([float(rank)]))
        assert all(torch.equal(p.grad, torch.full_like(p, 3.0)) for p in ps)
        assert torch.allclose(sc, torch.tensor([0.5]))
        sync.close()
        assert rollout._WORLD["size"] == 1
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_world_size_2():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}
