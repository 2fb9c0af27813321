"""In-kernel all-reduce over NVLink peer memory (csrc/nm_collective.cuh, nm_grad_allreduce in the C ABI) on TWO GPUs:
`gpurun --gpus 2 -- python -m pytest tests/test_peer_allreduce_gpu.py -m gpu`.  Skipped on a one-GPU box (the host-side
logic of the data-parallel path is covered by the gloo world-size-2 tests)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")]


def _worker(rank, world, port, no_multicast, q):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
        if no_multicast:
            os.environ["NM_AR_NO_MULTICAST"] = "1"
        import torch.distributed as dist
        dev = torch.device("cuda", rank)
        torch.cuda.set_device(dev)
        dist.init_process_group("nccl", device_id=dev)
        from neuromancer_b200 import configs, rollout as nr
        from neuromancer_b200.distributed import PeerAllReduce
        n, tail = 53_506 + 9, 53_506
        sync = PeerAllReduce(n, dev, tail_from=tail, tail_scale=1.0 / world)
        out = torch.empty(n, device=dev)
        g = torch.Generator(device="cpu"); g.manual_seed(5)
        for step in range(40):
            data = [torch.randn(n, generator=g) * (10.0 ** ((step % 7) - 3)) for _ in range(world)]     # every rank knows all shards
            sync.half().copy_(data[rank].to(dev))
            if (step + rank) % 3 == 0:
                torch.cuda._sleep(20_000_000)                    # rank skew: ~10 ms on one side
            sync.all_reduce(out)
            want = data[0].clone()
            for q_ in range(1, world):
                want += data[q_]                                  # rank order, fp32
            want[tail:] *= 1.0 / world
            got = out.cpu()
            if no_multicast:
                assert torch.equal(got, want), (step, (got - want).abs().max())
            else:
                torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-6 * float(want.abs().max()))
            gathered = [torch.empty_like(out) for _ in range(world)]
            dist.all_gather(gathered, out)
            assert all(torch.equal(gathered[0], t) for t in gathered)    # every rank holds the same bits
        assert sync.status() == 0
        # the data-parallel train step on the C ABI: gradients through the in-kernel exchange == through NCCL
        case = configs.make_case("c2", dev)
        batch = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in case.make_batch(4096, "cpu", 100 + rank).items()}
        plan = case.system.plan_for(case.system.init(dict(batch)), loss=case.problem.loss)
        a = nr.RawTrainStep(plan, batch, world_size=world, peer_sync=True)
        b = nr.RawTrainStep(plan, batch, world_size=world, peer_sync=False)
        for _ in range(3):
            a.forward(); a.backward(); a.all_reduce()
            b.forward(); b.backward(); b.all_reduce()
            torch.testing.assert_close(a.gbuf, b.gbuf, rtol=2e-6, atol=1e-7 * float(b.gbuf.abs().max()))
        mode = sync.mode
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok", mode))
    except Exception as e:                                            # noqa: BLE001
        import traceback
        q.put((rank, "error", traceback.format_exc()[-3000:]))


@pytest.mark.parametrize("no_multicast", [False, True], ids=["nvls", "peer-loads"])
def test_in_kernel_allreduce_on_two_gpus(no_multicast):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + (1 if no_multicast else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, no_multicast, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = []
    try:
        for _ in procs:
            res.append(q.get(timeout=240))
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.kill()
    assert all(r[1] == "ok" for r in res), res
    print("exchange mode:", res[0][2])
