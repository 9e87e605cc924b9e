"""CPU (gloo, world_size 2): the host-side logic of the N>1 path -- batch sharding without overlap or loss,
and the flat-bucket gradient all-reduce giving the single-process full-batch gradient."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tps_pp_b200 import parallel as PAR


def test_shard_bounds_cover_exactly():
    for n in (0, 1, 7, 8, 255, 256, 8192):
        for world in (1, 2, 3, 4, 8):
            spans = [PAR.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        PAR.shard_bounds(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
        data = torch.randn(10, 6)
        target = torch.randn(10, 3)
        xs, ts = PAR.shard_batch([data, target], rank, world)
        bucket = PAR.GradBucket(model.parameters())
        bucket.zero()
        # per-rank mean loss scaled so that the rank average equals the full-batch mean
        loss = ((model(xs) - ts) ** 2).sum() / data.shape[0] * world
        loss.backward()
        work = bucket.all_reduce_mean(async_op=True)
        work.wait()
        ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
        ref.load_state_dict(model.state_dict())
        (((ref(data) - target) ** 2).sum() / data.shape[0]).backward()
        err = max(float((p.grad - r.grad).abs().max()) for p, r in zip(model.parameters(), ref.parameters()))
        views_ok = all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in model.parameters())
        mean_loss, = PAR.all_reduce_scalars([float(loss) / world], "cpu")
        # second step after optimizer.zero_grad() (set_to_none=True drops the views): the bucket must pull the fresh
        # gradients in and re-attach, so the all-reduce still averages THIS step's gradients
        opt = torch.optim.SGD(model.parameters(), lr=0.0)
        opt.zero_grad()
        assert all(p.grad is None for p in model.parameters())
        loss = ((model(xs) - ts) ** 2).sum() / data.shape[0] * world
        loss.backward()
        bucket.all_reduce_mean()
        err2 = max(float((p.grad - r.grad).abs().max()) for p, r in zip(model.parameters(), ref.parameters()))
        base = bucket.flat.data_ptr()
        views2 = all(p.grad.data_ptr() == base + 4 * off for p, off in zip(bucket.params, bucket.offsets))
        model.zero_grad(set_to_none=False)          # in-place zeroing keeps the views
        views2 = views2 and all(p.grad.data_ptr() == base + 4 * off for p, off in zip(bucket.params, bucket.offsets))
        err = max(err, err2)
        views_ok = views_ok and views2 and bucket.reattached == len(bucket.params)
        q.put((rank, err, views_ok, xs.shape[0], mean_loss))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_sharded_step_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[3] for r in res] == [5, 5]
    for _, err, views_ok, _, _ in res:
        assert err < 1e-6 and views_ok
    assert abs(res[0][4] - res[1][4]) < 1e-12
