"""Host-side logic of the multi-GPU path on CPU: world_size 2 over gloo (127.0.0.1)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tf_raft_b200 import parallel


def test_shard_bounds_cover_the_batch():
    for b in (1, 4, 7, 32):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(b, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == b
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1
    with pytest.raises(ValueError):
        parallel.shard_bounds(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, batch, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        im1 = torch.rand(batch, 8, 8, 3, generator=g)
        im2 = torch.rand(batch, 8, 8, 3, generator=g)

        def fake_predict(a, b):          # any per-sample function: stands in for model.predict_step
            return (a - b)[..., :2] * torch.arange(1, 3)

        out = parallel.predict_sharded(fake_predict, im1, im2)
        ok = torch.equal(out, fake_predict(im1, im2))
        t = parallel.max_over_ranks(1.0 + rank, torch.device('cpu'))
        q.put((rank, bool(ok), t))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('batch', [4, 5])
def test_predict_sharded_world2_gloo(batch):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True]
    assert [r[2] for r in res] == [2.0, 2.0]          # max over ranks of (1 + rank)


def test_predict_stream_schedule_on_cpu():
    """Plumbing of the pipelined predict (parallel.predict_stream): order, one result per pair, look-ahead of one."""
    import torch
    from tf_raft_b200 import parallel
    calls = []

    def batches():
        for i in range(5):
            calls.append(('fetch', i))
            yield torch.full((2, 4, 4, 3), float(i)), torch.full((2, 4, 4, 3), float(10 * i))

    def predict(a, b):
        calls.append(('predict', int(a[0, 0, 0, 0])))
        return a[..., :2] + b[..., :2]

    outs = list(parallel.predict_stream(predict, batches(), 'cpu'))
    assert [float(o[0, 0, 0, 0]) for o in outs] == [0.0, 11.0, 22.0, 33.0, 44.0]
    # pair i+1 is fetched (uploaded) before pair i is computed
    assert calls[:4] == [('fetch', 0), ('fetch', 1), ('predict', 0), ('fetch', 2)]
    assert list(parallel.predict_stream(predict, iter(()), 'cpu')) == []
