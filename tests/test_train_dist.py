"""Data-parallel training logic on CPU, world_size 2 over gloo (127.0.0.1): sharding the batch over ranks, all-reducing
the context encoder's BatchNorm statistics (forward and backward) and averaging the gradients must reproduce the
single-process gradient of the whole batch -- the reference's single-device semantics (tf_raft/model.py:126-144)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _grads(rank, world, q_or_none, port):
    """Gradients of the sequence loss on this rank's shard (world > 1: inside a gloo group) or on the whole batch."""
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here); sys.path.insert(0, os.path.dirname(here))
    import cases
    from oracle import weights
    import tf_raft_b200.train as tr
    from tf_raft_b200.losses import sequence_loss
    from test_train import _CpuCorrBlock, _CpuLookup
    tr._Lookup, tr.CorrBlock = _CpuLookup, _CpuCorrBlock
    if world > 1:
        os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
        dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        p = weights.init_params('raft', 21, bias_scale=0.05, norm_jitter=0.1)
        im1, im2 = cases.images(2, 64, 96, 3, 4)
        rng = np.random.default_rng(9)
        flow_gt = torch.from_numpy((rng.standard_normal((2, 64, 96, 2)) * 5).astype(np.float32))
        valid = torch.ones((2, 64, 96), dtype=torch.bool)
        lo, hi = (rank, rank + 1) if world > 1 else (0, 2)
        frozen = ('moving_mean', 'moving_variance')
        P = {k: torch.tensor(v, dtype=torch.float32, requires_grad=not k.endswith(frozen)) for k, v in p.items()}
        names = [k for k, v in P.items() if v.requires_grad]
        moving = {k: v.detach().clone() for k, v in P.items() if k.endswith(frozen)}
        graph = tr.TrainGraph(P, 'raft', 'fp32', moving)
        preds = graph.forward(torch.from_numpy(im1[lo:hi]), torch.from_numpy(im2[lo:hi]), 2)
        loss = sequence_loss([flow_gt[lo:hi], valid[lo:hi]], preds)
        g = torch.autograd.grad(loss, [P[k] for k in names])
        flat = torch.cat([x.flatten() for x in g])
        if world > 1:
            dist.all_reduce(flat)                     # the ONE flat gradient all-reduce of Trainer.step
            flat /= world
        mov = torch.cat([moving[k].flatten() for k in sorted(moving)])
        if q_or_none is not None:
            q_or_none.put((rank, flat.numpy(), mov.numpy()))
        return flat.numpy(), mov.numpy()
    finally:
        if world > 1:
            dist.destroy_process_group()


def test_sharded_training_gradients_equal_the_whole_batch_world2_gloo():
    want, want_mov = _grads(0, 1, None, 0)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grads, args=(r, 2, q, port)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    scale = float(np.abs(want).max())
    for rank, got, mov in res:
        assert float(np.abs(got - want).max()) <= 2e-4 * scale + 2e-6, f'rank {rank}'
        np.testing.assert_allclose(mov, want_mov, atol=1e-6)      # moving statistics follow the GLOBAL batch statistics
