"""N>1 host logic on CPU with gloo, world_size 2: batch sharding and the single flat all-reduce of a training step
(SURVEY.md §8e).  The GPU arithmetic after the collective is covered by tests/test_train_state_gpu.py."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'mix-of-show_b200')


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, PKG)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from mos_b200.dp import FlatTrainState, allreduce_flat, shard_batch
    st = FlatTrainState(32, 768, 294912, 797184)
    assert st.n == 1116672                       # 4.47 MB of fp32, the whole wire payload of a step (+2 scalars)
    g = torch.Generator().manual_seed(100 + rank)
    st.grads[:st.n] = torch.randn(st.n, generator=g)
    scale, loss, norm = allreduce_flat(st, loss_value=1.0 + rank, norm_mean=0.5 * (rank + 1))
    q.put((rank, scale, loss, norm, st.grads[:8].clone(), st.grads[st.n - 3:st.n].clone(), shard_batch(7, rank, world)))
    dist.destroy_process_group()


def test_flat_allreduce_and_sharding_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, l0, n0, h0, t0, b0), (r1, s1, l1, n1, h1, t1, b1) = res
    assert s0 == s1 == 0.5
    assert abs(l0 - 1.5) < 1e-6 and abs(l1 - 1.5) < 1e-6            # mean of the logged losses (util.py:218-221)
    assert abs(n0 - 0.75) < 1e-6
    assert torch.equal(h0, h1) and torch.equal(t0, t1)                # every rank holds the same summed gradient
    ref = sum(torch.randn(1116672, generator=torch.Generator().manual_seed(100 + r)) for r in range(2))
    assert torch.allclose(h0, ref[:8]) and torch.allclose(t0, ref[-3:])
    assert b0 == [0, 1, 2, 3] and b1 == [4, 5, 6]                     # 7 samples over 2 ranks, no overlap, no gap
