"""CPU, world_size 2 over gloo: the N>1 host logic of the SDS loop -- independent per-rank views
and ONE all-reduce of the flattened parameter gradients (dwg/parallel.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, 'dreamwaltz-g_b200'))
    from dwg import camera, parallel
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    rng = parallel.rank_rng(1000, rank)
    cam = camera.random_camera(rng, 64, 64)
    params = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(2))]
    params[0].grad = torch.full((5, 3), float(rank + 1))
    params[1].grad = torch.arange(7.0) * (rank + 1)
    if rank == 1:
        params[2].grad = torch.tensor([5.0, 7.0])        # rank 0 has NO gradient for params[2] this step: fixed layout, zeros contributed
    nbytes = parallel.allreduce_grads(params)
    # persistent flat bucket: p.grad are views, autograd accumulates in place, one in-place collective
    w = [torch.nn.Parameter(torch.ones(4, 2) * (rank + 1)), torch.nn.Parameter(torch.ones(3)), torch.nn.Parameter(torch.ones(6))]
    bucket = parallel.GradBucket(w)
    bucket.zero()
    loss = (w[0] ** 2).sum() + (w[1] * float(rank + 2)).sum()          # w[2] unused on every rank -> stays zero
    loss.backward()
    views_ok = all(p.grad.data_ptr() == bucket.flat.data_ptr() + o * 4 for p, o in zip(w, bucket.offsets))
    nb2 = bucket.all_reduce()
    out.put((rank, float(cam['azimuth']), params[0].grad.tolist(), params[1].grad.tolist(), params[2].grad.tolist(), nbytes,
             views_ok, w[0].grad.tolist(), w[1].grad.tolist(), w[2].grad.tolist(), nb2, bucket.offsets))
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_and_independent_views():
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
    assert res[0][1] != res[1][1]                                        # ranks drew different views
    for _, _, g0, g1, g2, nb, views_ok, w0, w1, w2, nb2, offs in res:
        torch.testing.assert_close(torch.tensor(g0), torch.full((5, 3), 3.0))          # 1 + 2
        torch.testing.assert_close(torch.tensor(g1), torch.arange(7.0) * 3.0)
        assert g2 == [5.0, 7.0] and nb == (15 + 7 + 2) * 4                             # same buffer length on both ranks
        assert views_ok and offs == [0, 8, 12] and nb2 == (8 + 4 + 8) * 4              # 16-byte aligned segments
        torch.testing.assert_close(torch.tensor(w0), torch.full((4, 2), 2.0 * 1 + 2.0 * 2))
        torch.testing.assert_close(torch.tensor(w1), torch.full((3,), 2.0 + 3.0))
        assert w2 == [0.0] * 6
