"""CPU, world_size 2 over gloo: the data-parallel gradient exchange of txt2vid_b200.parallel (bucket pack ->
one all-reduce -> unpack, channels-last parameter layouts included) equals the mean of the per-rank gradients,
and identically seeded ranks draw identical frame offsets while permutations / z differ."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from txt2vid_b200.parallel import DistContext
    from txt2vid_b200 import hostrng
    ctx = DistContext(backend="gloo")
    assert ctx.enabled and ctx.world == world and ctx.gp_scale == float(world)
    torch.manual_seed(1234)                       # per-rank parameters would be identical in real use
    conv = torch.nn.Conv3d(4, 6, 3)
    lin = torch.nn.Linear(5, 3)
    # one weight in channels-last memory, like the product's re-homed conv weights
    conv.weight.data = conv.weight.data.contiguous(memory_format=torch.channels_last_3d)
    # a ConvLSTM-style 3x3 kernel on a 1x1 plane: only its centre tap is exchanged
    clstm = torch.nn.Conv2d(4, 6, 3, 1, 1, bias=False)
    clstm.weight._t2v_live_tap = (1, 1)
    params = list(conv.parameters()) + list(lin.parameters()) + list(clstm.parameters())
    g = torch.Generator().manual_seed(100 + rank)
    for p in params:
        p.grad = torch.empty_like(p, memory_format=torch.preserve_format)
        p.grad.copy_(torch.randn(p.shape, generator=g))
    local = [p.grad.clone() for p in params]
    opt = torch.optim.SGD(params, lr=0.1)
    ctx.reduce_grads(opt)
    # frame offsets: shared CPU seed -> same on all ranks; permutation: per-rank numpy seed
    torch.manual_seed(100)
    np.random.seed(100 + rank)
    bts = [hostrng.EagerDraws().bt(2) for _ in range(7)]
    perm = hostrng.EagerDraws().perm(8, "cpu").tolist()
    assert ctx.last_bucket_bytes == 4 * (sum(p.numel() for p in params[:-1]) + 6 * 4)
    # ---- the helpers of the multi-rank CLI path (txt2vid/train/gan.py under torchrun)
    # replicas that were initialised differently start from rank 0's weights and buffers
    torch.manual_seed(500 + rank)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.BatchNorm2d(4))
    net[1].running_mean.add_(float(rank + 1))
    ctx.broadcast_module(net)
    ctx.broadcast_module(None)
    state = [t.detach().clone() for t in list(net.parameters()) + list(net.buffers())]
    # per-rank random streams with a SHARED torch CPU generator (frame offsets), after seed_ranks
    ctx.seed_ranks(77)
    import random
    shared, own_np, own_py = float(torch.rand(())), float(np.random.rand()), random.random()
    # every rank iterates its own stride of the batches, equal counts (7 batches, world 2 -> 3 each)
    mine = list(ctx.shard(list(range(7))))
    assert len(ctx.shard(list(range(7)))) == 3 and ctx.is_main == (rank == 0)
    # summed (multi-scale) penalties are rescaled for gradient averaging, batch-mean penalties are not
    ms, plain = type("MS", (), {"sub_discrims": [1]})(), object()
    lam = (ctx.gp_lambda_for(0.5, [ms]), ctx.gp_lambda_for(0.5, [plain]), ctx.gp_lambda_for(0.5, [ms, plain]),
           ctx.gp_lambda_for(-1.0, [ms]))
    assert lam == (0.5 * world, 0.5, 0.5, -1.0), lam
    out[rank] = {"local": local, "reduced": [p.grad.clone() for p in params], "bts": bts, "perm": perm,
                 "strides": [p.grad.stride() for p in params], "state": state, "shared": shared, "own": (own_np, own_py),
                 "mine": mine}
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_reduce_grads_world2():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    n = len(r0["local"])
    for i, (a, b, m0, m1) in enumerate(zip(r0["local"], r1["local"], r0["reduced"], r1["reduced"])):
        mean = (a + b) / 2
        if i < n - 1:
            assert torch.allclose(m0, mean, atol=1e-6) and torch.allclose(m1, mean, atol=1e-6)
        else:                                    # tagged kernel: centre tap averaged, the other taps left as they were
            assert torch.allclose(m0[:, :, 1, 1], mean[:, :, 1, 1], atol=1e-6)
            assert torch.allclose(m1[:, :, 1, 1], mean[:, :, 1, 1], atol=1e-6)
            keep = torch.ones(3, 3, dtype=torch.bool)
            keep[1, 1] = False
            assert torch.equal(m0[:, :, keep], a[:, :, keep]) and torch.equal(m1[:, :, keep], b[:, :, keep])
    assert r0["bts"] == r1["bts"]
    assert r0["perm"] != r1["perm"]
    assert all(torch.equal(a, b) for a, b in zip(r0["state"], r1["state"]))           # broadcast_module
    assert float(r1["state"][-3].mean()) == 1.0                                        # ... rank 0's running_mean (0 + 1), not rank 1's (2)
    assert r0["shared"] == r1["shared"] and r0["own"][0] != r1["own"][0] and r0["own"][1] != r1["own"][1]
    assert r0["mine"] == [0, 2, 4] and r1["mine"] == [1, 3, 5]                          # shard: strided, truncated
    assert r0["strides"][0] == r1["strides"][0] and r0["strides"][0][1] == 1      # layout preserved (channels-last)
