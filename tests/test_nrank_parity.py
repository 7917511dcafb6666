"""SURVEY 8(e), last bullet: N-rank parity.  Two ranks, each with its OWN shard (clips, captions, caption-permutation
stream), identical replicas, SHARED frame offsets (identically seeded CPU generator), gradient penalty scaled by the
world size, gradients exchanged by txt2vid_b200.parallel inside train_iteration.  Checked against the single-process
oracle run on each rank's shard with the same per-rank RNG streams, its D and its G gradients averaged over the ranks
in fp32 before the respective Adam step (oracle train_iteration's `reduce` hook):

  * CPU (gloo, world 2, the kernels' executable spec in fp32 storage): 1e-3, the host logic of the exchange;
  * GPU (-m gpu): the real kernels in the fp32 storage mode; NCCL when the box has two GPUs, otherwise both ranks
    share cuda:0 and exchange over gloo (NCCL refuses two ranks on one device) -- the kernels, the bucket pack /
    unpack and the 1/N folded into the fused Adam are the same either way."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, device_kind, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // world))
    from helpers import golden
    from test_product_vs_oracle_cpu import run_product_iteration
    from txt2vid_b200 import ops, optim, trainer
    from txt2vid_b200.parallel import DistContext
    if device_kind == "cpu":
        import cpu_kernels
        for mod in (ops, optim, trainer):
            mod.K = cpu_kernels
        cpu_kernels.set_store_dtype(torch.float32)
        ops.BF16 = torch.float32
        backend, device, precision = "gloo", "cpu", None
    else:
        two = torch.cuda.device_count() >= world
        backend = "nccl" if two else "gloo"
        os.environ["LOCAL_RANK"] = str(rank if two else 0)
        device, precision = "cuda:%d" % (rank if two else 0), "fp32"
        torch.cuda.set_device(torch.device(device))
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
    ctx = DistContext(backend=backend)
    assert ctx.enabled and ctx.world == world
    fx = golden("tganv2_cond_B8.json")
    orc, got = run_product_iteration(True, fx, device, precision=precision, dist=ctx,
                                     gp_lambda=ctx.gp_lambda_for(0.5, [type("D", (), {"sub_discrims": 1})()]),
                                     data_seed=fx["config"]["data_seed"] + 17 * rank,
                                     np_seed=fx["config"]["seed"] + rank)
    out[rank] = {"orc": {k: orc[k] for k in ("lossD", "lossG", "gradD", "gradG")},
                 "got": {k: got[k] for k in ("lossD", "lossG", "gradD", "gradG")},
                 "bt": [int(b) for b in orc.get("bt_used", [])]}
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def _check(out, world, loss_tol, grad_tol):
    from test_product_vs_oracle_cpu import grad_stats
    # per-rank losses against the per-rank oracle (both carry gp_lambda x world)
    for r in range(world):
        for k in ("lossD", "lossG"):
            a, b = out[r]["got"][k], out[r]["orc"][k]
            assert abs(a - b) <= loss_tol * abs(b), (r, k, a, b)
    assert abs(out[0]["orc"]["lossD"] - out[1]["orc"]["lossD"]) > 1e-6          # the shards really differ
    rep = {}
    for part in ("gradD", "gradG"):
        for r in range(world):
            st = grad_stats(out[r]["got"][part], out[r]["orc"][part])            # the oracle's are rank-averaged
            rep[(part, r)] = st
            assert st["l2"] < grad_tol[part], (part, r, st)
        # every rank holds the same exchanged gradient
        same = grad_stats(out[0]["got"][part], out[1]["got"][part])
        assert same["l2"] < 1e-6, (part, same)
    print("2-rank parity:", {k: round(v["l2"], 6) for k, v in rep.items()})


def _run(device_kind):
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, device_kind, out), nprocs=world, join=True)
    return {r: out[r] for r in range(world)}


def test_two_rank_iteration_matches_the_oracle_on_shards_cpu():
    _check(_run("cpu"), 2, 1e-3, {"gradD": 1e-3, "gradG": 1e-3})


@pytest.mark.gpu
def test_two_rank_iteration_matches_the_oracle_on_shards_gpu():
    """fp32 storage mode through the real kernels; bars as in tests/test_iteration_gpu.py (fp32 floor)"""
    from helpers import fp32_bars
    loss_tol, g_l2, _, _, d_l2, _ = fp32_bars()
    _check(_run("cuda"), 2, loss_tol, {"gradD": d_l2, "gradG": g_l2})
