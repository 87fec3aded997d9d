"""Data-parallel path with world_size 2 on the gloo backend.

CPU test: the host-side DP logic (`shard_points`, `PointParallel.allreduce_terms/grads`, global
counts) wired around per-rank kernel stand-ins (the oracle's per-rank terms/gradients) must
reproduce the single-rank result on the concatenated batch.
GPU test: the real fused path, two ranks sharing cuda:0 over gloo, against one rank.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import build_problem, rel
from oracle import nsvd_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _init(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)


def _cpu_worker(rank, world, port, xg, seed, q):
    import neural_svd_b200 as N
    _init(rank, world, port)
    cfg = O.PathConfig.oscillator(neigs=4, fourier_mapping_size=16, sequential=True)
    params = {k: v.astype(np.float64) for k, v in O.init_params_like_reference(cfg, seed).items()}
    dp = N.PointParallel()
    x = N.shard_points(torch.from_numpy(xg), rank, world).numpy().astype(np.float64)
    n, b1 = x.shape[0], (x.shape[0] + 1) // 2
    u, acts, sigs = O.forward_streams(x, params, cfg, keep=True)
    Tf, f, aux = O.operator_apply(x, u, params, cfg)
    v, M = O.nesting_masks(cfg.neigs, cfg.sequential, cfg.step)
    G1, G2, ops = O.gram_terms(f, Tf, v.astype(np.float64), b1)
    terms = torch.from_numpy(np.concatenate([G1.ravel(), G2.ravel(), [ops], np.zeros(4)]))
    dp.allreduce_terms(terms, n, b1)                                    # all-reduce #1, counts travel inside it
    t = terms.numpy()
    L = cfg.neigs
    # what loss_finalize_kernel does with the four count floats (nsvd_simt.cu)
    Bg, B1g = int(t[-4] + 65536 * t[-3]), int(t[-2] + 65536 * t[-1])
    B2g = Bg - B1g
    assert (Bg, B1g, B2g) == dp.global_counts(n, b1, "cpu")            # the host-side variant agrees
    loss, lam1, lam2 = O.loss_from_terms(t[:L * L].reshape(L, L), t[L * L:2 * L * L].reshape(L, L), t[2 * L * L], Bg,
                                         B1g, B2g, M.astype(np.float64))
    dF = O.loss_dF(f, Tf, v, M, lam1, lam2, B=Bg, B1=B1g, B2=B2g, b1_local=b1)
    grads = O.mlp_backward(x, dF, params, cfg, u[0], acts, sigs, aux)
    names = sorted(grads)
    flat = torch.from_numpy(np.concatenate([grads[k].ravel() for k in names]))
    dp.allreduce_grads(flat)                                            # all-reduce #2
    if rank == 0:
        q.put((float(loss), flat.numpy(), (Bg, B1g, B2g)))
    dist.destroy_process_group()


def test_two_rank_data_parallel_equals_single_rank_cpu():
    world, seed = 2, 3
    cfg = O.PathConfig.oscillator(neigs=4, fourier_mapping_size=16, sequential=True)
    g = torch.Generator().manual_seed(0)
    xg = (cfg.sampling_scale * torch.randn(48, 2, generator=g)).numpy()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cpu_worker, args=(r, world, port, xg, seed, q)) for r in range(world)]
    [p.start() for p in procs]
    loss, flat, counts = q.get(timeout=120)
    [p.join(60) for p in procs]
    assert counts == (48, 24, 24)
    params = {k: v.astype(np.float64) for k, v in O.init_params_like_reference(cfg, seed).items()}
    r = O.train_step(xg.astype(np.float64), params, cfg)
    want = np.concatenate([r["grads"][k].ravel() for k in sorted(r["grads"])])
    assert abs(loss - r["loss"]) < 1e-12 * abs(r["loss"])
    assert rel(flat, want) < 1e-12


def _gpu_worker(rank, world, port, xg, seed, q):
    import neural_svd_b200 as N
    _init(rank, world, port)
    cfg = O.PathConfig.hydrogen(neigs=4, fourier_mapping_size=64)
    N.set_engine("f16x3")
    method, operator, importance, _ = build_problem(cfg, seed, "cuda:0")
    method.data_parallel = N.PointParallel()
    x = N.shard_points(torch.from_numpy(xg), rank, world).cuda()
    loss, aux = method.compute_loss_operator(operator, x, importance=importance)
    loss.backward()
    if rank == 0:
        q.put((float(loss.detach()), {n: p.grad.cpu().numpy() for n, p in method.named_parameters() if p.grad is not None}))
    dist.destroy_process_group()


@pytest.mark.gpu
def test_two_rank_fused_step_equals_single_rank_gpu():
    import neural_svd_b200 as N
    world, seed = 2, 7
    cfg = O.PathConfig.hydrogen(neigs=4, fourier_mapping_size=64)
    g = torch.Generator().manual_seed(1)
    xg = (cfg.sampling_scale * torch.randn(512, 2, generator=g)).numpy()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gpu_worker, args=(r, world, port, xg, seed, q)) for r in range(world)]
    [p.start() for p in procs]
    loss2, grads2 = q.get(timeout=300)
    [p.join(60) for p in procs]
    N.set_engine("f16x3")
    method, operator, importance, _ = build_problem(cfg, seed, "cuda:0")
    loss1, _ = method.compute_loss_operator(operator, torch.from_numpy(xg).cuda(), importance=importance)
    loss1.backward()
    assert abs(loss2 - float(loss1)) < 1e-5 * abs(float(loss1))
    for n, p in method.named_parameters():
        if p.grad is not None:
            assert rel(grads2[n], p.grad.cpu().numpy()) < 2e-5, n
