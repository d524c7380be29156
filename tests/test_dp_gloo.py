"""Host logic of the batch-sharded data-parallel path (immtsf/dp.py), exercised on CPU with the
gloo backend at world_size 2.  The model on each rank is the CPU oracle (test
infrastructure) -- what is under test is the sharding, the exact-loss normaliser and the single
SUM all-reduce of the gradient bucket: loss and every gradient of the 2-rank run must equal the
single-process run on the same global batch."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "imm-tsf_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem(ttf, mmf):
    from oracle import immtsf_oracle as O
    import gpu_common as G

    d_model, d, C, H, B, N, T = 24, 16, 3, 2, 7, 5, 6  # B odd: uneven shards
    shapes = O.param_shapes(ttf, mmf, d_model, d, C)
    g = torch.Generator().manual_seed(11)
    P = {k: (torch.randn(s, generator=g, dtype=torch.float64) * 0.2) for k, s in shapes.items()}
    notes, tau, t_hat, Y, _ = G.synth_batch(B, N, T, d_model, C, 5)
    truth = torch.randn(B, T, C, generator=g, dtype=torch.float64)
    mask = (torch.rand(B, T, C, generator=g) > 0.4).double()
    mask[:, :, 2] = 0  # a variable without any observation: excluded from the mean (lib/evaluation.py:51-62)
    return O, P, (notes.double(), tau.double(), t_hat.double(), Y.double(), truth, mask), dict(H=H, kappa=0.5)


def _loss_and_grads(O, P, batch, hp, ttf, mmf, group_on):
    from immtsf import dp

    P = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    notes, tau, t_hat, Y, truth, mask = batch
    Yr = Y.clone().requires_grad_(True)
    out = O.fusion_forward(P, ttf, mmf, notes, tau, t_hat, Yr, n_heads=hp["H"], kappa=hp["kappa"], p=0.0, masks=None,
                           faithful_expand=False)
    loss = dp.masked_mse_exact(out, truth, mask)
    loss.backward()
    params = [torch.nn.Parameter(v.detach()) for v in P.values()]
    for q, v in zip(params, P.values()):
        q.grad = v.grad.clone() if v.grad is not None else None
    if group_on:
        dp.allreduce_grads(params)
        tl = loss.detach().clone()
        dist.all_reduce(tl)
        loss = tl
    return loss.detach(), {k: q.grad for k, q in zip(P.keys(), params)}, Yr.grad


def _worker(rank, world, port, ttf, mmf, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from immtsf import dp

        torch.set_num_threads(1)
        O, P, batch, hp = _problem(ttf, mmf)
        local = dp.shard_batch(list(batch), rank, world)
        loss, grads, dY = _loss_and_grads(O, P, local, hp, ttf, mmf, True)
        lo, hi = dp.shard_bounds(batch[0].shape[0], rank, world)
        ret[rank] = (loss, grads, dY, (lo, hi))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ttf,mmf", [("TTF_RecAvg", "MMF_GR_Add"), ("TTF_T2V_XAttn", "MMF_XAttn_Add")])
def test_two_rank_gloo_equals_single_process(ttf, mmf):
    world = 2
    O, P, batch, hp = _problem(ttf, mmf)
    loss1, grads1, dY1 = _loss_and_grads(O, P, batch, hp, ttf, mmf, False)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), ttf, mmf, ret), nprocs=world, join=True)
        res = {r: ret[r] for r in range(world)}
    for r in range(world):
        loss, grads, dY, (lo, hi) = res[r]
        assert torch.allclose(loss, loss1, rtol=1e-12, atol=1e-14), (r, loss, loss1)
        for k in grads1:
            assert grads[k] is not None
            torch.testing.assert_close(grads[k], grads1[k] if grads1[k] is not None else torch.zeros_like(grads[k]),
                                       rtol=1e-10, atol=1e-13, msg=lambda m, k=k: f"{k}: {m}")
        torch.testing.assert_close(dY, dY1[lo:hi], rtol=1e-10, atol=1e-13)  # dY_ts stays local to its shard
    assert res[0][3] == (0, 4) and res[1][3] == (4, 7)


def test_shard_bounds_cover_batch_exactly():
    from immtsf import dp

    for n in (0, 1, 7, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [dp.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_batch_passes_shared_1d_t_hat_through():
    from immtsf import dp

    notes, tau, t_hat, Y = torch.zeros(6, 3, 4), torch.zeros(6, 3), torch.arange(5.0), torch.zeros(6, 5, 2)
    out = dp.shard_batch([notes, tau, t_hat, Y], 1, 2)
    assert out[0].shape[0] == 3 and out[1].shape[0] == 3 and out[3].shape[0] == 3
    assert out[2] is t_hat
