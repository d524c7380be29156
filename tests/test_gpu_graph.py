"""runtime.GraphedStep: the whole forward+loss+backward step captured in a CUDA graph must reproduce the eager
path on NEW inputs of the same shape (different ragged note counts included), and must draw fresh dropout masks
on every replay."""
import pytest
import torch

import gpu_common as G

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ttf,mmf", [("TTF_RecAvg", "MMF_GR_Add"), ("TTF_T2V_XAttn", "MMF_XAttn_Add"),
                                     ("TTF_T2V_XAttn", "MMF_GR_Add"), ("TTF_RecAvg", "MMF_XAttn_Add"),
                                     ("TTF_T2V_XAttn_old", "MMF_GR_Add"), ("TTF_T2V_XAttn_old", "MMF_XAttn_Add")])
def test_graph_replay_equals_eager_on_new_inputs(ttf, mmf):
    from immtsf import runtime

    cfg = dict(ttf=ttf, mmf=mmf, d_txt=64, C=4, H=2, kappa=0.5)
    fm = G.build_model(cfg, 96, dropout=0.0, seed=1)
    G.randomise_(fm, 2)
    fm.train()
    ex = [t.cuda() for t in G.synth_batch(16, 6, 10, 96, 4, 3)[:4]]
    Gw = torch.randn(16, 10, 4, generator=torch.Generator().manual_seed(5)).cuda()
    loss_fn = lambda out, g: (out * g).sum()
    step = runtime.GraphedStep(fm, example=ex, loss_fn=loss_fn, extras=(Gw,))
    static_grads = [p.grad for p in step.params]
    try:
        for seed in (11, 12):  # new content, new ragged counts, same shapes
            notes, tau, t_hat, Y, _ = G.synth_batch(16, 6, 10, 96, 4, seed)
            loss = step(notes.cuda(), tau.cuda(), t_hat.cuda(), Y.cuda(), Gw)
            got = {"loss": loss.clone(), "Y_out": step.Y_out.clone(), "dY": step.dY_ts.clone(),
                   "grads": {k: p.grad.clone() for k, p in fm.named_parameters()}}
            step.check_nan()
            ref = G.gpu_run(fm, notes, tau, t_hat, Y, Gw.cpu(), train=True)  # eager, fresh grads
            G.assert_close("Y_out", got["Y_out"].cpu(), ref["Y_out"], 1e-6)
            G.assert_close("dY", got["dY"].cpu(), ref["dY"], 1e-5)
            for k, g in ref["grads"].items():
                G.assert_close(k, got["grads"][k].cpu(), g, 1e-5, floor=1e-3)
            # the eager run above replaced p.grad with fresh tensors; put the graph's static buffers back
            for p, g in zip(step.params, static_grads):
                p.grad = g
    finally:
        step.close()


def test_graph_replay_draws_fresh_dropout_masks():
    from immtsf import runtime

    cfg = dict(ttf="TTF_T2V_XAttn", mmf="MMF_XAttn_Add", d_txt=64, C=4, H=1, kappa=0.5)
    fm = G.build_model(cfg, 96, dropout=0.3, seed=1)
    G.randomise_(fm, 2)
    fm.train()
    ex = [t.cuda() for t in G.synth_batch(8, 6, 10, 96, 4, 3)[:4]]
    step = runtime.GraphedStep(fm, example=ex)
    try:
        outs = []
        for _ in range(3):
            step(*ex)
            outs.append(step.Y_out.clone())
        assert torch.isfinite(outs[0]).all()
        assert not torch.equal(outs[0], outs[1]) and not torch.equal(outs[1], outs[2])
        assert int(step.seed_offset.item()) >= 3
    finally:
        step.close()
    # after close() the eager path is unaffected by the offset
    runtime.SEEDS.fixed = 77
    try:
        a = fm(*ex).detach().clone()
        b = fm(*ex).detach().clone()
        assert torch.equal(a, b)
    finally:
        runtime.SEEDS.fixed = None


def test_graph_flat_gradient_bucket():
    """flat_grads=True: every p.grad is a view into one buffer that the graph zeroes and autograd accumulates into;
    the bucket must equal the eager gradients (it is what the data-parallel all-reduce sends)."""
    from immtsf import runtime

    cfg = dict(ttf="TTF_T2V_XAttn", mmf="MMF_GR_Add", d_txt=64, C=4, H=1, kappa=0.5)
    fm = G.build_model(cfg, 96, dropout=0.0, seed=1)
    G.randomise_(fm, 2)
    fm.train()
    notes, tau, t_hat, Y, _ = G.synth_batch(12, 6, 10, 96, 4, 21)
    ex = [t.cuda() for t in (notes, tau, t_hat, Y)]
    Gw = torch.randn(12, 10, 4, generator=torch.Generator().manual_seed(5)).cuda()
    step = runtime.GraphedStep(fm, example=ex, loss_fn=lambda out, g: (out * g).sum(), extras=(Gw,), flat_grads=True)
    try:
        for _ in range(2):  # the second replay must not accumulate on top of the first
            step(*ex, Gw)
        flat = step.flat_grads.clone()
        views = {k: p.grad.clone() for k, p in fm.named_parameters()}
        assert flat.numel() == sum(p.numel() for p in fm.parameters())
        ref = G.gpu_run(fm, notes, tau, t_hat, Y, Gw.cpu(), train=True)
        for k, p in fm.named_parameters():
            G.assert_close(k, views[k].cpu(), ref["grads"][k], 1e-5, floor=1e-3)
        # bucket layout: step.params order (parameters whose gradients are born reduced under in-graph data parallelism
        # come first; without a process group there are none)
        name_of = {id(p): k for k, p in fm.named_parameters()}
        off = 0
        for p in step.params:
            assert torch.equal(flat[off:off + p.numel()].view_as(p), views[name_of[id(p)]])
            off += p.numel()
        assert step.n_first == 0 and step.group is None
    finally:
        step.close()


def test_prefetch_pipeline_equals_direct_calls():
    """GraphedStep.prefetch / step_prefetched (double-buffered H2D on a copy stream) must give the results of calling
    the step with the same batches directly, batch after batch, from pinned host memory."""
    from immtsf import runtime

    cfg = dict(ttf="TTF_T2V_XAttn", mmf="MMF_XAttn_Add", d_txt=64, C=4, H=1, kappa=0.5)
    fm = G.build_model(cfg, 96, dropout=0.0, seed=1)
    G.randomise_(fm, 2)
    fm.train()
    batches = [[t.pin_memory() for t in G.synth_batch(10, 6, 9, 96, 4, s)[:4]] for s in (3, 4, 5)]
    step = runtime.GraphedStep(fm, example=[t.cuda() for t in batches[0]])
    try:
        direct = []
        for b in batches:
            step(*b)
            torch.cuda.synchronize()
            direct.append((step.Y_out.clone(), step.dY_ts.clone(), [p.grad.clone() for p in step.params]))
        step.prefetch(*batches[0])
        for i in range(len(batches)):
            step.step_prefetched()
            if i + 1 < len(batches):
                step.prefetch(*batches[i + 1])  # overlaps the running step
            torch.cuda.synchronize()
            y, dy, gs = direct[i]
            assert torch.equal(step.Y_out, y)
            G.assert_close("dY", step.dY_ts.cpu(), dy.cpu(), 1e-6)
            for p, g in zip(step.params, gs):  # (some reductions over rows use atomics: equal to rounding, not bitwise)
                G.assert_close("grad", p.grad.cpu(), g.cpu(), 1e-5, floor=1e-3)
        assert step.prefetch_done()
    finally:
        step.close()
