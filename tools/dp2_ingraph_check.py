"""Two-GPU check (run under torchrun; not part of the 1-GPU test suite): the data-parallel step captured in ONE CUDA graph with
its two bucketed all-reduces (runtime.GraphedStep(allreduce_group=True)) must give every rank the gradients of the
whole batch, i.e. the same numbers as a single-process run over the concatenated shards."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "imm-tsf_b200"), os.path.join(ROOT, "tests")]  # gpu_common lives in tests/
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import gpu_common as G  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    from immtsf import dp, runtime

    worst = 0.0
    for ttf, mmf in (("TTF_T2V_XAttn", "MMF_XAttn_Add"), ("TTF_RecAvg", "MMF_XAttn_Add"), ("TTF_RecAvg", "MMF_GR_Add")):
        cfg = dict(ttf=ttf, mmf=mmf, d_txt=64, C=4, H=2, kappa=0.5)
        fm = G.build_model(cfg, 96, dropout=0.0, seed=1)  # identical parameters on every rank
        G.randomise_(fm, 2)
        fm.train()
        B = 8 * world
        notes, tau, t_hat, Y, Gw = G.synth_batch(B, 6, 10, 96, 4, 21)
        full = [notes, tau, t_hat, Y, Gw]
        mine = [t.cuda() for t in dp.shard_batch(full, rank, world)]
        loss_fn = lambda out, g: (out * g).sum()
        step = runtime.GraphedStep(fm, example=mine[:4], loss_fn=loss_fn, extras=(mine[4],), allreduce_group=True)
        # rank form: the MMF weight gradients (and RecAvg's folded `proj`) are born reduced and not communicated
        assert step.group is not None and (step.n_first > 0) == (mmf == "MMF_XAttn_Add"), step.n_first
        if rank == 0:
            print(f"{ttf}+{mmf}: {step.n_first} of {step.n_total} gradient floats born reduced; {step.dp_calls_per_step} collectives, "
                  f"{step.dp_floats_per_step} floats per step, {step.dp_nvls_calls_per_step} in-switch", flush=True)
        try:
            for _ in range(2):  # replays re-zero the flat bucket and reduce again
                step(*mine)
            step.check_comm()
            got = {k: p.grad.clone() for k, p in fm.named_parameters()}
            # reference: the same modules, eager, over the WHOLE batch in this process
            ref = G.gpu_run(fm, notes, tau, t_hat, Y, Gw, train=True)
            for k, g in ref["grads"].items():
                worst = max(worst, G.assert_close(f"{ttf}+{mmf} {k}", got[k].cpu(), g, 2e-5, floor=1e-3))
        finally:
            step.close()
    # train mode WITH dropout (T2V's final projection is folded into the rank operand, so its gradient is born reduced): the masks are
    # functions of the local sample index, so the reference is the plain schedule on the same shards -- one all-reduce of
    # the whole flat bucket after the replay -- with the same seeds
    cfg = dict(ttf="TTF_T2V_XAttn", mmf="MMF_XAttn_Add", d_txt=64, C=4, H=1, kappa=0.5)
    fm = G.build_model(cfg, 96, dropout=0.2, seed=1)
    G.randomise_(fm, 2)
    fm.train()
    notes, tau, t_hat, Y, Gw = G.synth_batch(8 * world, 6, 10, 96, 4, 33)
    mine = [t.cuda() for t in dp.shard_batch([notes, tau, t_hat, Y, Gw], rank, world)]
    loss_fn = lambda out, g: (out * g).sum()
    runtime.SEEDS.fixed = 0xD15EA5E
    grads = {}
    for mode in ("post", "ingraph"):
        step = runtime.GraphedStep(fm, example=mine[:4], loss_fn=loss_fn, extras=(mine[4],),
                                   **(dict(allreduce_group=True) if mode == "ingraph" else dict(flat_grads=True)))
        step(*mine)
        if mode == "post":
            dp.allreduce_grads(step.params, flat=step.flat_grads)
        else:
            assert step.n_first > 0.4 * step.n_total and step.coalesced, (step.n_first, step.n_total)
        torch.cuda.synchronize()
        grads[mode] = {k: p.grad.clone() for k, p in fm.named_parameters()}
        step.close()
        step = None
    runtime.SEEDS.fixed = None
    for k, g in grads["post"].items():
        worst = max(worst, G.assert_close(f"dropout {k}", grads["ingraph"][k].cpu(), g.cpu(), 2e-5, floor=1e-3))
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print(f"OK dp2 in-graph all-reduce, worst rel err {worst:.2e}", flush=True)
    step = None  # graphs that captured NCCL work go before the communicator
    import gc

    gc.collect()
    torch.cuda.synchronize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
