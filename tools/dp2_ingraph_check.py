"""Experimental (run under torchrun on 2 GPUs; NOT part of the test suite -- it hung at teardown on the first try): the data-parallel step captured in ONE CUDA graph with
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
    for ttf, mmf in (("TTF_T2V_XAttn", "MMF_XAttn_Add"), ("TTF_RecAvg", "MMF_GR_Add")):
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
        assert step.group is not None and 0 < step.n_first < step.flat_grads.numel()
        try:
            for _ in range(2):  # replays re-zero the flat bucket and reduce again
                step(*mine)
            got = {k: p.grad.clone() for k, p in fm.named_parameters()}
            # reference: the same modules, eager, over the WHOLE batch in this process
            ref = G.gpu_run(fm, notes, tau, t_hat, Y, Gw, train=True)
            for k, g in ref["grads"].items():
                worst = max(worst, G.assert_close(f"{ttf}+{mmf} {k}", got[k].cpu(), g, 2e-5, floor=1e-3))
        finally:
            step.close()
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print(f"OK dp2 in-graph all-reduce, worst rel err {worst:.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
