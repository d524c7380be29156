"""torchrun tool: the hand-written in-switch all-reduce (csrc/nvls.cu, immtsf/nvls.py) against NCCL on the same data --
bit-identical results on every rank, error flag clear -- and its latency next to NCCL's at the statistics sizes of the path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "imm-tsf_b200")]
import torch, torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
from immtsf import nvls

comm = nvls.NvlsComm.create(dist.group.WORLD)
if comm is None:
    if rank == 0:
        print("NVLS unavailable on this fabric (no multicast): NCCL carries the all-reduce")
    dist.destroy_process_group()
    sys.exit(0)

def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item() * 1e3

worst = 0.0
for n in (4, 7000, 6933, 597196, 887040 + 768 + 768, 1 << 20):
    comm.reset()
    buf = comm.alloc(n)
    g = torch.Generator(device=dev).manual_seed(1000 * rank + n)
    x = torch.randn(n, device=dev, generator=g)
    ref = x.clone(); dist.all_reduce(ref)
    buf.copy_(x)
    comm.all_reduce(buf)
    torch.cuda.synchronize()
    comm.check()
    err = (buf - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)
    worst = max(worst, err)
    # every rank must hold the SAME bits
    gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf.contiguous())
    same = all(torch.equal(gathered[0], t) for t in gathered)
    us_nvls = timeit(lambda: comm.all_reduce(buf))
    us_nccl = timeit(lambda: dist.all_reduce(ref))
    if rank == 0:
        print(f"world {world} n {n:8d} ({n * 4 / 1e6:6.3f} MB): rel err vs NCCL {err:.2e}, identical on all ranks {same}, "
              f"in-switch {us_nvls:6.1f} us, NCCL {us_nccl:6.1f} us", flush=True)
    assert err < 1e-5 and same
comm.check()
if rank == 0:
    print(f"OK nvls all-reduce, worst rel err vs NCCL {worst:.2e}", flush=True)
torch.cuda.synchronize(); dist.barrier()
dist.destroy_process_group()
