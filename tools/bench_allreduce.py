"""All-reduce latency at the gradient-bucket sizes of the path: NCCL vs torch symmetric-memory kernels (torchrun)."""
import os, sys
import torch, torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm_mem

gname = dist.group.WORLD.group_name
try:
    symm_mem.enable_symm_mem_for_group(gname)
except Exception as e:
    if rank == 0: print("enable_symm_mem_for_group:", e)

def timeit(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item() * 1e3

for mb in ([float(a) for a in sys.argv[1:]] or [4, 15.4, 32]):
    n = max(int(mb * 1e6 / 4) // 1024 * 1024, 1024)
    x = torch.randn(n, device=dev)
    res = {"nccl": timeit(lambda: dist.all_reduce(x))}
    try:
        s = symm_mem.empty(n, dtype=torch.float32, device=dev)
        symm_mem.rendezvous(s, gname)
        s.copy_(x)
        for name in ("one_shot_all_reduce", "two_shot_all_reduce_", "multimem_all_reduce_", "multimem_one_shot_all_reduce"):
            op = getattr(torch.ops.symm_mem, name, None)
            if op is None:
                continue
            try:
                res[name] = timeit(lambda: op(s, "sum", gname))
            except Exception as e:
                res[name] = f"ERR {type(e).__name__}: {str(e)[:80]}"
    except Exception as e:
        res["symm_mem"] = f"ERR {type(e).__name__}: {str(e)[:120]}"
    if rank == 0:
        print(f"world {world} {mb} MB:", {k: (round(v, 1) if isinstance(v, float) else v) for k, v in res.items()}, "us", flush=True)
torch.cuda.synchronize()
dist.barrier()
dist.destroy_process_group()
