"""How much host<->device bandwidth do N ranks get AT THE SAME TIME on this box?

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_concurrency.py

Every rank copies 512 MB between pinned host memory and its GPU (H2D, D2H, then both directions at
once on two streams), all ranks in lockstep.  The e2e leg of bench.py moves ~1 GB per rank and step over
PCIe; this is its ceiling, measured without any of our kernels in the loop."""
import json, os, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nbytes = 512 << 20
h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
d_a = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
d_b = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(kind, reps=8):
    for timed in (False, True):
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps if timed else 2):
            if kind in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_a.copy_(h_in, non_blocking=True)
            if kind in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_b, non_blocking=True)
        barrier()
        dt = time.perf_counter() - t0
    per_dir = nbytes * reps / dt / 1e9
    t = torch.tensor([per_dir], device="cuda", dtype=torch.float64)
    lo, hi = t.clone(), t.clone()
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return {"min_gbs_per_rank_and_direction": float(lo.item()), "max": float(hi.item())}


res = {k: run(k) for k in ("h2d", "d2h", "both")}
if rank == 0:
    print(json.dumps({"ranks": world, "bytes_per_copy": nbytes, "pinned_copy_bandwidth": res}))
if world > 1:
    dist.destroy_process_group()
