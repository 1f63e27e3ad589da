"""Host->device ceiling of the box: every rank streams pinned host buffers of the bench's batch size into its GPU with
cudaMemcpyAsync for ~1 s; aggregate GB/s = all ranks' bytes / the slowest rank's time.  Run under torchrun at the rank
counts of interest (profiles/r02_h2d_ceiling.md):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/h2d_ceiling.py"""
import json, os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voice100_b200.dist import bind_to_gpu_numa, max_over_ranks

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
bound = bind_to_gpu_numa(local) if world > 1 else False
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
out = {"n_gpus": world, "numa_bound": bound, "cpus": len(os.sched_getaffinity(0))}
for name, dtype in (("fp32_245MB", torch.float32), ("int16_123MB", torch.int16)):
    host = [torch.empty((256, 240000), dtype=dtype).pin_memory() for _ in range(2)]
    for h in host:
        h.zero_()
    devb = [torch.empty((256, 240000), dtype=dtype, device=dev) for _ in range(2)]
    stream = torch.cuda.Stream(dev)
    nbytes = host[0].numel() * host[0].element_size()
    reps = max(4, int(1.0 / (nbytes / 20e9)))
    with torch.cuda.stream(stream):
        for i in range(2):
            devb[i].copy_(host[i], non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        for i in range(reps):
            devb[i & 1].copy_(host[i & 1], non_blocking=True)
    torch.cuda.synchronize()
    dt = max_over_ranks(time.perf_counter() - t0, dev)
    out[name] = {"reps": reps, "per_gpu_GBps": round(nbytes * reps / dt / 1e9, 2), "aggregate_GBps": round(world * nbytes * reps / dt / 1e9, 2)}
    del host, devb
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
