"""All-reduce timing of the two gradient buckets (D 29.0 M, G 86.2 M fp32) on N GPUs; prints the NCCL transport."""
import os, torch, torch.distributed as dist
rank, lr = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl")
for n in (29_000_000, 86_200_000):
    x = torch.ones(n, device="cuda")
    for _ in range(3): dist.all_reduce(x)
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): dist.all_reduce(x)
    b.record(); torch.cuda.synchronize()
    if rank == 0:
        ms = a.elapsed_time(b) / 10
        print("all_reduce %d floats: %.3f ms  (%.0f GB/s algorithmic)" % (n, ms, n * 4 / ms / 1e6), flush=True)
dist.destroy_process_group()
