"""reduce-scatter + all-gather bandwidth of torch.distributed/NCCL on this box (what the direction-sharded Λ-iteration uses)"""
import os, time, torch, torch.distributed as dist
r = int(os.environ["LOCAL_RANK"]); w = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(r)
dist.init_process_group("nccl", device_id=torch.device("cuda", r))
n = (1 << 30) // 8 // w * w      # 1 GiB of doubles
t = torch.ones(n, dtype=torch.float64, device="cuda")
sl = t[r * (n // w):(r + 1) * (n // w)]
for name, fn in (("reduce_scatter", lambda: dist.reduce_scatter_tensor(sl, t)), ("all_gather", lambda: dist.all_gather_into_tensor(t, sl)),
                 ("all_reduce", lambda: dist.all_reduce(t))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    if r == 0:
        print(f"{name}: {n * 8 / 1e9:.2f} GB in {dt * 1e3:.1f} ms  algbw {n * 8 / dt / 1e9:.0f} GB/s", flush=True)
dist.destroy_process_group()
