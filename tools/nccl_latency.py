"""Development probe (torchrun): latency of a tiny NCCL all-reduce / all-gather on this box, and the transports NCCL chose."""
import os, time, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
x = torch.zeros(1, dtype=torch.int32, device="cuda"); g = torch.zeros(256, dtype=torch.int64, device="cuda"); ga = torch.zeros(256 * world, dtype=torch.int64, device="cuda")
for _ in range(20): dist.all_reduce(x)
torch.cuda.synchronize(); dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(200): dist.all_reduce(x)
b.record(); b.synchronize()
t1 = a.elapsed_time(b) / 200
a.record()
for _ in range(200): dist.all_gather_into_tensor(ga, g)
b.record(); b.synchronize()
t2 = a.elapsed_time(b) / 200
# with a host sync after every collective (the sharded step's pattern)
t0 = time.perf_counter()
for _ in range(100):
    dist.all_reduce(x); torch.cuda.synchronize()
t3 = (time.perf_counter() - t0) / 100 * 1e3
if rank == 0: print(f"world {world}: all_reduce(4 B) {t1*1e3:.1f} us, all_gather(2 KB) {t2*1e3:.1f} us back to back; all_reduce + host sync {t3*1e3:.1f} us", flush=True)
dist.destroy_process_group()
