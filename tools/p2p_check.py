"""Multi-GPU diagnosis (run under torchrun, 2+ ranks): sharded sort phase times, fused vs staged exchange; and on
rank 0 a plain peer-copy bandwidth probe."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
import gpusorting_b200 as g
from gpusorting_b200 import sharded
e = int(sys.argv[1]) if len(sys.argv) > 1 else 28
n = 1 << e
if rank == 0:
    print("can_access_peer 0->1:", torch.cuda.can_device_access_peer(0, 1), flush=True)
    a = torch.empty(1 << 28, dtype=torch.int32, device="cuda:0")
    b = torch.empty(1 << 28, dtype=torch.int32, device="cuda:1")
    for _ in range(2):
        torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        t0 = time.perf_counter(); b.copy_(a); torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        dt = time.perf_counter() - t0
    print(f"peer copy 1 GiB: {dt*1e3:.2f} ms = {(1<<30)/dt/1e9:.1f} GB/s", flush=True)
    del a, b
dist.barrier()
src = torch.empty(n, dtype=torch.int32, device="cuda")
g.init_random(src, 0, 10 + rank)
s = sharded.ShardedSorter(n, slack_percent=25)
for fused in (True, False, True):
    s.set_fused(fused)
    for it in range(3):
        res = s.sort_keys(src)
        torch.cuda.synchronize()
        tm = s.last_timing()
    ok = sharded.verify_global_order(res, rank, world)
    if rank == 0:
        print(f"n=2^{e}/rank world={world} fused={fused}: " + ", ".join(f"{k}={v:.3f}" for k, v in tm.items()) + f" ok={ok}", flush=True)
dist.barrier()
dist.destroy_process_group()
