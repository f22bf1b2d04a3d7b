"""Run a few sorts (for ncu captures). usage: one_sort.py log2n [pairs|u64] [variant] [reps]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpusorting_b200 as g
e = int(sys.argv[1]) if len(sys.argv) > 1 else 28
kind = sys.argv[2] if len(sys.argv) > 2 else "keys"
variant = int(sys.argv[3]) if len(sys.argv) > 3 else -1
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
n = 1 << e
if kind == "u64":
    w = torch.empty(2 * n, dtype=torch.int32, device="cuda"); g.init_random(w, 0, 10); t = w.view(torch.int64)
    s = g.OneSweepSorter(n, 8, 0)
else:
    t = torch.empty(n, dtype=torch.int32, device="cuda"); g.init_random(t, 0, 10)
    s = g.OneSweepSorter(n, 4, 4 if kind == "pairs" else 0)
if variant >= 0:
    s.set_option("variant", variant)
src = t.clone()
v = torch.arange(n, dtype=torch.int32, device="cuda") if kind == "pairs" else None
for _ in range(reps):
    t.copy_(src)
    if kind == "pairs":
        s.sort_pairs(t, v)
    else:
        s.sort_keys(t)
torch.cuda.synchronize()
assert s.validate(t) == 0
print("ok")
