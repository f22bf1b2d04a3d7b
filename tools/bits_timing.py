"""Development timing: osb200_sort_bits(0, 29/31) against the full-key sort at 2^30 (the sharded path's local sort)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpusorting_b200 as g  # noqa: E402
n = 1 << 30
src = torch.empty(n, dtype=torch.int32, device="cuda"); g.init_random(src, 0, 10)
w = src.clone()
with g.OneSweepSorter(n, 4, 0) as s:
    s.set_option("profile", 1)
    for name, fn in (("sort_keys", lambda: s.sort_keys(w)), ("sort_bits(0,31)", lambda: s.sort_bits(w, 0, 31)), ("sort_bits(0,29)", lambda: s.sort_bits(w, 0, 29)), ("sort_bits(3,32)", lambda: s.sort_bits(w, 3, 32))):
        for _ in range(3):
            w.copy_(src); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); b.synchronize()
        print(f"{name}: {a.elapsed_time(b):.3f} ms; per kernel {[round(x, 3) for x in s.last_profile()]}", flush=True)
