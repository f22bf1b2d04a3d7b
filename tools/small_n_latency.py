"""Development timing: latency of small sorts (one launch of the segment kernel against the ordinary 6-launch path) and
throughput of the segmented sort."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpusorting_b200 as g  # noqa: E402

def t_us(fn, reps=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) * 1000.0 / reps

with g.OneSweepSorter(1 << 17, 4, 4) as s:
    for e in (8, 10, 12, 13, 14, 15, 16):
        n = 1 << e
        k = torch.empty(n, dtype=torch.int32, device="cuda"); g.init_random(k, 0, 5)
        v = torch.arange(n, dtype=torch.int32, device="cuda")
        row = [f"n=2^{e}"]
        for small in (1, 0):
            s.set_option("small_path", small)
            row.append(f"small_path={small}: keys {t_us(lambda: s.sort_keys(k)):7.1f} us  pairs {t_us(lambda: s.sort_pairs(k, v)):7.1f} us")
        print("  ".join(row), flush=True)
    s.set_option("small_path", 1)
    for seglen, segs in ((32, 1 << 21), (256, 1 << 19), (2048, 1 << 16), (16384, 1 << 13)):
        n = seglen * segs
        k = torch.empty(n, dtype=torch.int32, device="cuda"); g.init_random(k, 0, 9)
        v = torch.arange(n, dtype=torch.int32, device="cuda")
        offs = torch.arange(segs + 1, dtype=torch.int64, device="cuda") * seglen
        src = k.clone()
        us = t_us(lambda: (k.copy_(src), s.segmented_sort(k, offs, v, max_segment_len=seglen)), reps=10)
        cp = t_us(lambda: k.copy_(src), reps=10)
        print(f"segmented pairs: {segs} segments x {seglen} keys: {(us - cp) / 1000:.3f} ms -> {n / (us - cp) / 1e3:.2f} Gpairs/s", flush=True)
