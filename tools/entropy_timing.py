"""Development timing + exactness: entropy presets 1-5 (keys and pairs) with the HOT instantiation on and off."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpusorting_b200 as g  # noqa: E402
SIGN = -(1 << 31)
e = int(os.environ.get("OSB_E", "30"))
n = 1 << e
src = torch.empty(n, dtype=torch.int32, device="cuda")
w = torch.empty_like(src)
def t_ms(fn, prep, reps=3):
    best = 1e9
    for _ in range(reps):
        prep(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
with g.OneSweepSorter(n, 4, 0) as s:
    for andc in range(5):
        g.init_random(src, andc, 10)
        row = [f"preset {andc + 1}:"]
        for hot in (1, 0):
            s.set_option("hot_passes", hot)
            ms = t_ms(lambda: s.sort_keys(w), lambda: w.copy_(src))
            row.append(f"hot_passes={hot} {ms:7.3f} ms {n / ms / 1e6:6.1f} Gkeys/s (hot mask {s.info('last_hot_mask'):04b})")
        print("  ".join(row), flush=True)
    s.set_option("hot_passes", 1)
# exactness, keys and pairs, 2^27 + ragged
m = (1 << 27) + 4321
for andc in (2, 3, 4):
    k = torch.empty(m, dtype=torch.int32, device="cuda"); v = torch.empty(m, dtype=torch.int32, device="cuda")
    g.init_random(k, andc, 77, payload=v, payload_is_index=True)
    want, order = torch.sort(k ^ SIGN, stable=True); want ^= SIGN
    k2 = k.clone()
    with g.OneSweepSorter(m, 4, 0) as s:
        s.sort_keys(k2); hm = s.info("last_hot_mask")
    okk = bool(torch.equal(k2, want))
    with g.OneSweepSorter(m, 4, 4) as s:
        s.sort_pairs(k, v)
    okp = bool(torch.equal(k, want)) and bool(torch.equal(v.to(torch.int64) & 0xFFFFFFFF, order))
    print(f"exact preset {andc + 1}: keys {'OK' if okk else 'MISMATCH'} pairs {'OK' if okp else 'MISMATCH'} hot mask {hm:04b}", flush=True)
    del k, v, k2, want, order
    torch.cuda.empty_cache()
