"""Development check: bit-exact (key, payload=index) sorts against torch's stable sort, library from OSB200_LIB."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpusorting_b200 as g  # noqa: E402

ok = True
for e, mask in ((20, 0xFFFF), (24, 0xFFFFFF), (28, 0xFFFFF), (28, -1), (30, 0xFFFFF)):
    n = (1 << e) + (777 if e < 28 else 0)
    k = torch.empty(n, dtype=torch.int32, device="cuda")
    v = torch.empty(n, dtype=torch.int32, device="cuda")
    g.init_random(k, 0, 11, payload=v, payload_is_index=True)
    k &= mask
    want_k, order = torch.sort(k ^ -(1 << 31), stable=True)
    want_k ^= -(1 << 31)
    with g.OneSweepSorter(n, 4, 4) as s:
        s.sort_pairs(k, v)
    good = bool(torch.equal(k, want_k)) and bool(torch.equal(v.to(torch.int64) & 0xFFFFFFFF, order))
    print(f"pairs 2^{e} mask {mask:#x}: {'OK' if good else 'MISMATCH'}", flush=True)
    ok &= good
    del want_k, order, k, v
    torch.cuda.empty_cache()
print("pairs_check", "OK" if ok else "MISMATCH")
