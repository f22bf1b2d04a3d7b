"""Development check: bit-exact keys-only and pairs sorts at large n against torch (stable) sort; prints the error pattern.
env: OSB_RANK_MODE (0 atomic, 1 ballot), OSB_REPS, OSB_E (log2 n), OSB_KEYS (1: also keys-only)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpusorting_b200 as g  # noqa: E402

SIGN = -(1 << 31)
mode = int(os.environ.get("OSB_RANK_MODE", "0"))
reps = int(os.environ.get("OSB_REPS", "2"))
e = int(os.environ.get("OSB_E", "28"))
n = 1 << e
tot = {"keys": 0, "pairs": 0}
for rep in range(reps):
    for mask in (-1, 0xFFFFF):
        k = torch.empty(n, dtype=torch.int32, device="cuda")
        v = torch.empty(n, dtype=torch.int32, device="cuda")
        g.init_random(k, 0, 11 + rep, payload=v, payload_is_index=True)
        k &= mask
        want_k, order = torch.sort(k ^ SIGN, stable=True)
        want_k ^= SIGN
        if os.environ.get("OSB_KEYS", "1") == "1":
            k2 = k.clone()
            with g.OneSweepSorter(n, 4, 0) as s:
                s.set_option("rank_mode", mode)
                s.sort_keys(k2)
            nb = int((k2 != want_k).sum())
            tot["keys"] += nb
            del k2
        with g.OneSweepSorter(n, 4, 4) as s:
            s.set_option("rank_mode", mode)
            s.sort_pairs(k, v)
        nk = int((k != want_k).sum())
        nv = int(((v.to(torch.int64) & 0xFFFFFFFF) != order).sum())
        tot["pairs"] += nk + nv
        print(f"rep {rep} 2^{e} mask {mask:#x} rank_mode {mode}: pairs key-mismatches {nk} payload-mismatches {nv}", flush=True)
        del want_k, order, k, v
        torch.cuda.empty_cache()
print("exact_check totals", tot)
