"""Development probe: where do the rare order violations of the atomic-ranked PAIRS kernel happen?  Sorts 2^28 uniform pairs
(payload = input index) and, for every output position whose payload differs from the stable reference, decodes the pass-0
coordinates (tile, warp, round, lane) of the two input indices involved."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpusorting_b200 as g  # noqa: E402
SIGN = -(1 << 31)
n = 1 << 28
for rep in range(int(os.environ.get("OSB_REPS", "2"))):
    k = torch.empty(n, dtype=torch.int32, device="cuda"); v = torch.empty(n, dtype=torch.int32, device="cuda")
    g.init_random(k, 0, 21 + rep, payload=v, payload_is_index=True)
    k0 = k.clone()
    want_k, order = torch.sort(k ^ SIGN, stable=True)
    if os.environ.get("OSB_KEYS_ONLY"):
        with g.OneSweepSorter(n, 4, 0) as s:
            s.sort_keys(k)
        print(f"rep {rep}: keys-only mismatches {int((k != (want_k ^ SIGN)).sum())}", flush=True)
        del k, v, k0, want_k, order
        torch.cuda.empty_cache()
        continue
    with g.OneSweepSorter(n, 4, 4) as s:
        s.set_option("rank_mode", int(os.environ.get("OSB_RANK_MODE", "0")))
        if os.environ.get("OSB_SPIN_CAP"):
            s.set_option("spin_cap", int(os.environ["OSB_SPIN_CAP"]))
        s.sort_pairs(k, v)
    got = v.to(torch.int64) & 0xFFFFFFFF
    bad = (got != order).nonzero().flatten()
    print(f"rep {rep}: {bad.numel()} payload mismatches")
    if bad.numel() == 0 or os.environ.get("OSB_BRIEF"):
        continue
    a, b = got[bad], order[bad]          # input indices found / expected at the bad positions
    same_instr = ((a >> 5) == (b >> 5))
    same_tile = ((a >> 14) == (b >> 14))
    ka, kb = k0[a], k0[b]
    x = (ka ^ kb).to(torch.int64) & 0xFFFFFFFF
    print("  same pass-0 warp-instruction:", int(same_instr.sum()), " same pass-0 tile:", int(same_tile.sum()), " equal keys:", int((x == 0).sum()))
    for byte in range(4):
        print(f"  keys share byte {byte}: {int((((x >> (8 * byte)) & 255) == 0).sum())}", end="")
    print()
    w = ((a >> 10) & 15)
    print("  warp-in-tile histogram of the found index:", torch.bincount(w, minlength=16).tolist())
    r = ((a >> 5) & 31)
    print("  round histogram:", torch.bincount(r, minlength=32).tolist())
    print("  lane distance |la-lb| histogram (same instr only):", torch.bincount(((a & 31) - (b & 31)).abs()[same_instr], minlength=32).tolist())
    print("  first 8:", [(int(p), int(ai), int(bi), hex(int(kk) & 0xFFFFFFFF), hex(int(kk2) & 0xFFFFFFFF)) for p, ai, bi, kk, kk2 in zip(bad[:8], a[:8], b[:8], ka[:8], kb[:8])])
    del k, v, k0, want_k, order, got
    torch.cuda.empty_cache()
