"""Development check for tools/sweep_run.sh: bit-exact keys-only sorts against torch.sort at a full and a ragged size, for the
library named by OSB200_LIB (typed descending keys included, so the decode-on-store scatter path runs too)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpusorting_b200 as g  # noqa: E402

ok = True
for n, andc in (((1 << 24) + 12345, 0), (1 << 22, 3), ((1 << 20) + 1, 0)):
    src = torch.empty(n, dtype=torch.int32, device="cuda")
    g.init_random(src, andc, 7)
    want = (torch.sort(src ^ -(1 << 31)).values) ^ -(1 << 31)
    with g.OneSweepSorter(n, 4, 0) as s:
        got = src.clone()
        s.sort_keys(got)
        ok &= bool(torch.equal(got, want))
        got = src.clone()
        s.sort_keys_typed(got, "i32", descending=True)
        ok &= bool(torch.equal(got, torch.sort(src, descending=True, stable=True).values))
print("sweep_check", "OK" if ok else "MISMATCH", flush=True)
