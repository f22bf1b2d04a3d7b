"""Development probe: per-phase wall clocks of the wide DigitBinningPass (library built with OSB_EXP bit 5 = 32)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpusorting_b200 as g  # noqa: E402

lib = ctypes.CDLL(os.environ["OSB200_LIB"])
n = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 30
src = torch.empty(n, dtype=torch.int32, device="cuda")
g.init_random(src, 0, 10)
work = src.clone()
s = g.OneSweepSorter(n, 4, 0)
s.sort_keys(work)
torch.cuda.synchronize()
out = (ctypes.c_ulonglong * 16)()
lib.osb200_debug_phases(out, 1)
work.copy_(src)
s.sort_keys(work)
torch.cuda.synchronize()
lib.osb200_debug_phases(out, 0)
names = ["ticket+clear", "load", "count", "reduce/scan/bases", "rank", "lookback", "scatter", None, "barrier after lookback"]
ctas = out[7]
tot = sum(out[i] for i in range(9) if i != 7)
print(f"CTAs {ctas}; mean clocks per CTA-tile: total {tot / ctas:.0f}")
for i, nm in enumerate(names):
    if nm is None:
        continue
    print(f"  {nm:20s} {out[i] / ctas:9.0f} clk  {100.0 * out[i] / tot:5.1f}%")
print(f"lookback windows per tile (digit 0): {out[9] / ctas:.2f}; stalled polls per tile: {out[10] / ctas:.2f}")
