"""Development probe (library built with OSB_EXP bit 7 = 128): cross-checks every lookback result of the PAIRS kernel."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpusorting_b200 as g  # noqa: E402
lib = ctypes.CDLL(os.environ["OSB200_LIB"])
n = 1 << 28
out = (ctypes.c_ulonglong * 16)()
for rep in range(3):
    k = torch.empty(n, dtype=torch.int32, device="cuda"); v = torch.empty(n, dtype=torch.int32, device="cuda")
    g.init_random(k, 0, 11 + rep, payload=v, payload_is_index=True)
    lib.osb200_debug_phases(out, 1)
    with g.OneSweepSorter(n, 4, 4) as s:
        s.sort_pairs(k, v)
        torch.cuda.synchronize()
    lib.osb200_debug_phases(out, 0)
    print(f"rep {rep}: wrong priors {out[11]} first: tile {out[12]} digit {out[13]} got {out[14]} want {out[15]}; windows/tile(d0) {out[9]/(4*n/16384):.2f} stalls {out[10]}")
