"""In-situ cost of the phases of the wide DigitBinningPass: duplicate/remove one phase, time the difference."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpusorting_b200 as g
e = int(sys.argv[1]) if len(sys.argv) > 1 else 30
n = 1 << e
src = torch.empty(n, dtype=torch.int32, device="cuda"); g.init_random(src, 0, 10)
dst = torch.empty_like(src)
s = g.OneSweepSorter(n, 4, 0)
s.set_option("profile", 1)
def pass_ms(ab):
    s.set_option("ablate", ab)
    best = []
    for _ in range(4):
        s.digit_binning_pass(src, dst, 8)
        torch.cuda.synchronize()
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # time only the binning kernel: hist+scan are launched first by the entry point; subtract via a hist-only timing
        a.record(); s.digit_binning_pass(src, dst, 8); b.record(); b.synchronize()
        best.append(a.elapsed_time(b))
    s.set_option("ablate", 0)
    return min(best)
hist = min((lambda: [ (lambda a,b: (a.record(), s.global_histogram(src), b.record(), b.synchronize(), a.elapsed_time(b))[-1])(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)])())
base = pass_ms(0)
print(f"n=2^{e}: hist {hist:.3f} ms; hist+scan+pass {base:.3f} ms -> pass ~{base-hist:.3f} ms")
for name, ab, div in [("count x3 (2 extra non-returning atomics/key)", 1, 2), ("transposing STS twice", 2, 1),
                      ("extra LDS sweep of sorted tile", 4, 1), ("extra returning atomic/key", 16, 1), ("NO global stores", 8, -1)]:
    t = pass_ms(ab)
    d = (t - base) / abs(div)
    print(f"  {name:48s}: {t:.3f} ms  delta/op {d*1:+.3f} ms  ({d/(base-hist)*100:+.1f}% of pass)")
