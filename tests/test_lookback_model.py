"""CPU model of the chained-scan lookback of digit_binning_wide_kernel (gpusorting_b200/csrc/osb_kernels.cu, lookback_wide):
the window arithmetic -- blocks of 8 reductions, the nearest block partial, one inclusive-prefix probe per block, the restart
at a predecessor that is not ready yet -- restated in Python and run against randomly progressing predecessor states.
Reference for the protocol itself: OneSweep.cu:306-327 (one predecessor per step).  No GPU, no library call."""
import random


def lookback(tile, nblk, agg, incl, advance):
    """agg[t]: None or the tile's digit count; incl[t]: None or the inclusive prefix over tiles 0..t.
    advance() lets other tiles make progress between two round trips."""
    s, cur = 0, tile - 1
    while True:
        if cur < 0:
            return s
        b0 = cur >> 3
        v = [[agg[(b0 - k) * 8 + t] if (b0 - k) >= 0 and (b0 - k) * 8 + t < len(agg) else (0 if (b0 - k) < 0 else None)
              for t in range(8)] for k in range(nblk)]
        c = [incl[(b0 - k) * 8 - 1] if (b0 - k) > 0 else None for k in range(nblk)]
        run, nxt, stalled, hi0 = s, (b0 - nblk + 1) * 8 - 1, False, cur & 7
        for k in range(nblk):
            b = b0 - k
            if b < 0:
                return run
            for t in range(7, -1, -1):
                if k == 0 and t > hi0:
                    continue
                a = v[k][t]
                if a is None:
                    nxt, stalled = b * 8 + t, True
                    break
                run += a
            if stalled:
                break
            if b == 0:
                return run
            if c[k] is not None:
                return run + c[k]
        s, cur = run, nxt
        advance()


def test_lookback_window_model_random_schedules():
    rng = random.Random(1)
    for _ in range(4000):
        ntiles = rng.randint(1, 90)
        counts = [rng.randint(0, 100) for _ in range(ntiles)]
        tile = ntiles - 1 if rng.random() < 0.5 else rng.randint(0, ntiles - 1)
        nblk = rng.choice([1, 2, 3, 4, 6])
        agg = [counts[t] if rng.random() < 0.6 else None for t in range(ntiles)]
        incl = [sum(counts[:t + 1]) if agg[t] is not None and rng.random() < 0.3 else None for t in range(ntiles)]

        def advance():
            for t in range(tile):
                if agg[t] is None and rng.random() < 0.5:
                    agg[t] = counts[t]
                if agg[t] is not None and incl[t] is None and rng.random() < 0.2:
                    incl[t] = sum(counts[:t + 1])

        assert lookback(tile, nblk, agg, incl, advance) == sum(counts[:tile])
