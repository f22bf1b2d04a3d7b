"""Generate tests/golden/ref_onesweep_golden.json from THE REFERENCE ITSELF.

Run on the GPU box (the reference's kernels are CUDA):
    gpurun -- 'python tests/golden/make_ref_golden.py gpurun_out/ref_onesweep_golden.json'
then copy the JSON into tests/golden/.  It drives oracle/_ref/libref_onesweep.so -- the reference's own
OneSweep.cu / UtilityKernels.cuh compiled from /root/reference by `make -C oracle ref` -- through
oracle/ref_harness.cu: InitRandom<<<256,256>>> (UtilityKernels.cuh:53-117) then the dispatcher's launch
order (OneSweepDispatcher.cuh:311-363), and records small order-sensitive digests (FNV-1a 64) plus the
first/last elements, for sizes the CPU oracle re-computes in seconds.  No product code is involved.
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import oraclelib  # noqa: E402

CASES = [  # (n, and_count, seed, pairs)
    (7680, 0, 7680, False),      # the first size of the reference's TestAllKeysOnly sweep (seed == n)
    (7681, 0, 7681, False),
    (12345, 0, 12345, True),
    (15360, 0, 15360, False),
    (65537, 0, 10, False),       # wraps the 65,536-stream generator
    (1 << 20, 0, 10, False),     # BASELINE.json configs[0]
    (1 << 20, 0, 10, True),
    (1 << 20, 1, 10, False),     # entropy presets 2..5 (Thearling-Smith AND-ing)
    (1 << 20, 2, 10, False),
    (1 << 20, 3, 10, False),
    (1 << 20, 4, 10, False),
    (1 << 22, 0, 22, False),
]


def main(out_path):
    ref = oraclelib.load_ref()
    orc = oraclelib.load_oracle()
    assert ref is not None, "oracle/_ref/libref_onesweep.so missing"
    max_n = max(c[0] for c in CASES)
    h = ref.lib.ref_create(max_n)
    assert h
    sort = torch.empty(max_n, dtype=torch.int32, device="cuda")
    alt = torch.empty_like(sort)
    pay = torch.empty_like(sort)
    altpay = torch.empty_like(sort)
    out = {"generator": "tests/golden/make_ref_golden.py", "source": "reference CUDA kernels via oracle/_ref", "cases": []}
    for n, andc, seed, pairs in CASES:
        if pairs:
            assert ref.lib.ref_init_random_pairs(sort.data_ptr(), pay.data_ptr(), n, andc, seed) == 0
        else:
            assert ref.lib.ref_init_random_keys(sort.data_ptr(), n, andc, seed) == 0
        torch.cuda.synchronize()
        inp = sort[:n].cpu().numpy().view(np.uint32).copy()
        if pairs:
            assert ref.lib.ref_sort_pairs(h, sort.data_ptr(), pay.data_ptr(), alt.data_ptr(), altpay.data_ptr(), n) == 0
        else:
            assert ref.lib.ref_sort_keys(h, sort.data_ptr(), alt.data_ptr(), n) == 0
        torch.cuda.synchronize()
        res = sort[:n].cpu().numpy().view(np.uint32).copy()
        hist = np.empty(1024, np.uint32)
        assert ref.lib.ref_get_global_histogram(h, hist.ctypes.data) == 0
        errs = int(ref.lib.ref_validate_keys(h, sort.data_ptr(), n))
        case = {
            "n": n, "and_count": andc, "seed": seed, "pairs": pairs,
            "input_head": [int(x) for x in inp[:8]], "input_digest": orc.digest(inp),
            "sorted_head": [int(x) for x in res[:8]], "sorted_tail": [int(x) for x in res[-8:]],
            "sorted_digest": orc.digest(res), "global_hist_digest": orc.digest(hist.astype(np.uint64)),
            "ref_validate_errors": errs,
        }
        if pairs:
            pres = pay[:n].cpu().numpy().view(np.uint32).copy()
            case["payload_digest"] = orc.digest(pres)
        out["cases"].append(case)
        print(n, andc, seed, pairs, "ok", hex(case["sorted_digest"]))
    ref.lib.ref_destroy(h)
    with open(out_path, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ref_onesweep_golden.json")
