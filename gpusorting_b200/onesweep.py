"""Host-side mirror of the reference's OneSweep interface, on top of the C-ABI.

Reference (paths relative to /root/reference):
  * class OneSweepDispatcher(bool keysOnly, uint32_t maxSize), GPUSortingCUDA/Sort/OneSweepDispatcher.cuh:17-392
    -- TestAllKeysOnly / TestAllPairs / BatchTimingKeysOnly / BatchTimingPairs keep their names and argument
    meaning here so the parity tests read like the reference's own.
  * OneSweep.Sort(...), GPUSortingUnity/Runtime/OneSweep.cs:297-306,358-370 -- the only public `Sort` in the
    reference; `Sort(keys[, values], n)` below is BASELINE.json's north-star shape of it.

PyTorch is used for device memory and streams only; every byte of sorting work happens in
libonesweep_b200.so.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import check, lib

ENTROPY_PRESET_1, ENTROPY_PRESET_2, ENTROPY_PRESET_3, ENTROPY_PRESET_4, ENTROPY_PRESET_5 = range(5)

# sort_keys / sort_pairs order keys by their UNSIGNED bit pattern: only integer containers are accepted there (a float
# tensor would silently sort negatives wrongly).  Float tensors go through sort_*_typed, which states the key type.
_KEY_DTYPES_4 = (torch.int32, torch.uint32)
_KEY_DTYPES_8 = (torch.int64, torch.uint64)
_TYPED_DTYPES_4 = (torch.int32, torch.uint32, torch.float32)
_TYPED_DTYPES_8 = (torch.int64, torch.uint64, torch.float64)
KEY_TYPES = {"u32": 0, "i32": 1, "f32": 2, "u64": 3, "i64": 4, "f64": 5}


def _stream_ptr(stream: Optional[torch.cuda.Stream]) -> int:
    s = stream if stream is not None else torch.cuda.current_stream()
    return int(s.cuda_stream)


def _check_dev_tensor(t: torch.Tensor, dtypes, name: str, n: Optional[int] = None, device: Optional[int] = None) -> None:
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.is_contiguous() and t.dtype in dtypes and t.dim() == 1):
        raise TypeError(f"{name} must be a contiguous 1-D CUDA tensor with dtype in {dtypes}")
    if device is not None and t.device.index != device:
        raise ValueError(f"{name} lives on cuda:{t.device.index}, the sorter on cuda:{device}")
    if n is not None and not (0 <= n <= t.numel()):
        raise ValueError(f"n={n} is outside 0..{name}.numel()={t.numel()}")


class OneSweepSorter:
    """Owns one C-ABI sorter handle (alt buffers, tile descriptors) for up to ``max_n`` elements.

    Keys are ordered by their UNSIGNED bit pattern, as in the reference CUDA path (uint32 keys,
    OneSweep.cuh:24-52); int32/int64 tensors are accepted as raw 32/64-bit containers.
    """

    def __init__(self, max_n: int, key_bytes: int = 4, value_bytes: int = 0, device: Optional[int] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("gpusorting_b200 needs a CUDA device (sm_100); there is no CPU fallback")
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.max_n, self.key_bytes, self.value_bytes = int(max_n), int(key_bytes), int(value_bytes)
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(lib.osb200_create(ctypes.byref(h), self.max_n, self.key_bytes, self.value_bytes), "osb200_create")
        self._h = h

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None):
            lib.osb200_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- options ----------------------------------------------------------------------------------
    def set_option(self, key: str, value: int) -> None:
        check(lib.osb200_set_option(self._h, key.encode(), int(value)), f"osb200_set_option({key})")

    def info(self, key: str) -> int:
        v = lib.osb200_get_info(self._h, key.encode())
        if v < 0:
            raise KeyError(key)
        return int(v)

    def last_profile(self):
        """Per-kernel milliseconds of the last sort when option 'profile' is on: [hist, scan, pass0, pass1, ...]."""
        buf = (ctypes.c_float * 16)()
        k = lib.osb200_get_profile(self._h, buf, 16)
        if k < 0:
            check(k, "osb200_get_profile")
        return [float(buf[i]) for i in range(k)]

    # -- device sorts -------------------------------------------------------------------------------
    # Every entry point validates dtype / device / contiguity / n against the tensors it is handed and makes the
    # sorter's device current for the launch (the C side launches on the current device).
    def _raw(self):
        return _KEY_DTYPES_4 if self.key_bytes == 4 else _KEY_DTYPES_8

    def _typed(self):
        return _TYPED_DTYPES_4 if self.key_bytes == 4 else _TYPED_DTYPES_8

    def sort_keys(self, keys: torch.Tensor, n: Optional[int] = None, stream=None) -> torch.Tensor:
        n = keys.numel() if n is None else int(n)
        _check_dev_tensor(keys, self._raw(), "keys", n, self.device)
        fn, what = ((lib.osb200_sort_keys_u32, "osb200_sort_keys_u32") if self.key_bytes == 4
                    else (lib.osb200_sort_keys_u64, "osb200_sort_keys_u64"))
        with torch.cuda.device(self.device):
            check(fn(self._h, keys.data_ptr(), n, _stream_ptr(stream)), what)
        return keys

    def sort_keys_typed(self, keys: torch.Tensor, key_type: str, descending: bool = False, n: Optional[int] = None,
                        stream=None) -> torch.Tensor:
        """Signed / float keys and descending order (reference HLSL: SortCommon.hlsl:134-154,594-656).  key_type in
        KEY_TYPES; it states how the bits are ordered, whatever the tensor dtype (which only has to have the width)."""
        n = keys.numel() if n is None else int(n)
        _check_dev_tensor(keys, self._typed(), "keys", n, self.device)
        with torch.cuda.device(self.device):
            check(lib.osb200_sort_keys_typed(self._h, keys.data_ptr(), n, KEY_TYPES[key_type], 1 if descending else 0,
                                             _stream_ptr(stream)), "osb200_sort_keys_typed")
        return keys

    def sort_pairs_typed(self, keys: torch.Tensor, values: torch.Tensor, key_type: str, descending: bool = False,
                         n: Optional[int] = None, stream=None):
        n = keys.numel() if n is None else int(n)
        _check_dev_tensor(keys, _TYPED_DTYPES_4, "keys", n, self.device)
        _check_dev_tensor(values, _TYPED_DTYPES_4, "values", n, self.device)
        with torch.cuda.device(self.device):
            check(lib.osb200_sort_pairs_typed(self._h, keys.data_ptr(), values.data_ptr(), n, KEY_TYPES[key_type],
                                              1 if descending else 0, _stream_ptr(stream)), "osb200_sort_pairs_typed")
        return keys, values

    def sort_bits(self, keys: torch.Tensor, begin_bit: int, end_bit: int, values: Optional[torch.Tensor] = None,
                  n: Optional[int] = None, stream=None):
        """Stable sort on the key bits [begin_bit, end_bit) only (osb200_sort_bits)."""
        n = keys.numel() if n is None else int(n)
        _check_dev_tensor(keys, self._raw(), "keys", n, self.device)
        if values is not None:
            _check_dev_tensor(values, _TYPED_DTYPES_4, "values", n, self.device)
        with torch.cuda.device(self.device):
            check(lib.osb200_sort_bits(self._h, keys.data_ptr(), values.data_ptr() if values is not None else None, n,
                                       int(begin_bit), int(end_bit), _stream_ptr(stream)), "osb200_sort_bits")
        return keys if values is None else (keys, values)

    def segmented_sort(self, keys: torch.Tensor, segment_offsets: torch.Tensor, values: Optional[torch.Tensor] = None,
                       max_segment_len: Optional[int] = None, stream=None):
        """Sort every segment [offsets[i], offsets[i+1]) of `keys` (and `values`) ascending and stable, in place, one thread
        block per segment (osb200_segmented_sort_u32; reference: SplitSort, SegSort/SplitSort/SplitSort.cuh:702-938).
        `segment_offsets`: int64 device tensor of num_segments + 1 offsets.  `max_segment_len` (default: computed here, which
        costs a device->host read) must not exceed 16,384."""
        _check_dev_tensor(keys, _KEY_DTYPES_4, "keys", keys.numel(), self.device)
        if values is not None:
            _check_dev_tensor(values, _TYPED_DTYPES_4, "values", keys.numel(), self.device)
        if segment_offsets.dtype != torch.int64 or not segment_offsets.is_cuda or not segment_offsets.is_contiguous():
            raise TypeError("segment_offsets must be a contiguous int64 CUDA tensor")
        segs = segment_offsets.numel() - 1
        if segs <= 0:
            return keys if values is None else (keys, values)
        if max_segment_len is None:
            max_segment_len = int((segment_offsets[1:] - segment_offsets[:-1]).max().item())
        with torch.cuda.device(self.device):
            check(lib.osb200_segmented_sort_u32(self._h, keys.data_ptr(), values.data_ptr() if values is not None else None,
                                                segment_offsets.data_ptr(), segs, int(max_segment_len), _stream_ptr(stream)),
                  "osb200_segmented_sort_u32")
        return keys if values is None else (keys, values)

    def sort_pairs(self, keys: torch.Tensor, values: torch.Tensor, n: Optional[int] = None, stream=None):
        n = keys.numel() if n is None else int(n)
        _check_dev_tensor(keys, _KEY_DTYPES_4, "keys", n, self.device)
        _check_dev_tensor(values, _TYPED_DTYPES_4, "values", n, self.device)  # payloads are opaque 32-bit words
        with torch.cuda.device(self.device):
            check(lib.osb200_sort_pairs_u32(self._h, keys.data_ptr(), values.data_ptr(), n, _stream_ptr(stream)),
                  "osb200_sort_pairs_u32")
        return keys, values

    # -- host-buffer sorts (end-to-end: H2D + sort + D2H inside the call) ---------------------------
    def sort_host(self, keys, values=None, n: Optional[int] = None):
        """keys/values: numpy arrays or CPU torch tensors (pinned or pageable), sorted in place."""
        kp, kn, kb = _host_ptr(keys)
        n = kn if n is None else int(n)
        if kb != self.key_bytes:
            raise TypeError("key width does not match the sorter")
        if values is None:
            fn = lib.osb200_sort_host_keys_u32 if kb == 4 else lib.osb200_sort_host_keys_u64
            check(fn(self._h, kp, n), "osb200_sort_host_keys")
        else:
            vp, vn, vb = _host_ptr(values)
            if vb != 4 or vn < n:
                raise TypeError("values must be 4-byte elements, at least n long")
            check(lib.osb200_sort_host_pairs_u32(self._h, kp, vp, n), "osb200_sort_host_pairs_u32")
        return keys if values is None else (keys, values)

    # -- kernel-level entry points (parity tests) ---------------------------------------------------
    def global_histogram(self, keys: torch.Tensor, n: Optional[int] = None, stream=None) -> torch.Tensor:
        n = keys.numel() if n is None else int(n)
        _check_dev_tensor(keys, self._typed(), "keys", n, self.device)
        hist = torch.empty(self.key_bytes * 256, dtype=torch.int64, device=keys.device)
        with torch.cuda.device(self.device):
            check(lib.osb200_global_histogram(self._h, keys.data_ptr(), n, hist.data_ptr(), _stream_ptr(stream)),
                  "osb200_global_histogram")
        return hist.view(self.key_bytes, 256)

    def digit_binning_pass(self, src: torch.Tensor, dst: torch.Tensor, radix_shift: int, src_values=None,
                           dst_values=None, n: Optional[int] = None, stream=None) -> None:
        n = src.numel() if n is None else int(n)
        _check_dev_tensor(src, self._typed(), "src", n, self.device)
        _check_dev_tensor(dst, self._typed(), "dst", n, self.device)
        if (src_values is None) != (dst_values is None):
            raise ValueError("src_values and dst_values go together")
        if src_values is not None:
            _check_dev_tensor(src_values, _TYPED_DTYPES_4, "src_values", n, self.device)
            _check_dev_tensor(dst_values, _TYPED_DTYPES_4, "dst_values", n, self.device)
        sv = src_values.data_ptr() if src_values is not None else None
        dv = dst_values.data_ptr() if dst_values is not None else None
        with torch.cuda.device(self.device):
            check(lib.osb200_digit_binning_pass(self._h, src.data_ptr(), dst.data_ptr(), sv, dv, n, int(radix_shift),
                                                _stream_ptr(stream)), "osb200_digit_binning_pass")

    def validate(self, keys: torch.Tensor, n: Optional[int] = None, stream=None) -> int:
        """Number of adjacent inversions (reference Validate, UtilityKernels.cuh:403-429); 0 == sorted."""
        n = keys.numel() if n is None else int(n)
        _check_dev_tensor(keys, self._typed(), "keys", n, self.device)
        err = ctypes.c_uint64(0)
        with torch.cuda.device(self.device):
            check(lib.osb200_validate(self._h, keys.data_ptr(), n, ctypes.byref(err), _stream_ptr(stream)), "osb200_validate")
        return int(err.value)


def _host_ptr(a):
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise TypeError("host array must be contiguous")
        return a.ctypes.data, a.size, a.dtype.itemsize
    if isinstance(a, torch.Tensor) and not a.is_cuda:
        if not a.is_contiguous():
            raise TypeError("host tensor must be contiguous")
        return a.data_ptr(), a.numel(), a.element_size()
    raise TypeError("expected a numpy array or a CPU torch tensor")


def init_random(keys: torch.Tensor, and_count: int, seed: int, n: Optional[int] = None,
                payload: Optional[torch.Tensor] = None, payload_is_index: bool = False, stream=None) -> None:
    """The reference's input generator InitRandom<<<256,256>>> (UtilityKernels.cuh:53-117), on the device."""
    n = keys.numel() if n is None else int(n)
    _check_dev_tensor(keys, _TYPED_DTYPES_4, "keys", n)
    if payload is not None:
        _check_dev_tensor(payload, _TYPED_DTYPES_4, "payload", n, keys.device.index)
    pp = payload.data_ptr() if payload is not None else None
    check(lib.osb200_init_random_u32(keys.data_ptr(), pp, n, int(and_count), int(seed) & 0xFFFFFFFF,
                                     1 if payload_is_index else 0, _stream_ptr(stream)), "osb200_init_random_u32")


# --------------------------------------------------------------------------------------------------
# module-level Sort(keys[, values], n): the north-star call shape.  Sorters are cached per
# (device, key width, pairs) and grown on demand, so repeated calls do not re-allocate.
# --------------------------------------------------------------------------------------------------
_CACHE: dict = {}


def _cached_sorter(device: int, key_bytes: int, value_bytes: int, n: int, stream_ptr: int) -> OneSweepSorter:
    # one handle per stream: the ABI allows one sort in flight per handle, and sorts on one stream are ordered
    k = (device, key_bytes, value_bytes, stream_ptr)
    s = _CACHE.get(k)
    if s is None or s.max_n < n:
        if s is not None:
            s.close()
        s = OneSweepSorter(max(n, 1), key_bytes, value_bytes, device)
        _CACHE[k] = s
    return s


def Sort(keys: torch.Tensor, values: Optional[torch.Tensor] = None, n: Optional[int] = None, stream=None):
    """OneSweep::Sort(keys[, values], n): ascending, stable, in place; returns its arguments."""
    n = keys.numel() if n is None else int(n)
    kb = keys.element_size()
    if not (isinstance(keys, torch.Tensor) and keys.is_cuda):
        raise TypeError("keys must be a CUDA tensor")
    with torch.cuda.device(keys.device.index):
        sp = _stream_ptr(stream)
    s = _cached_sorter(keys.device.index, kb, 0 if values is None else 4, n, sp)
    if values is None:
        return s.sort_keys(keys, n, stream)
    return s.sort_pairs(keys, values, n, stream)


class OneSweepDispatcher:
    """Mirror of the reference's class OneSweepDispatcher (Sort/OneSweepDispatcher.cuh:17-392).

    Same constructor arguments (keysOnly, maxSize) and the same public methods; like the reference it owns
    m_sort / m_sortPayload and sorts them in place.  Tests print nothing unless ``verbose``.
    """

    k_partitionSize = 7680  # the reference's sweep bounds (OneSweepDispatcher.cuh:23,98) are kept for TestAll*

    def __init__(self, keysOnly: bool, maxSize: int, verbose: bool = False):
        self.k_keysOnly, self.k_maxSize, self.verbose = bool(keysOnly), int(maxSize), verbose
        self.m_sort = torch.empty(self.k_maxSize, dtype=torch.int32, device="cuda")
        self.m_sortPayload = None if keysOnly else torch.empty(self.k_maxSize, dtype=torch.int32, device="cuda")
        self._sorter = OneSweepSorter(self.k_maxSize, 4, 0 if keysOnly else 4)

    # reference: DispatchKernelsKeysOnly / DispatchKernelsPairs (private there; :311-363)
    def DispatchKernelsKeysOnly(self, size: int) -> None:
        self._sorter.sort_keys(self.m_sort, size)

    def DispatchKernelsPairs(self, size: int) -> None:
        self._sorter.sort_pairs(self.m_sort, self.m_sortPayload, size)

    # reference: DispatchValidateKeys / DispatchValidatePairs (:365-391)
    def DispatchValidateKeys(self, size: int) -> bool:
        return self._sorter.validate(self.m_sort, size) == 0

    def DispatchValidatePairs(self, size: int) -> bool:
        # the reference relies on payload == key (UtilityKernels.cuh:432-479)
        return self._sorter.validate(self.m_sort, size) == 0 and self._sorter.validate(self.m_sortPayload, size) == 0

    def _sizes(self, small_step: int, large_exps):
        for i in range(self.k_partitionSize, self.k_partitionSize * 2 + 1, small_step):
            yield i, i
        for e in large_exps:
            if (1 << e) <= self.k_maxSize:
                yield 1 << e, e

    def TestAllKeysOnly(self, small_step: int = 1, large_exps=(26, 27, 28)) -> tuple[int, int]:
        """reference :87-134 -- every n in [7680, 15360] then 2^26..2^28; returns (passed, total)."""
        passed = total = 0
        for n, seed in self._sizes(small_step, large_exps):
            init_random(self.m_sort, ENTROPY_PRESET_1, seed, n)
            self.DispatchKernelsKeysOnly(n)
            ok = self.DispatchValidateKeys(n)
            passed += ok
            total += 1
            if not ok and self.verbose:
                print(f"Test failed at size {n}")
        return passed, total

    def TestAllPairs(self, small_step: int = 1, large_exps=(26, 27, 28)) -> tuple[int, int]:
        """reference :136-191"""
        passed = total = 0
        for n, seed in self._sizes(small_step, large_exps):
            init_random(self.m_sort, ENTROPY_PRESET_1, seed, n, payload=self.m_sortPayload)
            self.DispatchKernelsPairs(n)
            ok = self.DispatchValidatePairs(n)
            passed += ok
            total += 1
            if not ok and self.verbose:
                print(f"Test failed at size {n}")
        return passed, total

    def _batch(self, size: int, batchCount: int, seed: int, entropyPreset: int, pairs: bool) -> float:
        if size > self.k_maxSize:
            raise ValueError("Error, requested test size exceeds max initialized size.")
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        total_ms = 0.0
        for i in range(batchCount + 1):  # i == 0 is the discarded warm-up (reference :228)
            init_random(self.m_sort, entropyPreset, i + seed, size, payload=self.m_sortPayload if pairs else None)
            torch.cuda.synchronize()
            start.record()
            (self.DispatchKernelsPairs if pairs else self.DispatchKernelsKeysOnly)(size)
            stop.record()
            stop.synchronize()
            if i:
                total_ms += start.elapsed_time(stop)
        keys_per_sec = size / (total_ms / 1000.0) * batchCount
        if self.verbose:
            print(f"Total time elapsed: {total_ms / 1000.0}\nEstimated speed at {size} 32-bit elements: {keys_per_sec:E} keys/sec")
        return keys_per_sec

    def BatchTimingKeysOnly(self, size: int, batchCount: int, seed: int, entropyPreset: int = ENTROPY_PRESET_1) -> float:
        """reference :193-239; returns keys/sec."""
        return self._batch(size, batchCount, seed, entropyPreset, False)

    def BatchTimingPairs(self, size: int, batchCount: int, seed: int, entropyPreset: int = ENTROPY_PRESET_1) -> float:
        """reference :241-293; returns pairs/sec."""
        if self.k_keysOnly:
            raise ValueError("Error, object was initialized for keys only.")
        return self._batch(size, batchCount, seed, entropyPreset, True)
