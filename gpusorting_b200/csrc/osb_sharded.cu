// osb_sharded.cu -- multi-GPU sharded sort (MSD bucket exchange + local OneSweep).  Placeholder: filled in
// after the single-GPU path is parity-green; until then every entry point reports OSB200_ERR_UNSUPPORTED.
#include "../../include/onesweep_b200.h"

extern "C" {
int osb200_sharded_unique_id(void*) { return OSB200_ERR_UNSUPPORTED; }
int osb200_sharded_create(osb200_sharded_handle* out, const void*, int, int, uint64_t, int)
{
    if (out) *out = nullptr;
    return OSB200_ERR_UNSUPPORTED;
}
int osb200_sharded_destroy(osb200_sharded_handle) { return OSB200_ERR_UNSUPPORTED; }
int osb200_sharded_sort_keys_u32(osb200_sharded_handle, const uint32_t*, uint64_t, uint32_t**, uint64_t*, void*)
{
    return OSB200_ERR_UNSUPPORTED;
}
int osb200_sharded_last_timing(osb200_sharded_handle, float*) { return OSB200_ERR_UNSUPPORTED; }
}
