// Host-side unit checks of the device headers' pure functions (compiled with nvcc, executed on the CPU; no GPU needed).
//   descriptor packing: value / epoch / flag fields never bleed into each other (the first fused exchange died of
//   exactly that: peer ADDRESSES were fed through the 38-bit value field and corrupted the epoch);
//   key codec: encode is an order-preserving bijection and decode inverts it, for every key type and direction.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>
#include "../gpusorting_b200/csrc/osb_common.cuh"

// host copies of the two device-only codec helpers (same expressions as osb_common.cuh)
template <typename K> static K enc(K k, K a, K b, K d)
{
    using S = typename std::make_signed<K>::type;
    const K sar = static_cast<K>(static_cast<S>(k) >> (sizeof(K) * 8 - 1));
    return k ^ (((sar & a) | b) ^ d);
}
template <typename K> static K dec(K e, K a, K b, K d)
{
    using S = typename std::make_signed<K>::type;
    e ^= d;
    const K sar = static_cast<K>(static_cast<S>(e) >> (sizeof(K) * 8 - 1));
    return e ^ ((~sar & a) | b);
}

static int fails = 0;
#define CHECK(c) do { if (!(c)) { printf("FAILED %s:%d %s\n", __FILE__, __LINE__, #c); ++fails; } } while (0)

int main()
{
    using namespace osb;
    // ---- descriptors ----
    const uint64_t vmax = (1ull << kValueBits) - 1;
    for (uint32_t epoch : std::vector<uint32_t>{1u, 2u, 12345u, kEpochMax}) for (uint64_t flag : std::vector<uint64_t>{kFlagReduction, kFlagInclusive}) for (uint64_t v : std::vector<uint64_t>{0ull, 1ull, 16384ull, (1ull << 30), vmax}) {
        const uint64_t d = desc_pack(epoch, flag, v);
        CHECK(desc_epoch(d) == epoch); CHECK((d & kFlagMask) == flag); CHECK(desc_value(d) == v);
    }
    {   // an out-of-range value must not touch the epoch or the flag (it is truncated, and callers keep values < n)
        const uint64_t d = desc_pack(77, kFlagInclusive, (0x7f12ull << 40) | 123);
        CHECK(desc_epoch(d) == 77); CHECK((d & kFlagMask) == kFlagInclusive); CHECK(desc_value(d) == 123);
    }
    CHECK(desc_epoch(0) == 0);  // freshly zeroed memory never matches a live epoch (epochs start at 1)

    // ---- codec ----
    struct C { uint64_t a, b; const char* name; int bits; };
    const C kinds[] = {{0, 0, "u32", 32}, {0, 1ull << 31, "i32", 32}, {0xffffffffull, 1ull << 31, "f32", 32},
                       {0, 0, "u64", 64}, {0, 1ull << 63, "i64", 64}, {~0ull, 1ull << 63, "f64", 64}};
    uint64_t s = 88172645463325252ull;
    for (const C& c : kinds) for (int desc = 0; desc < 2; ++desc) {
        std::vector<uint64_t> bits;
        for (int i = 0; i < 20000; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; bits.push_back(c.bits == 32 ? (s & 0xffffffffull) : s); }
        for (uint64_t e : std::vector<uint64_t>{0ull, 1ull, 0x7fffffffull, 0x80000000ull, 0xffffffffull, 0x7f800000ull, 0xff800000ull}) bits.push_back(c.bits == 32 ? e : e << 32);
        const uint64_t all = c.bits == 32 ? 0xffffffffull : ~0ull, dd = desc ? all : 0;
        std::vector<std::pair<uint64_t, uint64_t>> pairs;
        for (uint64_t k : bits) {
            uint64_t e = c.bits == 32 ? enc<uint32_t>((uint32_t)k, (uint32_t)c.a, (uint32_t)c.b, (uint32_t)dd) : enc<uint64_t>(k, c.a, c.b, dd);
            uint64_t r = c.bits == 32 ? dec<uint32_t>((uint32_t)e, (uint32_t)c.a, (uint32_t)c.b, (uint32_t)dd) : dec<uint64_t>(e, c.a, c.b, dd);
            CHECK(r == k);
            pairs.push_back({e, k});
        }
        std::sort(pairs.begin(), pairs.end());
        // ascending encoded order must be the typed order
        for (size_t i = 1; i < pairs.size(); ++i) {
            const uint64_t x = pairs[i - 1].second, y = pairs[i].second;
            bool le;
            if (c.name[0] == 'u') le = x <= y;
            else if (c.name[0] == 'i') le = c.bits == 32 ? (int32_t)x <= (int32_t)y : (int64_t)x <= (int64_t)y;
            else {  // IEEE total order by sign-magnitude
                auto key = [&](uint64_t v) { const bool neg = (v >> (c.bits - 1)) & 1; const uint64_t mag = v & (all >> 1); return neg ? -(__int128)mag - 1 : (__int128)mag; };
                le = key(x) <= key(y);
            }
            if (desc) { // descending: reverse relation
                if (c.name[0] == 'u') le = x >= y;
                else if (c.name[0] == 'i') le = c.bits == 32 ? (int32_t)x >= (int32_t)y : (int64_t)x >= (int64_t)y;
                else { auto key = [&](uint64_t v) { const bool neg = (v >> (c.bits - 1)) & 1; const uint64_t mag = v & (all >> 1); return neg ? -(__int128)mag - 1 : (__int128)mag; }; le = key(x) >= key(y); }
            }
            if (!le) { printf("order violated for %s desc=%d\n", c.name, desc); ++fails; break; }
        }
    }
    printf(fails ? "host_unit: %d FAILURES\n" : "host_unit: ok\n", fails);
    return fails ? 1 : 0;
}
