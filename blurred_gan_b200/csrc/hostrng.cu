// Host-only: bit-exact replay of numpy's legacy RandomState.randint on a caller-held MT19937 state.
//
// The reference draws its patch positions with np.random.randint from numpy's global legacy generator
// (sliced_wasserstein.py:20-21), 2 x levels calls of n * 128 draws per fed minibatch, and parity ("identical seeds give
// identical patches") pins that stream.  numpy spends ~10 ns per draw there; at BASELINE configs[4] every rank of an
// 8-GPU evaluation replays the draws of all ranks (33.5 M draws), which made the host the slowest part of the feed phase.
// This is the same arithmetic (MT19937 + tempering, smallest-bit-mask rejection on 32-bit outputs: numpy
// random/src/mt19937/mt19937.h, random/src/distributions/distributions.c: buffered_bounded_masked_uint32) in a tight loop,
// operating on the (key, pos) pair that np.random.get_state() returns and np.random.set_state() takes back.
#include <cstdint>

#include "common.cuh"

#include <cstdlib>

namespace bgan {
// hostrng_avx2.cpp (g++ -mavx2): the same draws, eight MT19937 words at a time
bool cpu_has_avx2();
void randint_fill_avx2(uint32_t* key, int* pos_inout, uint32_t mask, uint32_t bound, int32_t lo32, int64_t count, int32_t* out);
// hostrng_avx512.cpp (g++ -mavx512f): sixteen words at a time
bool cpu_has_avx512f();
void randint_fill_avx512(uint32_t* key, int* pos_inout, uint32_t mask, uint32_t bound, int32_t lo32, int64_t count, int32_t* out);
}  // namespace bgan

namespace {

constexpr int kN = 624, kM = 397;
constexpr uint32_t kMatrixA = 0x9908b0dfu, kUpper = 0x80000000u, kLower = 0x7fffffffu;

inline void mt_regenerate(uint32_t* key) {
    int i = 0;
    for (; i < kN - kM; ++i) {
        const uint32_t y = (key[i] & kUpper) | (key[i + 1] & kLower);
        key[i] = key[i + kM] ^ (y >> 1) ^ ((0u - (y & 1u)) & kMatrixA);
    }
    for (; i < kN - 1; ++i) {
        const uint32_t y = (key[i] & kUpper) | (key[i + 1] & kLower);
        key[i] = key[i + (kM - kN)] ^ (y >> 1) ^ ((0u - (y & 1u)) & kMatrixA);
    }
    const uint32_t y = (key[kN - 1] & kUpper) | (key[0] & kLower);
    key[kN - 1] = key[kM - 1] ^ (y >> 1) ^ ((0u - (y & 1u)) & kMatrixA);
}

inline uint32_t temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

// One bounded draw sequence on the state; out may be NULL (advance only).  Returns false for unsupported ranges.
// Works block-wise: the tempering + masking of a whole 624-word block vectorises, the rejection is a compaction.
bool randint_fill(uint32_t* key, int* pos_inout, int64_t low, int64_t high, int64_t count, int32_t* out) {
    const uint64_t rng = static_cast<uint64_t>(high - 1 - low);
    if (!(high > low) || rng >= 0xffffffffull || low < INT32_MIN || high - 1 > INT32_MAX) return false;
    if (rng == 0) {  // numpy fills with `low` and consumes nothing
        if (out)
            for (int64_t i = 0; i < count; ++i) out[i] = static_cast<int32_t>(low);
        return true;
    }
    uint32_t mask = static_cast<uint32_t>(rng);
    mask |= mask >> 1;
    mask |= mask >> 2;
    mask |= mask >> 4;
    mask |= mask >> 8;
    mask |= mask >> 16;
    const uint32_t bound = static_cast<uint32_t>(rng);
    const int32_t lo32 = static_cast<int32_t>(low);
    // widest body the CPU (and the OS) supports; BGAN_RNG_SCALAR=1 / BGAN_RNG_NO_AVX512=1 pin a narrower one (tests)
    static const bool scalar_only = getenv("BGAN_RNG_SCALAR") && atoi(getenv("BGAN_RNG_SCALAR"));
    static const bool use_avx512 = !scalar_only && bgan::cpu_has_avx512f() && !(getenv("BGAN_RNG_NO_AVX512") && atoi(getenv("BGAN_RNG_NO_AVX512")));
    static const bool use_avx2 = !scalar_only && bgan::cpu_has_avx2();
    if (use_avx512) {
        bgan::randint_fill_avx512(key, pos_inout, mask, bound, lo32, count, out);
        return true;
    }
    if (use_avx2) {
        bgan::randint_fill_avx2(key, pos_inout, mask, bound, lo32, count, out);
        return true;
    }
    int pos = *pos_inout;
    int64_t done = 0;
    uint32_t v[kN];
    while (done < count) {
        if (pos == kN) {
            mt_regenerate(key);
            pos = 0;
        }
        const int avail = kN - pos;
        for (int i = 0; i < avail; ++i) v[i] = temper(key[pos + i]) & mask;
        int used = 0;
        if (count - done >= avail) {  // the whole rest of the block is consumed: no early exit to test for
            if (out) {
                for (; used < avail; ++used) {
                    out[done] = lo32 + static_cast<int32_t>(v[used]);
                    done += (v[used] <= bound);
                }
            } else {
                for (; used < avail; ++used) done += (v[used] <= bound);
            }
        } else {
            for (; used < avail && done < count; ++used) {
                if (out) out[done] = lo32 + static_cast<int32_t>(v[used]);
                done += (v[used] <= bound);
            }
        }
        pos += used;
    }
    *pos_inout = pos;
    return true;
}

}  // namespace

extern "C" int bgan_mt19937_randint_i32(uint32_t* key624, int* pos_inout, int64_t low, int64_t high, int64_t count,
                                        int32_t* out_host) {
    BGAN_REQUIRE(key624 && pos_inout && count >= 0, BGAN_E_BADARG, "bgan_mt19937_randint_i32: bad arguments");
    BGAN_REQUIRE(high > low, BGAN_E_BADARG, "bgan_mt19937_randint_i32: low >= high");
    BGAN_REQUIRE(*pos_inout >= 0 && *pos_inout <= kN, BGAN_E_BADARG, "bgan_mt19937_randint_i32: pos outside [0, 624]");
    BGAN_REQUIRE(randint_fill(key624, pos_inout, low, high, count, out_host), BGAN_E_UNSUPPORTED,
                 "bgan_mt19937_randint_i32: only ranges that numpy serves with 32-bit masked rejection and that fit int32");
    return 0;
}

extern "C" int bgan_swd_positions_i32(uint32_t* key624, int* pos_inout, int levels, const int* heights, const int* widths,
                                      int nhood, int64_t count, int world, int rank, int32_t* out_host) {
    BGAN_REQUIRE(key624 && pos_inout && heights && widths && out_host && levels > 0 && count >= 0 && nhood > 0, BGAN_E_BADARG,
                 "bgan_swd_positions_i32: bad arguments");
    BGAN_REQUIRE(world >= 1 && rank >= 0 && rank < world, BGAN_E_BADARG, "bgan_swd_positions_i32: bad rank / world");
    BGAN_REQUIRE(*pos_inout >= 0 && *pos_inout <= kN, BGAN_E_BADARG, "bgan_swd_positions_i32: pos outside [0, 624]");
    const int half = nhood / 2;
    for (int l = 0; l < levels; ++l)
        BGAN_REQUIRE(widths[l] - half > half && heights[l] - half > half, BGAN_E_BADARG,
                     "bgan_swd_positions_i32: level %d (%d x %d) is smaller than the neighbourhood", l, heights[l], widths[l]);
    for (int r = 0; r < world; ++r)
        for (int l = 0; l < levels; ++l) {
            int32_t* ox = r == rank ? out_host + static_cast<int64_t>(2 * l) * count : nullptr;
            int32_t* oy = r == rank ? out_host + static_cast<int64_t>(2 * l + 1) * count : nullptr;
            // the order of sliced_wasserstein.py:20-21: columns first, then rows
            BGAN_REQUIRE(randint_fill(key624, pos_inout, half, widths[l] - half, count, ox), BGAN_E_UNSUPPORTED, "range");
            BGAN_REQUIRE(randint_fill(key624, pos_inout, half, heights[l] - half, count, oy), BGAN_E_UNSUPPORTED, "range");
        }
    return 0;
}
