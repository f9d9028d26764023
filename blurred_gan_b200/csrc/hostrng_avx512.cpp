// AVX-512 body of the numpy randint replay: hostrng_avx2.cpp with sixteen MT19937 words per step (see there and
// hostrng.cu, which picks the widest body the CPU and the OS support at run time).  Host-only, g++ -mavx512f.
// Accepted lanes are compressed in a register (vpcompressd zmm, zmm) and written with a full-width store: the
// memory-destination form of the instruction is microcoded on some CPUs.
#include <immintrin.h>

#include <cstdint>

namespace bgan {

namespace {

constexpr int kN = 624, kM = 397;
constexpr uint32_t kMatrixA = 0x9908b0dfu, kUpper = 0x80000000u, kLower = 0x7fffffffu;

inline __m512i twist16(const uint32_t* a, const uint32_t* far) {
    const __m512i cur = _mm512_loadu_si512(a);
    const __m512i nxt = _mm512_loadu_si512(a + 1);
    const __m512i y = _mm512_or_si512(_mm512_and_si512(cur, _mm512_set1_epi32(static_cast<int>(kUpper))),
                                      _mm512_and_si512(nxt, _mm512_set1_epi32(static_cast<int>(kLower))));
    const __mmask16 odd = _mm512_test_epi32_mask(y, _mm512_set1_epi32(1));
    const __m512i mag = _mm512_maskz_set1_epi32(odd, static_cast<int>(kMatrixA));
    return _mm512_xor_si512(_mm512_xor_si512(_mm512_loadu_si512(far), _mm512_srli_epi32(y, 1)), mag);
}

inline uint32_t twist1(uint32_t cur, uint32_t nxt, uint32_t far) {
    const uint32_t y = (cur & kUpper) | (nxt & kLower);
    return far ^ (y >> 1) ^ ((0u - (y & 1u)) & kMatrixA);
}

void regenerate(uint32_t* key) {
    int i = 0;
    // words [0, 227): partner key[i + 397] is an old word; words [i, i + 16) are stored after word i + 16 has been loaded
    for (; i + 16 <= kN - kM; i += 16) _mm512_storeu_si512(key + i, twist16(key + i, key + i + kM));
    for (; i < kN - kM; ++i) key[i] = twist1(key[i], key[i + 1], key[i + kM]);
    // words [227, 623): partner key[i - 227] is a new word, written at least 211 words ago
    for (; i + 16 <= kN - 1; i += 16) _mm512_storeu_si512(key + i, twist16(key + i, key + i + (kM - kN)));
    for (; i < kN - 1; ++i) key[i] = twist1(key[i], key[i + 1], key[i + (kM - kN)]);
    key[kN - 1] = twist1(key[kN - 1], key[0], key[kM - 1]);
}

inline __m512i temper16(__m512i y) {
    y = _mm512_xor_si512(y, _mm512_srli_epi32(y, 11));
    y = _mm512_xor_si512(y, _mm512_and_si512(_mm512_slli_epi32(y, 7), _mm512_set1_epi32(static_cast<int>(0x9d2c5680u))));
    y = _mm512_xor_si512(y, _mm512_and_si512(_mm512_slli_epi32(y, 15), _mm512_set1_epi32(static_cast<int>(0xefc60000u))));
    return _mm512_xor_si512(y, _mm512_srli_epi32(y, 18));
}
inline uint32_t temper1(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

}  // namespace

// (trivial scalar code: safe to execute on a CPU without AVX-512 although this file is compiled with -mavx512f)
bool cpu_has_avx512f() { return __builtin_cpu_supports("avx512f") != 0; }

// Same contract as randint_fill_avx2.
void randint_fill_avx512(uint32_t* key, int* pos_inout, uint32_t mask, uint32_t bound, int32_t lo32, int64_t count, int32_t* out) {
    const __m512i vmask = _mm512_set1_epi32(static_cast<int>(mask));
    const __m512i vbound = _mm512_set1_epi32(static_cast<int>(bound));
    const __m512i vlo = _mm512_set1_epi32(lo32);
    int pos = *pos_inout;
    int64_t done = 0;
    while (done < count) {
        if (pos == kN) {
            regenerate(key);
            pos = 0;
        }
        // sixteen words at a time while a whole group is available and cannot overshoot `count` (all sixteen may be accepted):
        // the full-width store of the compressed lanes then stays inside out[done .. done + 16)
        while (pos + 16 <= kN && count - done >= 16) {
            const __m512i v = _mm512_and_si512(temper16(_mm512_loadu_si512(key + pos)), vmask);
            const __mmask16 ok = _mm512_cmple_epu32_mask(v, vbound);
            if (out) _mm512_storeu_si512(out + done, _mm512_maskz_compress_epi32(ok, _mm512_add_epi32(v, vlo)));
            done += __builtin_popcount(static_cast<unsigned>(ok));
            pos += 16;
        }
        // the rest of the block, or the last few draws, one word at a time
        while (pos < kN && done < count && (pos + 16 > kN || count - done < 16)) {
            const uint32_t v = temper1(key[pos++]) & mask;
            if (v <= bound) {
                if (out) out[done] = lo32 + static_cast<int32_t>(v);
                ++done;
            }
        }
    }
    *pos_inout = pos;
}

}  // namespace bgan
