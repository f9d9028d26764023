// AVX2 body of the numpy randint replay (see hostrng.cu, which selects it at run time when the CPU has AVX2).
// Host-only, compiled by g++ with -mavx2; same arithmetic as the scalar path, eight MT19937 words at a time:
//   regeneration  key[i] = key[i + 397] ^ (y >> 1) ^ (y & 1 ? 0x9908b0df : 0),  y = (key[i] & 0x80000000) | (key[i + 1] & 0x7fffffff)
//                 (the recurrence reaches back 227 words, so eight consecutive words are independent)
//   tempering, masking, the rejection test, and a left-pack of the accepted lanes (permutation table) for the store.
// Every rank of a multi-rank SWD evaluation replays the draws of all ranks (only its own are stored), which made this
// loop the largest single item of an 8-GPU evaluation.
#include <immintrin.h>

#include <cstdint>

namespace bgan {

namespace {

constexpr int kN = 624, kM = 397;
constexpr uint32_t kMatrixA = 0x9908b0dfu, kUpper = 0x80000000u, kLower = 0x7fffffffu;

inline __m256i twist8(const uint32_t* a, const uint32_t* far) {
    const __m256i cur = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(a));
    const __m256i nxt = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(a + 1));
    const __m256i y = _mm256_or_si256(_mm256_and_si256(cur, _mm256_set1_epi32(static_cast<int>(kUpper))),
                                      _mm256_and_si256(nxt, _mm256_set1_epi32(static_cast<int>(kLower))));
    const __m256i odd = _mm256_sub_epi32(_mm256_setzero_si256(), _mm256_and_si256(y, _mm256_set1_epi32(1)));  // 0 or ~0
    const __m256i mag = _mm256_and_si256(odd, _mm256_set1_epi32(static_cast<int>(kMatrixA)));
    return _mm256_xor_si256(_mm256_xor_si256(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(far)), _mm256_srli_epi32(y, 1)), mag);
}

inline uint32_t twist1(uint32_t cur, uint32_t nxt, uint32_t far) {
    const uint32_t y = (cur & kUpper) | (nxt & kLower);
    return far ^ (y >> 1) ^ ((0u - (y & 1u)) & kMatrixA);
}

void regenerate(uint32_t* key) {
    int i = 0;
    // words [0, 227): partner key[i + 397] is an old word; the store of words [i, i + 8) comes after the load of word i + 8
    for (; i + 8 <= kN - kM; i += 8) _mm256_storeu_si256(reinterpret_cast<__m256i*>(key + i), twist8(key + i, key + i + kM));
    for (; i < kN - kM; ++i) key[i] = twist1(key[i], key[i + 1], key[i + kM]);
    // words [227, 623): partner key[i - 227] is a new word, written at least 227 words ago
    for (; i + 8 <= kN - 1; i += 8) _mm256_storeu_si256(reinterpret_cast<__m256i*>(key + i), twist8(key + i, key + i + (kM - kN)));
    for (; i < kN - 1; ++i) key[i] = twist1(key[i], key[i + 1], key[i + (kM - kN)]);
    key[kN - 1] = twist1(key[kN - 1], key[0], key[kM - 1]);
}

inline __m256i temper8(__m256i y) {
    y = _mm256_xor_si256(y, _mm256_srli_epi32(y, 11));
    y = _mm256_xor_si256(y, _mm256_and_si256(_mm256_slli_epi32(y, 7), _mm256_set1_epi32(static_cast<int>(0x9d2c5680u))));
    y = _mm256_xor_si256(y, _mm256_and_si256(_mm256_slli_epi32(y, 15), _mm256_set1_epi32(static_cast<int>(0xefc60000u))));
    return _mm256_xor_si256(y, _mm256_srli_epi32(y, 18));
}
inline uint32_t temper1(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

// left-pack table: for an 8-bit accept mask, the lane indices of the accepted lanes in order
struct PackTable {
    alignas(32) uint32_t idx[256][8];
    PackTable() {
        for (int m = 0; m < 256; ++m) {
            int k = 0;
            for (int l = 0; l < 8; ++l)
                if (m & (1 << l)) idx[m][k++] = static_cast<uint32_t>(l);
            for (; k < 8; ++k) idx[m][k] = 0;
        }
    }
};

}  // namespace

// (trivial scalar code: safe to execute on a CPU without AVX2 although this file is compiled with -mavx2)
bool cpu_has_avx2() { return __builtin_cpu_supports("avx2") != 0; }

// Same contract as randint_fill in hostrng.cu (range already validated there: 0 < rng < 2^32 - 1, results fit int32).
void randint_fill_avx2(uint32_t* key, int* pos_inout, uint32_t mask, uint32_t bound, int32_t lo32, int64_t count, int32_t* out) {
    static const PackTable table;
    const __m256i vmask = _mm256_set1_epi32(static_cast<int>(mask));
    const __m256i vbound = _mm256_set1_epi32(static_cast<int>(bound));
    const __m256i vlo = _mm256_set1_epi32(lo32);
    int pos = *pos_inout;
    int64_t done = 0;
    while (done < count) {
        if (pos == kN) {
            regenerate(key);
            pos = 0;
        }
        // eight words at a time while a whole group is available and cannot overshoot `count` (all eight may be accepted);
        // out[done .. done + 8) is then inside the caller's buffer
        while (pos + 8 <= kN && count - done >= 8) {
            const __m256i v = _mm256_and_si256(temper8(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(key + pos))), vmask);
            const __m256i ok = _mm256_cmpeq_epi32(_mm256_min_epu32(v, vbound), v);  // v <= bound, unsigned
            const int m = _mm256_movemask_ps(_mm256_castsi256_ps(ok));
            if (out) {
                const __m256i packed = _mm256_permutevar8x32_epi32(_mm256_add_epi32(v, vlo),
                                                                   _mm256_load_si256(reinterpret_cast<const __m256i*>(table.idx[m])));
                _mm256_storeu_si256(reinterpret_cast<__m256i*>(out + done), packed);
            }
            done += __builtin_popcount(static_cast<unsigned>(m));
            pos += 8;
        }
        // the rest of the block, or the last few draws, one word at a time
        while (pos < kN && done < count && (pos + 8 > kN || count - done < 8)) {
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
