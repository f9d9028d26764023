// Segmented sort of fp32 keys by splitters + shared-memory buckets (sm_100a) -- the np.sort(proj, axis=0) of the reference's
// sliced_wasserstein.py:47-48 for segments of up to 2^21 keys, and its fusion with the mean |A - B| of :49-50.
//
// The LSD radix sort of sort.cu moves every key through HBM four times and ranks it four times with eight ballots (it is
// instruction-bound: profiles/r02_sort.md).  Here a key crosses HBM twice and is ranked with shared-memory atomics:
//
//   splitters  one CTA per segment: 16 x B stratified samples, bitonic sort in shared memory, every 16th sample is a splitter.
//              B range buckets (mean ~2048 keys, whatever the distribution: the mass between order statistics of a sample
//              does not depend on it) + B "equality" buckets (keys equal to a splitter: a value that occurs so often that it
//              is sampled twice gets a bucket of its own that needs no sorting).
//   classify   per block of 16384 keys: bucket id by binary search over the splitters (shared memory) -> ids (uint16),
//              per-block bucket counts (shared-memory atomics) -> counts[seg][block][id]
//   scan       per segment: bucket starts, and the start of every (block, id) run -> bases[seg][block][id]; flags a range
//              bucket that does not fit a CTA (never seen; the caller then falls back to the radix sort)
//   scatter    per block: keys regrouped by id in shared memory (slot = local start + shared-memory atomic), written as
//              runs to scratch.  Not stable -- the last step sorts the buckets anyway.
//   local sort persistent CTAs take runs of adjacent buckets (<= 8192 keys): sub-bucket = linear map of the value between the
//              run's bounding splitters (~2 keys per sub-bucket), count with atomics, scan, place, rank the few keys of a
//              sub-bucket against each other, write the sorted run (or, fused: meet the sorted A and accumulate |a - b|).
//              A run whose sub-buckets are badly loaded (clustered values, non-finite bounds) is sorted by a bitonic network
//              in shared memory instead: slower, never wrong.
//
// HBM traffic per key: 4 (classify) + 2 + 2 (ids) + 4 + 4 (scatter) + 4 + 4 (local sort) = 24 bytes + ~3 of tables, against 48.
// Segments of <= 8192 keys skip everything but the last step.
#include <algorithm>
#include <cstdlib>

#include "sort.cuh"

namespace bgan {
namespace {

constexpr int kSsOversample = 16;
constexpr int kSsMeanBucket = 2048;
constexpr int kSsMaxBuckets = 1024;
constexpr int kSsCap = 8192;           // keys a local-sort CTA holds
constexpr int kSsBlock = 16384;        // keys per classify / scatter CTA
constexpr int kSsThreads = 256;        // classify / scatter
constexpr int64_t kSsMaxRows = int64_t(1) << 21;
constexpr int kSplThreads = 1024;      // splitters / scan kernels
constexpr int kLsThreads = 512;        // local sort
constexpr int kLsWarps = kLsThreads / 32;
constexpr int kLsItems = kSsCap / kLsThreads;  // 16 keys per thread
constexpr int kLsMaxSub = 4096;
constexpr int kLsBatchIds = 16;        // bucket ids per ticket
constexpr int kLsEqPass = 64;          // an equality bucket larger than this is passed through on its own

struct SsPlan {
    int buckets, ids, blocks, samples;
    bool direct;
};

SsPlan ss_plan(int64_t rows) {
    SsPlan p;
    p.direct = rows <= kSsCap;
    const int64_t want = ceil_div64(rows, kSsMeanBucket);
    int b = 2;
    while (b < want && b < kSsMaxBuckets) b <<= 1;
    p.buckets = b;
    p.ids = 2 * b;
    p.blocks = static_cast<int>(ceil_div64(rows, kSsBlock));
    p.samples = b * kSsOversample;
    return p;
}

size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

struct SsLayout {
    size_t scratch, ids, counts, bases, starts, bounds, ctrl, total;
};
SsLayout ss_layout(int64_t rows, int segments) {
    const SsPlan p = ss_plan(rows);
    SsLayout l;
    size_t at = 0;
    auto take = [&](size_t bytes) {
        const size_t here = at;
        at += align256(bytes);
        return here;
    };
    const size_t n = static_cast<size_t>(rows) * segments;
    const size_t tab = static_cast<size_t>(segments) * p.blocks * p.ids;
    l.ctrl = take(256);
    if (p.direct) {
        l.scratch = l.ids = l.counts = l.bases = l.starts = l.bounds = 0;
    } else {
        l.scratch = take(n * sizeof(uint32_t));
        l.ids = take(n * sizeof(uint16_t));
        l.counts = take(tab * sizeof(uint16_t));
        l.bases = take(tab * sizeof(uint32_t));
        l.starts = take(static_cast<size_t>(segments) * (p.ids + 1) * sizeof(uint32_t));
        l.bounds = take(static_cast<size_t>(segments) * (p.buckets + 1) * sizeof(uint32_t));
    }
    l.total = at;
    return l;
}

// ---- bitonic network over n (a power of two) ordered keys in shared memory; ends with a barrier --------------------------
template <int THREADS>
__device__ __forceinline__ void bitonic_sort_smem(uint32_t* s, int n, int tid) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (n >> 1); t += THREADS) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int p = i | j;
                const bool asc = (i & k) == 0;
                const uint32_t a = s[i], b = s[p];
                if ((a > b) == asc) {
                    s[i] = b;
                    s[p] = a;
                }
            }
            __syncthreads();
        }
    }
}

// ---- splitters ---------------------------------------------------------------------------------------------------------------
// bounds[seg][0] = smallest sample, [1 .. B-1] = the B-1 splitters, [B] = largest sample (ordered-uint images)
__global__ void __launch_bounds__(kSplThreads) ss_splitters_kernel(const uint32_t* __restrict__ keys, int64_t rows, int buckets,
                                                                   uint32_t* __restrict__ bounds) {
    extern __shared__ uint32_t s_sample[];
    const int seg = blockIdx.x, tid = threadIdx.x;
    const int n = buckets * kSsOversample;
    const uint32_t* src = keys + static_cast<size_t>(seg) * rows;
    for (int i = tid; i < n; i += kSplThreads) {
        // one sample from each of n equal strata, at a hashed position inside the stratum
        const int64_t lo = static_cast<int64_t>(i) * rows / n, hi = static_cast<int64_t>(i + 1) * rows / n;
        uint32_t h = static_cast<uint32_t>(i) * 0x9E3779B1u + static_cast<uint32_t>(seg) * 0x85EBCA77u;
        h ^= h >> 15;
        h *= 0x2C1B3C6Du;
        h ^= h >> 12;
        const uint32_t width = static_cast<uint32_t>(hi - lo);
        s_sample[i] = float_to_ordered(__ldg(src + lo + (width > 1 ? h % width : 0u)));
    }
    __syncthreads();
    bitonic_sort_smem<kSplThreads>(s_sample, n, tid);
    uint32_t* out = bounds + static_cast<size_t>(seg) * (buckets + 1);
    for (int j = tid; j <= buckets; j += kSplThreads)
        out[j] = j == 0 ? s_sample[0] : (j == buckets ? s_sample[n - 1] : s_sample[j * kSsOversample - 1]);
}

// ---- classify ----------------------------------------------------------------------------------------------------------------
// id = 2 j for a key strictly between splitter j-1 and splitter j, 2 j + 1 for a key equal to splitter j: ids order like keys
__device__ __forceinline__ uint32_t bucket_id(const uint32_t* sp, int buckets, uint32_t u) {
    int j = 0;
    for (int step = buckets >> 1; step > 0; step >>= 1) j += (sp[j + step - 1] < u) ? step : 0;
    const bool eq = j < buckets - 1 && sp[j] == u;
    return 2u * j + (eq ? 1u : 0u);
}

__global__ void __launch_bounds__(kSsThreads) ss_classify_kernel(const uint32_t* __restrict__ keys, int64_t rows, int buckets, int blocks,
                                                                 const uint32_t* __restrict__ bounds, uint16_t* __restrict__ ids,
                                                                 uint16_t* __restrict__ counts) {
    extern __shared__ uint32_t s_cls[];  // splitters [buckets] | hist [2 buckets]
    uint32_t* sp = s_cls;
    uint32_t* hist = s_cls + buckets;
    const int blk = blockIdx.x, seg = blockIdx.y, tid = threadIdx.x;
    const uint32_t* b = bounds + static_cast<size_t>(seg) * (buckets + 1);
    for (int j = tid; j < buckets; j += kSsThreads) sp[j] = j < buckets - 1 ? b[j + 1] : 0xffffffffu;
    for (int j = tid; j < 2 * buckets; j += kSsThreads) hist[j] = 0;
    __syncthreads();
    const int64_t base = static_cast<int64_t>(blk) * kSsBlock;
    const int valid = rows - base < kSsBlock ? static_cast<int>(rows - base) : kSsBlock;
    const size_t at = static_cast<size_t>(seg) * rows + base;
    const uint32_t* src = keys + at;
    uint16_t* dst = ids + at;
    if ((at & 3) == 0) {
        const int vec = valid >> 2;
#pragma unroll 4
        for (int v = tid; v < vec; v += kSsThreads) {
            const uint4 k = __ldg(reinterpret_cast<const uint4*>(src) + v);
            const uint32_t i0 = bucket_id(sp, buckets, float_to_ordered(k.x)), i1 = bucket_id(sp, buckets, float_to_ordered(k.y));
            const uint32_t i2 = bucket_id(sp, buckets, float_to_ordered(k.z)), i3 = bucket_id(sp, buckets, float_to_ordered(k.w));
            atomicAdd(&hist[i0], 1u);
            atomicAdd(&hist[i1], 1u);
            atomicAdd(&hist[i2], 1u);
            atomicAdd(&hist[i3], 1u);
            reinterpret_cast<uint2*>(dst)[v] = make_uint2(i0 | (i1 << 16), i2 | (i3 << 16));
        }
        for (int i = (vec << 2) + tid; i < valid; i += kSsThreads) {
            const uint32_t id = bucket_id(sp, buckets, float_to_ordered(__ldg(src + i)));
            atomicAdd(&hist[id], 1u);
            dst[i] = static_cast<uint16_t>(id);
        }
    } else {
        for (int i = tid; i < valid; i += kSsThreads) {
            const uint32_t id = bucket_id(sp, buckets, float_to_ordered(__ldg(src + i)));
            atomicAdd(&hist[id], 1u);
            dst[i] = static_cast<uint16_t>(id);
        }
    }
    __syncthreads();
    uint16_t* c = counts + (static_cast<size_t>(seg) * blocks + blk) * (2 * buckets);
    for (int j = tid; j < 2 * buckets; j += kSsThreads) c[j] = static_cast<uint16_t>(hist[j]);
}

// ---- scan ----------------------------------------------------------------------------------------------------------------------
// starts[seg][id] (ids + 1 entries) and bases[seg][block][id] = starts[id] + keys of that id in earlier blocks
__global__ void __launch_bounds__(kSplThreads) ss_scan_kernel(const uint16_t* __restrict__ counts, int64_t rows, int ids, int blocks,
                                                              uint32_t* __restrict__ starts, uint32_t* __restrict__ bases,
                                                              uint32_t* __restrict__ ctrl) {
    __shared__ uint32_t s_tot[2 * kSsMaxBuckets];
    __shared__ uint32_t s_warp[kSplThreads / 32];
    const int seg = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint16_t* c = counts + static_cast<size_t>(seg) * blocks * ids;
    for (int id = tid; id < 2 * kSsMaxBuckets; id += kSplThreads) {
        uint32_t t = 0;
        if (id < ids) {
#pragma unroll 8
            for (int b = 0; b < blocks; ++b) t += c[static_cast<size_t>(b) * ids + id];
            if ((id & 1) == 0 && t > static_cast<uint32_t>(kSsCap)) atomicOr(ctrl, 1u);  // a range bucket that no CTA can hold
        }
        s_tot[id] = t;
    }
    __syncthreads();
    // exclusive scan over the ids: thread t owns ids 2 t and 2 t + 1
    const uint32_t a = s_tot[2 * tid], b2 = s_tot[2 * tid + 1];
    uint32_t incl = a + b2;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t excl = incl - (a + b2);
    for (int w = 0; w < warp; ++w) excl += s_warp[w];
    __syncthreads();
    s_tot[2 * tid] = excl;
    s_tot[2 * tid + 1] = excl + a;
    __syncthreads();
    uint32_t* st = starts + static_cast<size_t>(seg) * (ids + 1);
    for (int id = tid; id <= ids; id += kSplThreads) st[id] = id < ids ? s_tot[id] : static_cast<uint32_t>(rows);
    uint32_t* bs = bases + static_cast<size_t>(seg) * blocks * ids;
    for (int id = tid; id < ids; id += kSplThreads) {
        uint32_t run = s_tot[id];
#pragma unroll 8
        for (int b = 0; b < blocks; ++b) {
            bs[static_cast<size_t>(b) * ids + id] = run;
            run += c[static_cast<size_t>(b) * ids + id];
        }
    }
}

// ---- scatter -------------------------------------------------------------------------------------------------------------------
constexpr size_t kScatterSmem = kSsBlock * sizeof(uint32_t) + kSsBlock * sizeof(uint16_t) + 2 * (2 * kSsMaxBuckets) * sizeof(uint32_t) + 64;

__global__ void __launch_bounds__(kSsThreads, 2) ss_scatter_kernel(const uint32_t* __restrict__ keys, const uint16_t* __restrict__ ids,
                                                                   int64_t rows, int n_ids, int blocks, const uint16_t* __restrict__ counts,
                                                                   const uint32_t* __restrict__ bases, uint32_t* __restrict__ out,
                                                                   const uint32_t* __restrict__ ctrl) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    uint32_t* s_key = reinterpret_cast<uint32_t*>(s_raw);
    uint16_t* s_id = reinterpret_cast<uint16_t*>(s_raw + kSsBlock * sizeof(uint32_t));
    uint32_t* s_cur = reinterpret_cast<uint32_t*>(s_raw + kSsBlock * (sizeof(uint32_t) + sizeof(uint16_t)));
    uint32_t* s_delta = s_cur + 2 * kSsMaxBuckets;
    uint32_t* s_warp = s_delta + 2 * kSsMaxBuckets;
    if (*ctrl != 0) return;
    const int blk = blockIdx.x, seg = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t row = (static_cast<size_t>(seg) * blocks + blk) * n_ids;
    // local starts: exclusive scan of this block's counts; thread t owns `per` consecutive ids
    const int per = (n_ids + kSsThreads - 1) / kSsThreads;  // 1 .. 8
    uint32_t c[8];
    uint32_t sum = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int id = tid * per + e;
        c[e] = (e < per && id < n_ids) ? counts[row + id] : 0u;
        sum += c[e];
    }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t run = incl - sum;
    for (int w = 0; w < warp; ++w) run += s_warp[w];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int id = tid * per + e;
        if (e < per && id < n_ids) {
            s_cur[id] = run;
            s_delta[id] = bases[row + id] - run;
            run += c[e];
        }
    }
    __syncthreads();
    const int64_t base = static_cast<int64_t>(blk) * kSsBlock;
    const int valid = rows - base < kSsBlock ? static_cast<int>(rows - base) : kSsBlock;
    const size_t at = static_cast<size_t>(seg) * rows + base;
    const uint32_t* src = keys + at;
    const uint16_t* sid = ids + at;
    if ((at & 3) == 0) {
        const int vec = valid >> 2;
#pragma unroll 4
        for (int v = tid; v < vec; v += kSsThreads) {
            const uint4 k = __ldg(reinterpret_cast<const uint4*>(src) + v);
            const uint2 i2 = __ldg(reinterpret_cast<const uint2*>(sid) + v);
            const uint32_t id0 = i2.x & 0xffff, id1 = i2.x >> 16, id2 = i2.y & 0xffff, id3 = i2.y >> 16;
            const uint32_t p0 = atomicAdd(&s_cur[id0], 1u), p1 = atomicAdd(&s_cur[id1], 1u);
            const uint32_t p2 = atomicAdd(&s_cur[id2], 1u), p3 = atomicAdd(&s_cur[id3], 1u);
            s_key[p0] = k.x, s_id[p0] = static_cast<uint16_t>(id0);
            s_key[p1] = k.y, s_id[p1] = static_cast<uint16_t>(id1);
            s_key[p2] = k.z, s_id[p2] = static_cast<uint16_t>(id2);
            s_key[p3] = k.w, s_id[p3] = static_cast<uint16_t>(id3);
        }
        for (int i = (vec << 2) + tid; i < valid; i += kSsThreads) {
            const uint32_t id = sid[i], p = atomicAdd(&s_cur[id], 1u);
            s_key[p] = __ldg(src + i), s_id[p] = static_cast<uint16_t>(id);
        }
    } else {
        for (int i = tid; i < valid; i += kSsThreads) {
            const uint32_t id = sid[i], p = atomicAdd(&s_cur[id], 1u);
            s_key[p] = __ldg(src + i), s_id[p] = static_cast<uint16_t>(id);
        }
    }
    __syncthreads();
    uint32_t* dst = out + static_cast<size_t>(seg) * rows;
#pragma unroll 4
    for (int slot = tid; slot < valid; slot += kSsThreads) dst[s_delta[s_id[slot]] + slot] = s_key[slot];
}

// ---- local sort ------------------------------------------------------------------------------------------------------------------
struct LocalSmem {
    uint32_t buf[kSsCap];
    uint32_t fin[kSsCap];
    uint32_t cnt[kLsMaxSub + 32];
    uint32_t starts[kLsBatchIds + 1];
    uint32_t warp[kLsWarps];
    uint32_t red[2][kLsWarps];
    double dsum[kLsWarps];
    uint32_t ticket;
    uint32_t cost;
};

struct LocalArgs {
    const uint32_t* src;     // bucketed keys (scratch), or the keys themselves when a segment is one run (direct)
    uint32_t* dst;           // sorted floats (unused with da.ref)
    int64_t rows;
    int segments;
    int n_ids;               // 0: direct
    int buckets;
    const uint32_t* starts;  // [seg][n_ids + 1]
    const uint32_t* bounds;  // [seg][buckets + 1]
    uint32_t* ctrl;          // [0] overflow flag, [1] ticket counter
    DiffArgs da;
};

template <bool DIFF>
__global__ void __launch_bounds__(kLsThreads, 2) ss_local_sort_kernel(const LocalArgs a) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    LocalSmem& sm = *reinterpret_cast<LocalSmem*>(s_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool direct = a.n_ids == 0;
    if (!direct && a.ctrl[0] != 0) return;
    const int batch_ids = direct ? 1 : min(kLsBatchIds, a.n_ids);
    const int batches_per_seg = direct ? 1 : a.n_ids / batch_ids;
    const uint32_t total = static_cast<uint32_t>(a.segments) * batches_per_seg;
    for (int i = tid; i < kLsMaxSub + 32; i += kLsThreads) sm.cnt[i] = 0;

    double acc = 0.0;
    int acc_sum = -1;
    auto flush = [&]() {  // (uniform) this CTA's |a - b| so far -> one of the partial accumulators of its sum
        const double t = warp_sum(acc);
        if (lane == 0) sm.dsum[warp] = t;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < kLsWarps; ++w) s += sm.dsum[w];
            atomicAdd(a.da.partial + static_cast<size_t>(acc_sum) * kDiffSlots + (blockIdx.x % kDiffSlots), s);
        }
        __syncthreads();
        acc = 0.0;
    };

    for (;;) {
        __syncthreads();
        if (tid == 0) sm.ticket = atomicAdd(a.ctrl + 1, 1u);
        __syncthreads();
        const uint32_t ticket = sm.ticket;
        if (ticket >= total) break;
        const int seg = static_cast<int>(ticket / batches_per_seg);
        const int id0 = static_cast<int>(ticket % batches_per_seg) * batch_ids;
        if (tid <= batch_ids)
            sm.starts[tid] = direct ? (tid == 0 ? 0u : static_cast<uint32_t>(a.rows))
                                    : a.starts[static_cast<size_t>(seg) * (a.n_ids + 1) + id0 + tid];
        __syncthreads();
        if constexpr (DIFF) {
            const int sum_idx = (seg + a.da.seg_offset) / a.da.segs_per_sum;
            if (sum_idx != acc_sum) {
                if (acc_sum >= 0) flush();
                acc_sum = sum_idx;
            }
        }
        const size_t seg_at = static_cast<size_t>(seg) * a.rows;
        int g0 = 0;
        while (g0 < batch_ids) {
            // a run of adjacent buckets that fits the CTA; a large equality bucket (all keys equal) goes through unsorted
            auto len = [&](int g) { return static_cast<int>(sm.starts[g + 1] - sm.starts[g]); };
            auto big_eq = [&](int g) { return !direct && ((id0 + g) & 1) && len(g) > kLsEqPass; };
            int g1 = g0 + 1;
            bool pass = big_eq(g0);
            if (!pass)
                while (g1 < batch_ids && !big_eq(g1) && static_cast<int>(sm.starts[g1 + 1] - sm.starts[g0]) <= kSsCap) ++g1;
            const uint32_t start = sm.starts[g0];
            const int n = static_cast<int>(sm.starts[g1] - start);
            const int first_id = id0 + g0, last_id = id0 + g1 - 1;
            g0 = g1;
            if (n == 0) continue;
            const uint32_t* src = a.src + seg_at + start;
            if (pass || n > kSsCap) {  // (n > kSsCap cannot happen: the scan kernel raised the flag and nothing runs)
                for (int i = tid; i < n; i += kLsThreads) {
                    const uint32_t v = __ldg(src + i);
                    if constexpr (DIFF)
                        acc += static_cast<double>(fabsf(__ldg(a.da.ref + seg_at + start + i) - __uint_as_float(v)));
                    else
                        a.dst[seg_at + start + i] = v;
                }
                continue;
            }

            // ---- keys -> registers ----
            uint32_t k[kLsItems];
            uint32_t mn = 0xffffffffu, mx = 0u;
#pragma unroll
            for (int q = 0; q < kLsItems; ++q) {
                const int i = q * kLsThreads + tid;
                k[q] = i < n ? float_to_ordered(__ldg(src + i)) : 0xffffffffu;
                if (i < n) mn = min(mn, k[q]), mx = max(mx, k[q]);
            }
            uint32_t lo_u, hi_u;
            if (direct) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                }
                if (lane == 0) sm.red[0][warp] = mn, sm.red[1][warp] = mx;
                __syncthreads();
                lo_u = sm.red[0][0], hi_u = sm.red[1][0];
#pragma unroll
                for (int w = 1; w < kLsWarps; ++w) lo_u = min(lo_u, sm.red[0][w]), hi_u = max(hi_u, sm.red[1][w]);
            } else {
                const uint32_t* b = a.bounds + static_cast<size_t>(seg) * (a.buckets + 1);
                lo_u = b[(first_id >> 1) + (first_id & 1)];
                hi_u = b[(last_id >> 1) + 1];
            }
            const float flo = __uint_as_float(ordered_to_float(lo_u)), fhi = __uint_as_float(ordered_to_float(hi_u));
            int n_sub = 64;
            while (n_sub < (n >> 1)) n_sub <<= 1;  // ~2 keys per sub-bucket, <= kLsMaxSub
            const float scale = static_cast<float>(n_sub) / (fhi - flo);
            bool network = !(isfinite(flo) && isfinite(fhi)) || n < 2;
            uint32_t pk[kLsItems];  // sub-bucket << 16 | arrival rank inside it
            if (!network) {
#pragma unroll
                for (int q = 0; q < kLsItems; ++q) {
                    if (q * kLsThreads + tid < n) {
                        // monotone in the key whatever flo and scale are: subtraction, product, truncation and clamp all are
                        const float x = __uint_as_float(ordered_to_float(k[q]));
                        int sub = min(n_sub - 1, max(0, __float2int_rz((x - flo) * scale)));
                        if (x != x) sub = (k[q] >> 31) ? n_sub - 1 : 0;  // NaNs sit at the ends of the ordered-uint range
                        pk[q] = (static_cast<uint32_t>(sub) << 16) | atomicAdd(&sm.cnt[sub], 1u);
                    }
                }
                if (tid == 0) sm.cost = 0;
                __syncthreads();
                // exclusive scan of the counts in place; sum of squares = the compares the ranking step will make
                const int per = n_sub >= kLsThreads ? n_sub / kLsThreads : 1;  // 1 .. 8
                uint32_t c[kLsMaxSub / kLsThreads];
                uint32_t sum = 0, sq = 0;
#pragma unroll
                for (int e = 0; e < kLsMaxSub / kLsThreads; ++e) {
                    const int sidx = tid * per + e;
                    c[e] = (e < per && sidx < n_sub) ? sm.cnt[sidx] : 0u;
                    sum += c[e];
                    sq += c[e] * c[e];
                }
                uint32_t incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
                if (lane == 31) sm.warp[warp] = incl;
                if (lane == 0 && sq) atomicAdd(&sm.cost, sq);
                __syncthreads();
                uint32_t run = incl - sum;
                for (int w = 0; w < warp; ++w) run += sm.warp[w];
#pragma unroll
                for (int e = 0; e < kLsMaxSub / kLsThreads; ++e) {
                    const int sidx = tid * per + e;
                    if (e < per && sidx < n_sub) {
                        sm.cnt[sidx] = run;
                        run += c[e];
                    }
                }
                if (tid == 0) sm.cnt[n_sub] = static_cast<uint32_t>(n);
                // badly loaded sub-buckets (clustered values): the all-pairs ranking would cost more than the network
                network = sm.cost > static_cast<uint32_t>(n) * 64u;
                __syncthreads();
            }
            const uint32_t* sorted;
            if (!network) {
#pragma unroll
                for (int q = 0; q < kLsItems; ++q)
                    if (q * kLsThreads + tid < n) sm.buf[sm.cnt[pk[q] >> 16] + (pk[q] & 0xffff)] = k[q];
                __syncthreads();
#pragma unroll
                for (int q = 0; q < kLsItems; ++q) {
                    if (q * kLsThreads + tid < n) {
                        const uint32_t sub = pk[q] >> 16;
                        const uint32_t s0 = sm.cnt[sub], s1 = sm.cnt[sub + 1], mine = s0 + (pk[q] & 0xffff);
                        uint32_t rank = s0;
                        for (uint32_t j = s0; j < s1; ++j) {
                            const uint32_t o = sm.buf[j];
                            rank += (o < k[q] || (o == k[q] && j < mine)) ? 1u : 0u;
                        }
                        sm.fin[rank] = k[q];
                    }
                }
                sorted = sm.fin;
            } else {
                int padded = 32;
                while (padded < n) padded <<= 1;
#pragma unroll
                for (int q = 0; q < kLsItems; ++q) {
                    const int i = q * kLsThreads + tid;
                    if (i < padded) sm.buf[i] = k[q];  // (k[q] = 0xffffffff beyond n: sorts last)
                }
                __syncthreads();
                bitonic_sort_smem<kLsThreads>(sm.buf, padded, tid);
                sorted = sm.buf;
            }
            __syncthreads();
            for (int i = tid; i <= n_sub; i += kLsThreads) sm.cnt[i] = 0;  // ready for the next run
            for (int i = tid; i < n; i += kLsThreads) {
                const uint32_t v = ordered_to_float(sorted[i]);
                if constexpr (DIFF)
                    acc += static_cast<double>(fabsf(__ldg(a.da.ref + seg_at + start + i) - __uint_as_float(v)));
                else
                    a.dst[seg_at + start + i] = v;
            }
            __syncthreads();
        }
    }
    if constexpr (DIFF) {
        if (acc_sum >= 0) flush();
    }
}

}  // namespace

bool sample_sort_applicable(int64_t rows) { return rows > 0 && rows <= kSsMaxRows; }

size_t sample_sort_workspace_bytes(int64_t rows, int segments) {
    if (rows <= 0 || segments <= 0) return 0;
    return ss_layout(rows, segments).total;
}

int sample_sort_launch(float* keys_device, int64_t rows, int segments, void* ws, cudaStream_t stream, const DiffArgs& da,
                       const uint32_t** flag_device) {
    const SsPlan p = ss_plan(rows);
    const SsLayout l = ss_layout(rows, segments);
    char* w = static_cast<char*>(ws);
    uint32_t* ctrl = reinterpret_cast<uint32_t*>(w + l.ctrl);
    uint32_t* keys = reinterpret_cast<uint32_t*>(keys_device);
    BGAN_CUDA_OK(cudaMemsetAsync(ctrl, 0, 256, stream));
    if (const char* f = getenv("BGAN_SORT_FORCE_FALLBACK")) {  // tests only: pretend a bucket overflowed
        if (atoi(f) && !p.direct) BGAN_CUDA_OK(cudaMemsetAsync(ctrl, 1, 1, stream));
    }

    thread_local int attr_dev = -1;
    int dev = 0;
    BGAN_CUDA_OK(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        BGAN_CUDA_OK(cudaFuncSetAttribute(ss_splitters_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSsMaxBuckets * kSsOversample * 4));
        BGAN_CUDA_OK(cudaFuncSetAttribute(ss_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kScatterSmem)));
        BGAN_CUDA_OK(cudaFuncSetAttribute(ss_local_sort_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(LocalSmem))));
        BGAN_CUDA_OK(cudaFuncSetAttribute(ss_local_sort_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(LocalSmem))));
        attr_dev = dev;
    }

    LocalArgs a;
    a.rows = rows;
    a.segments = segments;
    a.ctrl = ctrl;
    a.da = da;
    a.dst = keys;
    if (p.direct) {
        a.src = keys;
        a.n_ids = 0;
        a.buckets = 0;
        a.starts = nullptr;
        a.bounds = nullptr;
        *flag_device = nullptr;
    } else {
        uint32_t* scratch = reinterpret_cast<uint32_t*>(w + l.scratch);
        uint16_t* ids = reinterpret_cast<uint16_t*>(w + l.ids);
        uint16_t* counts = reinterpret_cast<uint16_t*>(w + l.counts);
        uint32_t* bases = reinterpret_cast<uint32_t*>(w + l.bases);
        uint32_t* starts = reinterpret_cast<uint32_t*>(w + l.starts);
        uint32_t* bounds = reinterpret_cast<uint32_t*>(w + l.bounds);
        ss_splitters_kernel<<<segments, kSplThreads, p.samples * sizeof(uint32_t), stream>>>(keys, rows, p.buckets, bounds);
        const dim3 grid(p.blocks, segments);
        ss_classify_kernel<<<grid, kSsThreads, 3 * p.buckets * sizeof(uint32_t), stream>>>(keys, rows, p.buckets, p.blocks, bounds, ids, counts);
        ss_scan_kernel<<<segments, kSplThreads, 0, stream>>>(counts, rows, p.ids, p.blocks, starts, bases, ctrl);
        ss_scatter_kernel<<<grid, kSsThreads, kScatterSmem, stream>>>(keys, ids, rows, p.ids, p.blocks, counts, bases, scratch, ctrl);
        a.src = scratch;
        a.n_ids = p.ids;
        a.buckets = p.buckets;
        a.starts = starts;
        a.bounds = bounds;
        *flag_device = ctrl;
    }
    const int batches = p.direct ? segments : segments * (p.ids / std::min(kLsBatchIds, p.ids));
    const int grid_ls = std::max(1, std::min(batches, 2 * device_sm_count()));
    if (da.ref != nullptr)
        ss_local_sort_kernel<true><<<grid_ls, kLsThreads, sizeof(LocalSmem), stream>>>(a);
    else
        ss_local_sort_kernel<false><<<grid_ls, kLsThreads, sizeof(LocalSmem), stream>>>(a);
    BGAN_LAUNCH_OK();
    return 0;
}

}  // namespace bgan
