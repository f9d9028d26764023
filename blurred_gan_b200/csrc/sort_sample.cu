// Segmented sort of fp32 keys by value buckets in shared memory (sm_100a) -- the np.sort(proj, axis=0) of the reference's
// sliced_wasserstein.py:47-48 for segments of up to 2^21 keys, and its fusion with the mean |A - B| of :49-50.
//
// The LSD radix sort of sort.cu moves every key through HBM four times and ranks it four times with eight ballots (it is
// instruction-bound: profiles/r02_sort.md).  Here a key crosses HBM twice and is ranked with shared-memory atomics; the
// order comes from two levels of linear interpolation on the VALUE (the keys of a sliced Wasserstein distance are
// projections of normalised descriptors: a smooth, light-tailed density):
//
//   model      one CTA per segment: 16 B stratified samples -> [lo, hi]; histogram of the samples over C = 16 B equal cells of
//              [lo, hi]; prefix sums -> lut[cell] = bucket (a bucket = a run of cells holding ~16 samples, i.e. ~rows / B =
//              ~2048 keys).  Monotone by construction: cell is a non-decreasing function of the key, lut of the cell.
//   classify   bucket id = lut[clamp((x - lo) * scale)] -> ids (uint16); per block of 16384 keys the bucket counts
//              (shared-memory atomics) -> counts[seg][block][id]
//   scan       per segment: bucket starts, and the start of every (block, id) run -> bases[seg][block][id]; raises a flag when
//              a bucket holds more keys than a CTA can (outliers that stretch [lo, hi], one value that occurs thousands of
//              times): nothing is written then and the caller falls back to the radix sort
//   scatter    per block: keys regrouped by id in shared memory (slot = local start + shared-memory atomic), written as
//              runs to scratch.  Not stable -- the last step sorts the buckets anyway.
//   local sort persistent CTAs take runs of adjacent buckets (<= 8192 keys): sub-bucket = linear map of the value between the
//              run's bounds (0.5 - 1 key per sub-bucket), count with atomics, scan, place, then -- in slot order, so that neighbouring
//              lanes read neighbouring words -- rank every key against its sub-bucket (four unconditional probes) and write it to its
//              final position (or, fused: meet the sorted A there and accumulate |a - b|).  A run whose sub-buckets are badly
//              loaded (clustered values, non-finite bounds) is sorted by a bitonic network in shared memory instead: slower,
//              never wrong.
//
// One or a few values repeated in several per cent of the keys (saturated image patches give identical descriptors) would overflow
// their bucket however the cells are drawn: the model looks at every cell that holds >= 24 samples, and if most of them are ONE
// value, that value gets an "equality bucket" of its own (never sorted, any size) between the keys below and above it.
//
// HBM traffic per key: 4 (classify) + 2 + 2 (ids) + 4 + 4 (scatter) + 4 + 4 (local sort) = 24 bytes + ~2 of tables, against 48.
// Segments of <= 8192 keys skip everything but the last step.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "sort.cuh"

namespace bgan {
namespace {

constexpr int kSsOversample = 16;      // samples per bucket
constexpr int kSsOversampleLog2 = 4;
constexpr int kSsCellsPerBucket = 16;
constexpr int kSsMeanBucket = 2048;
constexpr int kSsMaxBuckets = 1024;
constexpr int kSsMaxCells = kSsMaxBuckets * kSsCellsPerBucket;
constexpr int kSsMaxHeavy = 16;        // values per segment that get a bucket of their own ("equality buckets")
constexpr int kSsHeavySamples = 24;    // a cell with this many of the samples (1.5 buckets' worth) is examined for one
constexpr int kSsMaxIds = kSsMaxBuckets + 2 * kSsMaxHeavy;  // bucket ids: a heavy value splits its bucket in three
constexpr int kSsCap = 8192;           // keys a local-sort CTA holds
constexpr int kSsBlock = 16384;        // keys per (classify, scatter) block: one row of counts / bases
constexpr int kSsClsBlocks = 4;        // blocks per classify CTA (one load of the lut)
constexpr int kSsThreads = 256;        // classify
constexpr int kScThreads = 256;        // scatter (512: measured slower, 365 against 340 us)
constexpr int64_t kSsMaxRows = int64_t(1) << 21;
constexpr int kModelThreads = 1024;    // model / scan kernels
constexpr int kLsThreads = 512;        // local sort
constexpr int kLsWarps = kLsThreads / 32;
constexpr int kLsItems = kSsCap / kLsThreads;  // 16 keys per thread
constexpr int kLsMaxSub = 8192;         // sub-buckets of a run: the power of two >= its keys (0.5 .. 1 key per sub-bucket)
constexpr int kLsBatchIds = 16;        // buckets per ticket

struct SsPlan {
    int buckets, ids, cells, blocks, samples;
    bool direct;
};

SsPlan ss_plan(int64_t rows) {
    SsPlan p;
    p.direct = rows <= kSsCap;
    const int64_t want = ceil_div64(rows, kSsMeanBucket);
    int b = 2;
    while (b < want && b < kSsMaxBuckets) b <<= 1;
    p.buckets = b;
    p.ids = b + 2 * kSsMaxHeavy;
    p.cells = b * kSsCellsPerBucket;
    p.blocks = static_cast<int>(ceil_div64(rows, kSsBlock));
    p.samples = b * kSsOversample;
    return p;
}

size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

struct SsLayout {
    size_t scratch, ids, counts, bases, starts, lut, idinfo, heavy, segpar, ctrl, total;
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
    l.scratch = l.ids = l.counts = l.bases = l.starts = l.lut = l.idinfo = l.heavy = l.segpar = 0;
    if (!p.direct) {
        l.scratch = take(n * sizeof(uint32_t));
        l.ids = take(n * sizeof(uint16_t));
        l.counts = take(tab * sizeof(uint16_t));
        l.bases = take(tab * sizeof(uint32_t));
        l.starts = take(static_cast<size_t>(segments) * (p.ids + 1) * sizeof(uint32_t));
        l.lut = take(static_cast<size_t>(segments) * p.cells * sizeof(uint16_t));
        l.idinfo = take(static_cast<size_t>(segments) * p.ids * sizeof(uint2));
        l.heavy = take(static_cast<size_t>(segments) * kSsMaxHeavy * sizeof(uint2));
        l.segpar = take(static_cast<size_t>(segments) * sizeof(float2));
    }
    l.total = at;
    return l;
}

// cell (or sub-bucket) of a value: non-decreasing in the key's ordered-uint image whatever lo and scale are -- subtraction,
// product, truncation and clamp all are monotone; NaNs go to the end their sign puts them at in the ordered-uint order
__device__ __forceinline__ int linear_cell(uint32_t float_bits, float lo, float scale, int cells) {
    const float x = __uint_as_float(float_bits);
    int c = min(cells - 1, max(0, __float2int_rz((x - lo) * scale)));
    if (x != x) c = (float_bits >> 31) ? 0 : cells - 1;
    return c;
}
// the same for keys known not to be NaN
__device__ __forceinline__ int linear_cell_finite(uint32_t float_bits, float lo, float scale, int cells) {
    return min(cells - 1, max(0, __float2int_rz((__uint_as_float(float_bits) - lo) * scale)));
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

// ---- model ------------------------------------------------------------------------------------------------------------------
// lut[cell]: the cell's bucket id, or 0x8000 | k for the cell of heavy value k.  A heavy value -- a cell that holds >= 24 samples, all
// of them the same number: a value repeated in several per cent of the keys (saturated patches give identical descriptors) -- splits
// its bucket in three: ids id0 (keys of the cell below it, together with the cells before), id0 + 1 (the value itself: an
// "equality bucket" that needs no sorting and may have any size) and id0 + 2 (above it, together with the cells behind).
// idinfo[id] = (first cell | equality flag << 31, last cell + 1); heavy[k] = (the value's bits, id0).
__global__ void __launch_bounds__(kModelThreads) ss_model_kernel(const uint32_t* __restrict__ keys, int64_t rows, int buckets,
                                                                 uint16_t* __restrict__ lut, uint2* __restrict__ idinfo,
                                                                 uint2* __restrict__ heavy, float2* __restrict__ segpar) {
    extern __shared__ uint32_t s_hist[];  // [cells]: sample histogram (| tag << 24), then the cell's first id (| special << 31)
    __shared__ float s_lo[kModelThreads / 32], s_hi[kModelThreads / 32];
    __shared__ uint32_t s_warp[kModelThreads / 32], s_warp2[kModelThreads / 32];
    __shared__ uint32_t s_nheavy, s_hcell[kSsMaxHeavy], s_hfirst[kSsMaxHeavy], s_hval[kSsMaxHeavy], s_heq[kSsMaxHeavy], s_hid0[kSsMaxHeavy];
    __shared__ uint32_t s_cell_lo[kSsMaxIds], s_cell_hi[kSsMaxIds];
    const int seg = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = buckets * kSsOversample, cells = buckets * kSsCellsPerBucket, ids = buckets + 2 * kSsMaxHeavy;
    const uint32_t* src = keys + static_cast<size_t>(seg) * rows;
    constexpr int kPerThread = kSsMaxBuckets * kSsOversample / kModelThreads;  // samples, and cells, per thread (at most)
    static_assert(kSsMaxCells / kModelThreads == kPerThread, "as many cells as samples");
    uint32_t x[kPerThread];
    float lo = INFINITY, hi = -INFINITY;
#pragma unroll
    for (int e = 0; e < kPerThread; ++e) {
        const int i = e * kModelThreads + tid;
        x[e] = 0;
        if (i < n) {
            // one sample from each of n equal strata, at a hashed position inside the stratum
            const int64_t a = static_cast<int64_t>(i) * rows / n, b = static_cast<int64_t>(i + 1) * rows / n;
            uint32_t h = static_cast<uint32_t>(i) * 0x9E3779B1u + static_cast<uint32_t>(seg) * 0x85EBCA77u;
            h ^= h >> 15;
            h *= 0x2C1B3C6Du;
            h ^= h >> 12;
            const uint32_t width = static_cast<uint32_t>(b - a);
            x[e] = __ldg(src + a + (width > 1 ? h % width : 0u));
            lo = fminf(lo, __uint_as_float(x[e]));  // (fminf / fmaxf skip NaNs)
            hi = fmaxf(hi, __uint_as_float(x[e]));
        }
    }
    for (int c = tid; c < cells; c += kModelThreads) s_hist[c] = 0;
    for (int i = tid; i < ids; i += kModelThreads) s_cell_lo[i] = s_cell_hi[i] = static_cast<uint32_t>(cells);
    if (tid == 0) s_nheavy = 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) s_lo[warp] = lo, s_hi[warp] = hi;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < kModelThreads / 32; ++w) lo = fminf(lo, s_lo[w]), hi = fmaxf(hi, s_hi[w]);
    // (hi == lo, or an infinite bound: scale 0 or NaN -> every key in one cell; a heavy value if it is one number, else the
    // scan kernel raises the flag)
    const float scale = hi > lo ? static_cast<float>(cells) / (hi - lo) : 0.0f;
    int cell_of[kPerThread];
#pragma unroll
    for (int e = 0; e < kPerThread; ++e) {
        cell_of[e] = 0;
        if (e * kModelThreads + tid < n) {
            cell_of[e] = linear_cell(x[e], lo, scale, cells);
            atomicAdd(&s_hist[cell_of[e]], 1u);
        }
    }
    __syncthreads();
    // ---- heavy cells: tag them; candidate value = the cell's sample with the lowest index; count the samples equal to it ----
    for (int c = tid; c < cells; c += kModelThreads) {
        if (s_hist[c] >= static_cast<uint32_t>(kSsHeavySamples)) {
            const uint32_t k = atomicAdd(&s_nheavy, 1u);
            if (k < static_cast<uint32_t>(kSsMaxHeavy)) {
                s_hcell[k] = c;
                s_hfirst[k] = 0xffffffffu;
                s_heq[k] = 0u;
                s_hist[c] |= (k + 1) << 24;
            }
        }
    }
    __syncthreads();
    const bool any_heavy = s_nheavy != 0;  // (uniform; smooth data: no cell is heavy and the three passes below are skipped)
    if (any_heavy) {
        uint32_t tag_of[kPerThread];
#pragma unroll
        for (int e = 0; e < kPerThread; ++e) {
            const uint32_t i = static_cast<uint32_t>(e * kModelThreads + tid);
            tag_of[e] = i < static_cast<uint32_t>(n) ? s_hist[cell_of[e]] >> 24 : 0u;
            if (tag_of[e]) atomicMin(&s_hfirst[tag_of[e] - 1], i);
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < kPerThread; ++e)
            if (tag_of[e] && s_hfirst[tag_of[e] - 1] == static_cast<uint32_t>(e * kModelThreads + tid)) s_hval[tag_of[e] - 1] = x[e];
        __syncthreads();
#pragma unroll
        for (int e = 0; e < kPerThread; ++e)
            if (tag_of[e] && __uint_as_float(x[e]) == __uint_as_float(s_hval[tag_of[e] - 1])) atomicAdd(&s_heq[tag_of[e] - 1], 1u);
        __syncthreads();
    }
    if (tid < kSsMaxHeavy) {
        const bool tagged = static_cast<uint32_t>(tid) < min(s_nheavy, static_cast<uint32_t>(kSsMaxHeavy));
        const bool single = tagged && s_heq[tid] >= static_cast<uint32_t>(kSsHeavySamples);  // (a NaN candidate equals nothing)
        if (tagged && !single) s_hist[s_hcell[tid]] &= 0xffffffu;  // no dominant value: an ordinary cell (it may still overflow a bucket)
        if (!single) s_hcell[tid] = 0xffffffffu;
        s_hid0[tid] = 0xffffffffu;
    }
    __syncthreads();
    // ---- exclusive prefixes over the cells (thread t owns `per` consecutive cells): samples outside heavy values -> bucket,
    //      heavy values before the cell -> two more ids each ----
    const int per = cells >= kModelThreads ? cells / kModelThreads : 1;  // 1 .. 16
    uint32_t raw[kPerThread];
    uint32_t sum = 0, sum2 = 0;
#pragma unroll
    for (int e = 0; e < kPerThread; ++e) {
        const int cell = tid * per + e;
        raw[e] = (e < per && cell < cells) ? s_hist[cell] : 0u;
        if (raw[e] >> 24)
            sum2 += 2;  // (its few other samples are left out of the bucket prefix)
        else
            sum += raw[e];
    }
    uint32_t incl = sum, incl2 = sum2;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o), t2 = __shfl_up_sync(0xffffffffu, incl2, o);
        if (lane >= o) incl += t, incl2 += t2;
    }
    if (lane == 31) s_warp[warp] = incl, s_warp2[warp] = incl2;
    __syncthreads();
    uint32_t run = incl - sum, run2 = incl2 - sum2;
    for (int w = 0; w < warp; ++w) run += s_warp[w], run2 += s_warp2[w];
    uint16_t* lut_out = lut + static_cast<size_t>(seg) * cells;
#pragma unroll
    for (int e = 0; e < kPerThread; ++e) {
        const int cell = tid * per + e;
        if (e < per && cell < cells) {
            const uint32_t bucket = min(static_cast<uint32_t>(buckets - 1), run >> kSsOversampleLog2) + run2;
            const uint32_t tag = raw[e] >> 24;
            if (tag) {
                s_hid0[tag - 1] = bucket;
                s_hist[cell] = bucket | 0x80000000u;
                lut_out[cell] = static_cast<uint16_t>(0x8000u | (tag - 1));
                run2 += 2;
            } else {
                s_hist[cell] = bucket;
                lut_out[cell] = static_cast<uint16_t>(bucket);
                run += raw[e];
            }
        }
    }
    __syncthreads();
    // ---- the cells of every id: an id first appears in exactly one cell and last appears in exactly one cell ----
    for (int cell = tid; cell < cells; cell += kModelThreads) {
        const uint32_t v = s_hist[cell];
        const int first = static_cast<int>(v & 0x7fffffffu), last = first + ((v >> 31) ? 2 : 0);
        int prev_last = -1, next_first = 0x7fffffff;
        if (cell > 0) {
            const uint32_t pv = s_hist[cell - 1];
            prev_last = static_cast<int>(pv & 0x7fffffffu) + ((pv >> 31) ? 2 : 0);
        }
        if (cell + 1 < cells) next_first = static_cast<int>(s_hist[cell + 1] & 0x7fffffffu);
        for (int id = max(first, prev_last + 1); id <= last; ++id) s_cell_lo[id] = static_cast<uint32_t>(cell);
        for (int id = first; id <= min(last, next_first - 1); ++id) s_cell_hi[id] = static_cast<uint32_t>(cell + 1);
    }
    __syncthreads();
    uint2* info = idinfo + static_cast<size_t>(seg) * ids;
    for (int id = tid; id < ids; id += kModelThreads) {
        uint32_t eq = 0;
        if (any_heavy) {
#pragma unroll
            for (int k = 0; k < kSsMaxHeavy; ++k) eq |= (s_hid0[k] != 0xffffffffu && s_hid0[k] + 1 == static_cast<uint32_t>(id)) ? 1u : 0u;
        }
        info[id] = make_uint2(s_cell_lo[id] | (eq << 31), s_cell_hi[id]);
    }
    if (tid < kSsMaxHeavy)
        heavy[static_cast<size_t>(seg) * kSsMaxHeavy + tid] =
            make_uint2(s_hcell[tid] != 0xffffffffu ? s_hval[tid] : 0u, s_hid0[tid]);
    if (tid == 0) segpar[seg] = make_float2(lo, scale);
}

// ---- classify ----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSsThreads) ss_classify_kernel(const uint32_t* __restrict__ keys, int64_t rows, int buckets, int blocks,
                                                                 const uint16_t* __restrict__ lut, const uint2* __restrict__ heavy,
                                                                 const float2* __restrict__ segpar, uint16_t* __restrict__ ids,
                                                                 uint16_t* __restrict__ counts) {
    extern __shared__ __align__(16) unsigned char s_cls[];  // lut [cells] uint16 | hist [n_ids] uint32 | heavy values [kSsMaxHeavy] uint2
    const int cells = buckets * kSsCellsPerBucket, n_ids = buckets + 2 * kSsMaxHeavy;
    uint16_t* s_lut = reinterpret_cast<uint16_t*>(s_cls);
    uint32_t* hist = reinterpret_cast<uint32_t*>(s_cls + cells * sizeof(uint16_t));
    uint2* s_hv = reinterpret_cast<uint2*>(hist + n_ids);
    const int seg = blockIdx.y, tid = threadIdx.x;
    bool mine_heavy = false;
    if (tid < kSsMaxHeavy) {
        s_hv[tid] = heavy[static_cast<size_t>(seg) * kSsMaxHeavy + tid];
        mine_heavy = s_hv[tid].y != 0xffffffffu;
    }
    const bool seg_heavy = __syncthreads_or(mine_heavy) != 0;
    {
        const uint4* g = reinterpret_cast<const uint4*>(lut + static_cast<size_t>(seg) * cells);  // (cells is a multiple of 32)
        for (int i = tid; i < cells / 8; i += kSsThreads) reinterpret_cast<uint4*>(s_lut)[i] = __ldg(g + i);
    }
    const float2 par = segpar[seg];
    const float lo = par.x, scale = par.y;
    for (int bb = 0; bb < kSsClsBlocks; ++bb) {
        const int blk = blockIdx.x * kSsClsBlocks + bb;
        if (blk >= blocks) break;
        for (int j = tid; j < n_ids; j += kSsThreads) hist[j] = 0;
        __syncthreads();
        const int64_t base = static_cast<int64_t>(blk) * kSsBlock;
        const int valid = rows - base < kSsBlock ? static_cast<int>(rows - base) : kSsBlock;
        const size_t at = static_cast<size_t>(seg) * rows + base;
        const uint32_t* src = keys + at;
        uint16_t* dst = ids + at;
        auto block = [&](auto heavy_tag) {  // (two bodies: segments without heavy values skip the test per key)
            constexpr bool kHeavy = decltype(heavy_tag)::value;
            auto one = [&](uint32_t bits) {
                uint32_t id = s_lut[linear_cell(bits, lo, scale, cells)];
                if constexpr (kHeavy) {
                    if (id & 0x8000u) {  // the cell of a heavy value: below it / the value / above it
                        const uint2 h = s_hv[id & 0x7fffu];
                        const float x = __uint_as_float(bits), v = __uint_as_float(h.x);
                        id = h.y + (x == v ? 1u : (x > v ? 2u : 0u));
                        if (x != x) id = (bits >> 31) ? h.y : h.y + 2u;
                    }
                }
                atomicAdd(&hist[id], 1u);
                return id;
            };
            int done = 0;
            if ((at & 3) == 0) {
                const int vec = valid >> 2;
#pragma unroll 4
                for (int v = tid; v < vec; v += kSsThreads) {
                    const uint4 k = __ldg(reinterpret_cast<const uint4*>(src) + v);
                    const uint32_t i0 = one(k.x), i1 = one(k.y), i2 = one(k.z), i3 = one(k.w);
                    reinterpret_cast<uint2*>(dst)[v] = make_uint2(i0 | (i1 << 16), i2 | (i3 << 16));
                }
                done = vec << 2;
            }
            for (int i = done + tid; i < valid; i += kSsThreads) dst[i] = static_cast<uint16_t>(one(__ldg(src + i)));
        };
        if (seg_heavy)
            block(std::true_type{});
        else
            block(std::false_type{});
        __syncthreads();
        uint16_t* c = counts + (static_cast<size_t>(seg) * blocks + blk) * n_ids;
        for (int j = tid; j < n_ids; j += kSsThreads) c[j] = static_cast<uint16_t>(hist[j]);
        __syncthreads();
    }
}

// ---- scan ----------------------------------------------------------------------------------------------------------------------
// starts[seg][id] (n_ids + 1 entries) and bases[seg][block][id] = starts[id] + keys of that id in earlier blocks
__global__ void __launch_bounds__(kModelThreads) ss_scan_kernel(const uint16_t* __restrict__ counts, int64_t rows, int n_ids, int blocks,
                                                                const uint2* __restrict__ idinfo, uint32_t* __restrict__ starts,
                                                                uint32_t* __restrict__ bases, uint32_t* __restrict__ ctrl) {
    __shared__ uint32_t s_warp[kModelThreads / 32];
    __shared__ uint32_t s_base;
    const int seg = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint16_t* c = counts + static_cast<size_t>(seg) * blocks * n_ids;
    const uint2* info = idinfo + static_cast<size_t>(seg) * n_ids;
    uint32_t* st = starts + static_cast<size_t>(seg) * (n_ids + 1);
    uint32_t* bs = bases + static_cast<size_t>(seg) * blocks * n_ids;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int id0 = 0; id0 < n_ids; id0 += kModelThreads) {  // (one or two rounds)
        const int id = id0 + tid;
        uint32_t total = 0;
        if (id < n_ids) {
#pragma unroll 8
            for (int b = 0; b < blocks; ++b) total += c[static_cast<size_t>(b) * n_ids + id];
            // a bucket that no CTA can hold (the bucket of a heavy value is never sorted and may have any size)
            if (total > static_cast<uint32_t>(kSsCap) && (info[id].x >> 31) == 0) atomicOr(ctrl, 1u);
        }
        uint32_t incl = total;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t run = s_base + incl - total;
        for (int w = 0; w < warp; ++w) run += s_warp[w];
        if (id < n_ids) {
            st[id] = run;
            uint32_t r = run;
#pragma unroll 8
            for (int b = 0; b < blocks; ++b) {
                bs[static_cast<size_t>(b) * n_ids + id] = r;
                r += c[static_cast<size_t>(b) * n_ids + id];
            }
        }
        __syncthreads();
        if (tid == kModelThreads - 1) s_base = run + total;
        __syncthreads();
    }
    if (tid == 0) st[n_ids] = static_cast<uint32_t>(rows);
}

// ---- scatter -------------------------------------------------------------------------------------------------------------------
// Two ways out of shared memory.  RUNS: a group of lanes copies every bucket's run -- no id per slot, 72 KB, three CTAs per SM;
// right when a block holds long runs (<= 512 buckets: >= 32 keys per run).  Otherwise every slot carries its bucket id and the
// threads walk the slots (104 KB, two CTAs per SM): measured at 1024 buckets, 728 us against 826 us per 268 M keys.
template <bool RUNS>
constexpr size_t scatter_smem() {
    return kSsBlock * sizeof(uint32_t) + (RUNS ? 0 : kSsBlock * sizeof(uint16_t)) + 2 * kSsMaxIds * sizeof(uint32_t) + 64;
}

template <bool RUNS>
__global__ void __launch_bounds__(kScThreads, RUNS ? 3 : 2) ss_scatter_kernel(const uint32_t* __restrict__ keys, const uint16_t* __restrict__ ids,
                                                                             int64_t rows, int n_ids, int blocks,
                                                                             const uint16_t* __restrict__ counts, const uint32_t* __restrict__ bases,
                                                                             uint32_t* __restrict__ out, const uint32_t* __restrict__ ctrl) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    uint32_t* s_key = reinterpret_cast<uint32_t*>(s_raw);
    uint32_t* s_cur = s_key + kSsBlock;  // first free local slot of a bucket; after the regrouping: the end of its run
    uint32_t* s_delta = s_cur + kSsMaxIds;
    uint32_t* s_warp = s_delta + kSsMaxIds;
    uint16_t* s_id = reinterpret_cast<uint16_t*>(s_warp + 16);  // (!RUNS only)
    if (*ctrl != 0) return;
    const int blk = blockIdx.x, seg = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t row = (static_cast<size_t>(seg) * blocks + blk) * n_ids;
    // local starts: exclusive scan of this block's counts; thread t owns `per` consecutive ids
    constexpr int kMaxPer = (kSsMaxIds + kScThreads - 1) / kScThreads;
    const int per = (n_ids + kScThreads - 1) / kScThreads;  // 1 .. 5
    uint32_t c[kMaxPer];
    uint32_t sum = 0;
#pragma unroll
    for (int e = 0; e < kMaxPer; ++e) {
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
    for (int e = 0; e < kMaxPer; ++e) {
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
    auto put = [&](uint32_t id, uint32_t key) {
        const uint32_t p = atomicAdd(&s_cur[id], 1u);
        s_key[p] = key;
        if constexpr (!RUNS) s_id[p] = static_cast<uint16_t>(id);
    };
    int done = 0;
    if ((at & 3) == 0) {
        const int vec = valid >> 2;
#pragma unroll 4
        for (int v = tid; v < vec; v += kScThreads) {
            const uint4 k = __ldg(reinterpret_cast<const uint4*>(src) + v);
            const uint2 i2 = __ldg(reinterpret_cast<const uint2*>(sid) + v);
            put(i2.x & 0xffff, k.x);
            put(i2.x >> 16, k.y);
            put(i2.y & 0xffff, k.z);
            put(i2.y >> 16, k.w);
        }
        done = vec << 2;
    }
    for (int i = done + tid; i < valid; i += kScThreads) put(sid[i], __ldg(src + i));
    __syncthreads();
    uint32_t* dst = out + static_cast<size_t>(seg) * rows;
    if constexpr (RUNS) {
        for (int id = warp; id < n_ids; id += kScThreads / 32) {
            const uint32_t ls = id ? s_cur[id - 1] : 0u, le = s_cur[id], d = s_delta[id];
            for (uint32_t slot = ls + lane; slot < le; slot += 32) dst[d + slot] = s_key[slot];
        }
    } else {
#pragma unroll 4
        for (int slot = tid; slot < valid; slot += kScThreads) dst[s_delta[s_id[slot]] + slot] = s_key[slot];
    }
}

// ---- local sort ------------------------------------------------------------------------------------------------------------------
struct LocalSmem {
    uint32_t buf[kSsCap + 4];  // (+4: the ranking step probes up to four slots past a run's last key)
    uint32_t cnt[kLsMaxSub + 32];
    uint32_t starts[kLsBatchIds + 1];
    uint2 info[kLsBatchIds];   // (first cell | equality flag << 31, last cell + 1) of the ticket's buckets
    uint32_t warp[kLsWarps];
    uint32_t red[2][kLsWarps];
    double dsum[kLsWarps];
    uint32_t ticket;
    uint32_t cost;     // sum of squares of the sub-bucket counts of the current run
    uint32_t has_nan;  // the current run holds a NaN
};

struct LocalArgs {
    const uint32_t* src;         // bucketed keys (scratch), or the keys themselves when a segment is one run (direct)
    uint32_t* dst;               // sorted floats (unused with da.ref)
    int64_t rows;
    int segments;
    int n_ids;                   // bucket ids per segment; 0: direct
    const uint32_t* starts;      // [seg][n_ids + 1]
    const uint2* idinfo;         // [seg][n_ids]
    const float2* segpar;        // [seg] (lo, cells / (hi - lo))
    uint32_t* ctrl;              // [0] overflow flag, [1] ticket counter
    DiffArgs da;
};

// exclusive scan of cnt[0, n_sub) in place, thread t owning PER consecutive counters (vector accesses: a stride of PER words
// between lanes would be a PER-way bank conflict); returns this thread's sum of squares
template <int PER>
__device__ __forceinline__ void scan_counts(LocalSmem& sm, int n_sub, int tid, int lane, int warp) {
    uint32_t c[PER];
    const int base = tid * PER;
    const bool on = base < n_sub;
    if constexpr (PER >= 4) {
#pragma unroll
        for (int v = 0; v < PER / 4; ++v) {
            const uint4 a = on ? *reinterpret_cast<const uint4*>(&sm.cnt[base + 4 * v]) : make_uint4(0, 0, 0, 0);
            c[4 * v] = a.x, c[4 * v + 1] = a.y, c[4 * v + 2] = a.z, c[4 * v + 3] = a.w;
        }
    } else if constexpr (PER == 2) {
        const uint2 a = on ? *reinterpret_cast<const uint2*>(&sm.cnt[base]) : make_uint2(0, 0);
        c[0] = a.x, c[1] = a.y;
    } else {
        c[0] = on ? sm.cnt[base] : 0u;
    }
    uint32_t sum = 0, sq = 0;
#pragma unroll
    for (int e = 0; e < PER; ++e) sum += c[e], sq += c[e] * c[e];
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
    uint32_t e0[PER];
#pragma unroll
    for (int e = 0; e < PER; ++e) e0[e] = run, run += c[e];
    if (on) {
        if constexpr (PER >= 4) {
#pragma unroll
            for (int v = 0; v < PER / 4; ++v)
                *reinterpret_cast<uint4*>(&sm.cnt[base + 4 * v]) = make_uint4(e0[4 * v], e0[4 * v + 1], e0[4 * v + 2], e0[4 * v + 3]);
        } else if constexpr (PER == 2) {
            *reinterpret_cast<uint2*>(&sm.cnt[base]) = make_uint2(e0[0], e0[1]);
        } else {
            sm.cnt[base] = e0[0];
        }
    }
}

// a run that is already in order (the bucket of a heavy value: all keys equal): copied, or met with the sorted A, as it is.
// (Not inlined: its unrolled loads would cost the sorting path registers.)
template <bool DIFF>
__device__ __noinline__ double pass_through(const uint32_t* __restrict__ src, const float* __restrict__ ref_run, uint32_t* __restrict__ dst_run,
                                            int n, int tid) {
    double acc = 0.0;
    int i = tid;
    for (; i + 3 * kLsThreads < n; i += 4 * kLsThreads) {  // four loads in flight per thread
        uint32_t v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldg(src + i + u * kLsThreads);
        if constexpr (DIFF) {
            float r[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) r[u] = __ldg(ref_run + i + u * kLsThreads);
#pragma unroll
            for (int u = 0; u < 4; ++u) acc += static_cast<double>(fabsf(r[u] - __uint_as_float(v[u])));
        } else {
#pragma unroll
            for (int u = 0; u < 4; ++u) dst_run[i + u * kLsThreads] = v[u];
        }
    }
    for (; i < n; i += kLsThreads) {
        const uint32_t v = __ldg(src + i);
        if constexpr (DIFF)
            acc += static_cast<double>(fabsf(__ldg(ref_run + i) - __uint_as_float(v)));
        else
            dst_run[i] = v;
    }
    return acc;
}

template <bool DIFF>
__global__ void __launch_bounds__(kLsThreads, 2) ss_local_sort_kernel(const LocalArgs a) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    LocalSmem& sm = *reinterpret_cast<LocalSmem*>(s_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool direct = a.n_ids == 0;
    if (!direct && a.ctrl[0] != 0) return;
    const int batch_ids = direct ? 1 : kLsBatchIds;
    const int batches_per_seg = direct ? 1 : (a.n_ids + kLsBatchIds - 1) / kLsBatchIds;
    const uint32_t total = static_cast<uint32_t>(a.segments) * batches_per_seg;
    for (int i = tid; i < kLsMaxSub + 32; i += kLsThreads) sm.cnt[i] = 0;
    if (tid == 0) sm.cost = 0, sm.has_nan = 0;

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
        if (tid <= batch_ids)  // (ids past the last one: empty)
            sm.starts[tid] = direct ? (tid == 0 ? 0u : static_cast<uint32_t>(a.rows))
                                    : a.starts[static_cast<size_t>(seg) * (a.n_ids + 1) + min(id0 + tid, a.n_ids)];
        if (!direct && tid < batch_ids)
            sm.info[tid] = id0 + tid < a.n_ids ? a.idinfo[static_cast<size_t>(seg) * a.n_ids + id0 + tid] : make_uint2(0u, 0u);
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
            // a run of adjacent buckets that fits the CTA; the bucket of a heavy value (all keys equal) goes through unsorted when it
            // is large, and inside a run when it is not
            auto big_eq = [&](int g) { return !direct && (sm.info[g].x >> 31) && sm.starts[g + 1] - sm.starts[g] > 64u; };
            const bool pass = big_eq(g0);
            int g1 = g0 + 1;
            if (!pass)
                while (g1 < batch_ids && !big_eq(g1) && static_cast<int>(sm.starts[g1 + 1] - sm.starts[g0]) <= kSsCap) ++g1;
            const uint32_t start = sm.starts[g0];
            const int n = static_cast<int>(sm.starts[g1] - start);
            const int first_g = g0, last_g = g1 - 1;
            g0 = g1;
            if (n == 0) continue;
            const uint32_t* src = a.src + seg_at + start;
            const float* const ref_run = DIFF ? a.da.ref + seg_at + start : nullptr;
            uint32_t* const dst_run = DIFF ? nullptr : a.dst + seg_at + start;
            auto emit = [&](int pos, uint32_t float_bits) {  // the key whose final position in the run is `pos`
                if constexpr (DIFF)
                    acc += static_cast<double>(fabsf(__ldg(ref_run + pos) - __uint_as_float(float_bits)));
                else
                    dst_run[pos] = float_bits;
            };
            if (pass || n > kSsCap) {  // (n > kSsCap without `pass` cannot happen: the scan kernel raised the flag and nothing runs)
                acc += pass_through<DIFF>(src, ref_run, dst_run, n, tid);
                continue;
            }

            // ---- keys -> registers; bounds of the run ----
            uint32_t k[kLsItems];  // raw float bits
            {
                const uint32_t* mine = src + tid;
                const int left = n - tid;
#pragma unroll
                for (int q = 0; q < kLsItems; ++q) k[q] = q * kLsThreads < left ? __ldg(mine + q * kLsThreads) : 0u;
            }
            float flo, fhi;
            if (direct) {
                uint32_t mn = 0xffffffffu, mx = 0u;
#pragma unroll
                for (int q = 0; q < kLsItems; ++q) {
                    if (q * kLsThreads + tid < n) {
                        const uint32_t u = float_to_ordered(k[q]);
                        mn = min(mn, u), mx = max(mx, u);
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                }
                if (lane == 0) sm.red[0][warp] = mn, sm.red[1][warp] = mx;
                __syncthreads();
                mn = sm.red[0][0], mx = sm.red[1][0];
#pragma unroll
                for (int w = 1; w < kLsWarps; ++w) mn = min(mn, sm.red[0][w]), mx = max(mx, sm.red[1][w]);
                flo = __uint_as_float(ordered_to_float(mn)), fhi = __uint_as_float(ordered_to_float(mx));
            } else {
                const float2 par = a.segpar[seg];
                const float cell_w = 1.0f / par.y;
                flo = par.x + static_cast<float>(sm.info[first_g].x & 0x7fffffffu) * cell_w;
                fhi = par.x + static_cast<float>(sm.info[last_g].y) * cell_w;
            }
            int n_sub = 64;
            while (n_sub < n) n_sub <<= 1;  // 0.5 .. 1 key per sub-bucket, <= kLsMaxSub
            const float scale = static_cast<float>(n_sub) / (fhi - flo);
            bool network = !(isfinite(flo) && isfinite(fhi) && fhi > flo);
            if (!network) {
                uint32_t pk[kLsItems];  // sub-bucket << 16 | arrival rank inside it
#pragma unroll
                for (int q = 0; q < kLsItems; ++q) {
                    if (q * kLsThreads + tid < n) {
                        int sub = linear_cell_finite(k[q], flo, scale, n_sub);
                        if (__uint_as_float(k[q]) != __uint_as_float(k[q])) {  // (the run then takes the network below)
                            sub = 0;
                            sm.has_nan = 1;
                        }
                        pk[q] = (static_cast<uint32_t>(sub) << 16) | atomicAdd(&sm.cnt[sub], 1u);
                    }
                }
                __syncthreads();
                // sum of squares of the counts = the compares the ranking step will make
                if (n_sub >= 16 * kLsThreads)
                    scan_counts<16>(sm, n_sub, tid, lane, warp);
                else if (n_sub >= 8 * kLsThreads)
                    scan_counts<8>(sm, n_sub, tid, lane, warp);
                else if (n_sub >= 4 * kLsThreads)
                    scan_counts<4>(sm, n_sub, tid, lane, warp);
                else if (n_sub >= 2 * kLsThreads)
                    scan_counts<2>(sm, n_sub, tid, lane, warp);
                else
                    scan_counts<1>(sm, n_sub, tid, lane, warp);
                if (tid == 0) sm.cnt[n_sub] = static_cast<uint32_t>(n);
                // badly loaded sub-buckets (clustered values): the all-pairs ranking would cost more than the network
                // (and a run with NaNs: the ranking step below assumes that u + 1 does not wrap and that values map to sub-buckets)
                network = sm.cost > static_cast<uint32_t>(n) * 64u || sm.has_nan != 0;
                __syncthreads();
                if (!network) {
#pragma unroll
                    for (int q = 0; q < kLsItems; ++q)
                        if (q * kLsThreads + tid < n) sm.buf[sm.cnt[pk[q] >> 16] + (pk[q] & 0xffff)] = float_to_ordered(k[q]);
                    if (tid < 4) sm.buf[n + tid] = 0xffffffffu;
                    __syncthreads();
                    // slot order: lanes next to each other read the same few sub-buckets (broadcasts, neighbouring words) and
                    // write neighbouring final positions
                    // (four keys per thread and round: with one, the fused variant had a single load of the sorted A in flight per
                    // thread -- 1125 us against 867 us for the plain sort of 134 M keys)
                    constexpr int kBatch = 4;
                    for (int i0 = tid; i0 < n; i0 += kBatch * kLsThreads) {
                        uint32_t pos[kBatch], val[kBatch];
#pragma unroll
                        for (int v = 0; v < kBatch; ++v) {
                            const int i = i0 + v * kLsThreads;
                            pos[v] = 0xffffffffu;
                            val[v] = 0;
                            if (i < n) {
                                const uint32_t u = sm.buf[i];
                                const uint32_t bits = ordered_to_float(u);
                                const int sub = linear_cell_finite(bits, flo, scale, n_sub);
                                const uint32_t s0 = sm.cnt[sub], c = sm.cnt[sub + 1] - s0, self = static_cast<uint32_t>(i) - s0;
                                // position = keys of the sub-bucket that sort before this one (ties: by slot).  Nearly every
                                // sub-bucket holds <= 4 keys: four unconditional probes.  A probe past the sub-bucket's end meets a
                                // key of a LATER sub-bucket (or the 0xffffffff behind the run), which is larger and counts nothing.
                                // "<= for the slots before mine, < for those behind" is one 64-bit compare of (key, slot) pairs
                                uint32_t rank = s0;
                                const uint32_t* probe = sm.buf + s0;
                                const uint64_t mine = (static_cast<uint64_t>(u) << 32) | self;
#pragma unroll
                                for (uint32_t e = 0; e < 4; ++e) rank += ((static_cast<uint64_t>(probe[e]) << 32) | e) < mine ? 1u : 0u;
                                for (uint32_t e = 4; e < c; ++e) rank += ((static_cast<uint64_t>(probe[e]) << 32) | e) < mine ? 1u : 0u;
                                pos[v] = rank;
                                val[v] = bits;
                            }
                        }
                        if constexpr (DIFF) {
                            float ref[kBatch];
#pragma unroll
                            for (int v = 0; v < kBatch; ++v) ref[v] = pos[v] != 0xffffffffu ? __ldg(ref_run + pos[v]) : 0.0f;
#pragma unroll
                            for (int v = 0; v < kBatch; ++v)
                                if (pos[v] != 0xffffffffu) acc += static_cast<double>(fabsf(ref[v] - __uint_as_float(val[v])));
                        } else {
#pragma unroll
                            for (int v = 0; v < kBatch; ++v)
                                if (pos[v] != 0xffffffffu) dst_run[pos[v]] = val[v];
                        }
                    }
                }
            }
            if (network) {
                int padded = 32;
                while (padded < n) padded <<= 1;
                __syncthreads();
#pragma unroll
                for (int q = 0; q < kLsItems; ++q) {
                    const int i = q * kLsThreads + tid;
                    if (i < padded) sm.buf[i] = i < n ? float_to_ordered(k[q]) : 0xffffffffu;
                }
                __syncthreads();
                bitonic_sort_smem<kLsThreads>(sm.buf, padded, tid);
                for (int i = tid; i < n; i += kLsThreads) emit(i, ordered_to_float(sm.buf[i]));
            }
            __syncthreads();
            for (int i = tid; i <= n_sub; i += kLsThreads) sm.cnt[i] = 0;  // ready for the next run
            if (tid == 0) sm.cost = 0, sm.has_nan = 0;
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
        BGAN_CUDA_OK(cudaFuncSetAttribute(ss_model_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSsMaxCells * 4));
        BGAN_CUDA_OK(cudaFuncSetAttribute(ss_classify_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSsMaxCells * 2 + kSsMaxIds * 4 + kSsMaxHeavy * 8));
        BGAN_CUDA_OK(cudaFuncSetAttribute(ss_scatter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(scatter_smem<true>())));
        BGAN_CUDA_OK(cudaFuncSetAttribute(ss_scatter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(scatter_smem<false>())));
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
        a.starts = nullptr;
        a.idinfo = nullptr;
        a.segpar = nullptr;
        *flag_device = nullptr;
    } else {
        uint32_t* scratch = reinterpret_cast<uint32_t*>(w + l.scratch);
        uint16_t* ids = reinterpret_cast<uint16_t*>(w + l.ids);
        uint16_t* counts = reinterpret_cast<uint16_t*>(w + l.counts);
        uint32_t* bases = reinterpret_cast<uint32_t*>(w + l.bases);
        uint32_t* starts = reinterpret_cast<uint32_t*>(w + l.starts);
        uint16_t* lut = reinterpret_cast<uint16_t*>(w + l.lut);
        uint2* idinfo = reinterpret_cast<uint2*>(w + l.idinfo);
        uint2* heavy = reinterpret_cast<uint2*>(w + l.heavy);
        float2* segpar = reinterpret_cast<float2*>(w + l.segpar);
        ss_model_kernel<<<segments, kModelThreads, p.cells * sizeof(uint32_t), stream>>>(keys, rows, p.buckets, lut, idinfo, heavy, segpar);
        ss_classify_kernel<<<dim3(ceil_div(p.blocks, kSsClsBlocks), segments), kSsThreads, p.cells * 2 + p.ids * 4 + kSsMaxHeavy * 8, stream>>>(
            keys, rows, p.buckets, p.blocks, lut, heavy, segpar, ids, counts);
        ss_scan_kernel<<<segments, kModelThreads, 0, stream>>>(counts, rows, p.ids, p.blocks, idinfo, starts, bases, ctrl);
        if (p.buckets <= 512)
            ss_scatter_kernel<true><<<dim3(p.blocks, segments), kScThreads, scatter_smem<true>(), stream>>>(keys, ids, rows, p.ids, p.blocks,
                                                                                                           counts, bases, scratch, ctrl);
        else
            ss_scatter_kernel<false><<<dim3(p.blocks, segments), kScThreads, scatter_smem<false>(), stream>>>(keys, ids, rows, p.ids, p.blocks,
                                                                                                             counts, bases, scratch, ctrl);
        a.src = scratch;
        a.n_ids = p.ids;
        a.starts = starts;
        a.idinfo = idinfo;
        a.segpar = segpar;
        *flag_device = ctrl;
    }
    const int batches = p.direct ? segments : segments * ceil_div(p.ids, kLsBatchIds);
    const int grid_ls = std::max(1, std::min(batches, 2 * device_sm_count()));
    if (da.ref != nullptr)
        ss_local_sort_kernel<true><<<grid_ls, kLsThreads, sizeof(LocalSmem), stream>>>(a);
    else
        ss_local_sort_kernel<false><<<grid_ls, kLsThreads, sizeof(LocalSmem), stream>>>(a);
    BGAN_LAUNCH_OK();
    return 0;
}

}  // namespace bgan
