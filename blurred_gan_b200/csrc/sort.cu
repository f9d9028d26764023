// Segmented LSD radix sort of fp32 keys for sm_100a -- the np.sort(proj, axis=0) of the reference's
// sliced_wasserstein.py:47-48, one segment per projection direction (direction-major layout).  Since the end of round 2 this is
// the sort of segments longer than 2^21 keys and the fall-back of the value-bucket sort (sort_sample.cu), which takes the rest;
// the entry points at the bottom of this file choose.
//
// 4 passes of 8-bit digits over order-preserving uint32 images of the floats.  Per pass:
//   hist    : per-(segment, tile) digit counts                       -> counts[seg][tile][digit]
//   scan    : exclusive prefix over (digit-major, tile-minor)        -> offsets[seg][tile][digit] (two small kernels)
//   scatter : stable rank inside the tile (lanes with equal digits found by eight ballots -- match_any for the top
//             byte, which takes few distinct values -- plus per-warp counters), keys regrouped by
//             digit in shared memory, then written as coalesced runs to their global positions.
// A "one sweep" variant (all four histograms from one read of the keys, decoupled look-back in the scatter: 36 instead of
// 48 bytes per key) was built and measured in round 2 and is NOT here: the histogram is bound by shared-memory atomics
// (3 per key: 365 us against 4 x 135 us for the per-pass kernels), the look-back scatter took 520 - 560 us per pass
// against ~400, 54 instead of 61 Gkeys/s at 512 x 262144 keys (profiles/r02_sort.md).
// Tile = 256 threads x 16 keys.  The float<->uint transform is fused into the first load and the
// last store, so the sort runs in place from the caller's point of view (ping-pong through ws).
#include <algorithm>
#include <cstdlib>

#include "sort.cuh"

namespace bgan {
namespace {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kItems = 16;
constexpr int kTile = kSortThreads * kItems;  // 4096 keys
constexpr int kRadix = 256;

// A single statement that "uses" all 16 loaded keys: nothing that consumes one of them can be scheduled before ALL the
// loads have been issued.  Without it the first pass's float -> ordered-uint transforms were interleaved with the loads
// (SASS: 16 LDG spread over 110 instructions), each transform waiting for its load, a handful of loads in flight per warp:
// the first histogram pass ran at 2.8 TB/s (192 us) where the later ones, which have no transform, reach 5.5 TB/s (97 us).
__device__ __forceinline__ void pin_loads(uint32_t (&k)[kItems]) {
    static_assert(kItems == 16, "pin_loads lists 16 registers");
    asm volatile("" : "+r"(k[0]), "+r"(k[1]), "+r"(k[2]), "+r"(k[3]), "+r"(k[4]), "+r"(k[5]), "+r"(k[6]), "+r"(k[7]), "+r"(k[8]), "+r"(k[9]),
                      "+r"(k[10]), "+r"(k[11]), "+r"(k[12]), "+r"(k[13]), "+r"(k[14]), "+r"(k[15]));
}

// counts[(seg * tiles + tile) * 256 + digit]: one coalesced 1 KB row per tile, for the histogram's write and the
// scatter's read alike
// lanes of the warp that hold the same 8-bit digit (bit 8 set = padding lane, matches nobody real).
// MATCH: one match_any -- fast when a 32-key row holds a handful of distinct digits (the top byte of the keys: sign and
// exponent), but its cost grows with the number of distinct values (measured: ~240 cycles per row for uniform digits).
// Otherwise eight ballots, one per digit bit: constant time.
// One digit bit: peers &= lanes whose bit equals mine.  Written as PTX so that it stays four instructions (test, vote and
// two predicated logic operations); the C++ form `peers &= bit ? m : ~m` compiled to six or seven per bit, and the ranking
// loop is what bounds the scatter kernel (instruction-bound: profiles/r01e_swd.md).
#define BGAN_MATCH_BIT(B)                                  \
    "and.b32 t, %1, " #B ";\n\t"                           \
    "setp.ne.u32 p, t, 0;\n\t"                             \
    "vote.sync.ballot.b32 m, p, 0xffffffff;\n\t"           \
    "@!p not.b32 m, m;\n\t"                                \
    "and.b32 %0, %0, m;\n\t"

template <bool MATCH, bool PADDED>
__device__ __forceinline__ uint32_t same_digit_lanes(uint32_t d) {
    if constexpr (MATCH) {
        return __match_any_sync(0xffffffffu, d);
    } else {
        uint32_t peers;
        if constexpr (PADDED) {
            peers = __ballot_sync(0xffffffffu, d < 256u);
            if (d >= 256u) peers = ~peers;
        } else {
            peers = 0xffffffffu;
        }
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            ".reg .b32 m, t;\n\t" BGAN_MATCH_BIT(1) BGAN_MATCH_BIT(2) BGAN_MATCH_BIT(4) BGAN_MATCH_BIT(8) BGAN_MATCH_BIT(16) BGAN_MATCH_BIT(32)
                BGAN_MATCH_BIT(64) BGAN_MATCH_BIT(128) "}"
            : "+r"(peers)
            : "r"(d));
        return peers;
    }
}

// One shared-memory atomic per key into the warp's own histogram.  (Round 1 aggregated the top byte -- sign and exponent, a
// handful of values per 32 keys -- with match_any first; measured in round 2, the sixteen match_any per thread cost more
// than the conflicts they avoid: 209 us against ~100 us per pass of 134 M keys, 60.9 -> 64.3 Gkeys/s for the whole sort.)
template <bool FIRST>
__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const uint32_t* __restrict__ keys, int64_t rows, int tiles, int shift,
                                                                  uint32_t* __restrict__ counts) {
    __shared__ uint32_t s_cnt[kSortWarps][kRadix];
    const int tile = blockIdx.x, seg = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < kSortWarps * kRadix / 4; i += kSortThreads) reinterpret_cast<uint4*>(&s_cnt[0][0])[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    const uint32_t* src = keys + static_cast<size_t>(seg) * rows + static_cast<size_t>(tile) * kTile;
    const int valid = (rows - static_cast<int64_t>(tile) * kTile < kTile) ? static_cast<int>(rows - static_cast<int64_t>(tile) * kTile) : kTile;
    // all 16 loads of the thread first: issued between the shared-memory atomics they were serialised (18 registers, one
    // load in flight per warp, "long scoreboard" 41 - 51 per issue in ncu: 2.3 TB/s); up front they are 16 x 128 B per warp
    uint32_t key[kItems];
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        const int local = i * kSortThreads + tid;
        key[i] = local < valid ? __ldg(src + local) : 0u;
    }
    pin_loads(key);
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        const uint32_t k = FIRST ? float_to_ordered(key[i]) : key[i];
        if (i * kSortThreads + tid < valid) atomicAdd(&s_cnt[warp][(k >> shift) & 0xff], 1u);
    }
    __syncthreads();
    uint32_t total = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) total += s_cnt[w][tid];
    counts[(static_cast<size_t>(seg) * tiles + tile) * kRadix + tid] = total;
}

// Offsets: counts[seg][tile][digit] -> global start of every (digit, tile) run, i.e. the exclusive prefix in (digit-major,
// tile-minor) order, written to a second array of the same shape.  Two small kernels over (chunk of tiles, segment) CTAs of 256 threads
// (thread = digit; every access is a coalesced 1 KB row):
//   partial : part[seg][chunk][digit] = sum of the chunk's counts
//   offsets : digit totals -> exclusive scan over the digits -> start of this chunk -> walk the chunk
constexpr int kScanChunkTiles = 16;  // tiles per chunk (fully unrolled loads)
constexpr int kMaxScanChunks = 1024;

__global__ void __launch_bounds__(kRadix) radix_scan_partial_kernel(const uint32_t* __restrict__ counts, int tiles, int chunks,
                                                                    uint32_t* __restrict__ part) {
    const int chunk = blockIdx.x, seg = blockIdx.y, d = threadIdx.x;
    const int per = (tiles + chunks - 1) / chunks;
    const int t0 = min(chunk * per, tiles), t1 = min(t0 + per, tiles);
    const uint32_t* c = counts + static_cast<size_t>(seg) * tiles * kRadix + d;
    uint32_t sum = 0;
#pragma unroll 8
    for (int t = t0; t < t1; ++t) sum += __ldg(c + static_cast<size_t>(t) * kRadix);
    part[(static_cast<size_t>(seg) * chunks + chunk) * kRadix + d] = sum;
}

__global__ void __launch_bounds__(kRadix) radix_scan_offsets_kernel(const uint32_t* __restrict__ counts, int tiles, int chunks,
                                                                    const uint32_t* __restrict__ part, uint32_t* __restrict__ offsets) {
    __shared__ uint32_t s_warp[kRadix / 32];
    const int chunk = blockIdx.x, seg = blockIdx.y, d = threadIdx.x, lane = d & 31, warp = d >> 5;
    const uint32_t* p = part + static_cast<size_t>(seg) * chunks * kRadix + d;
    uint32_t total = 0, before = 0;  // this digit: all chunks / the chunks in front of mine
    for (int ch = 0; ch < chunks; ++ch) {
        const uint32_t v = __ldg(p + static_cast<size_t>(ch) * kRadix);
        total += v;
        if (ch < chunk) before += v;
    }
    uint32_t incl = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t run = incl - total + before;
    for (int w = 0; w < warp; ++w) run += s_warp[w];
    const int per = (tiles + chunks - 1) / chunks;
    const int t0 = min(chunk * per, tiles), t1 = min(t0 + per, tiles);
    // out of place (reads and writes provably do not alias: all loads of the chunk are in flight together)
    const uint32_t* c = counts + static_cast<size_t>(seg) * tiles * kRadix + d;
    uint32_t* o = offsets + static_cast<size_t>(seg) * tiles * kRadix + d;
    for (int tb = t0; tb < t1; tb += 8) {
        uint32_t v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = tb + j < t1 ? __ldg(c + static_cast<size_t>(tb + j) * kRadix) : 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (tb + j < t1) o[static_cast<size_t>(tb + j) * kRadix] = run;
            run += v[j];
        }
    }
}

// Scatter of one tile.  Ranking: per 32-key row, lanes with the same digit find each other with match_any; the lowest one
// bumps the warp's counter of that digit (one shared-memory atomic per distinct digit) and hands the old value to its
// peers.  Keys are then regrouped by digit in shared memory and leave as runs of consecutive addresses.
struct ScatterSmem {
    uint32_t cnt[kSortWarps][kRadix];  // per-warp digit counts, then: first local slot of the warp's keys of that digit
    uint32_t gdelta[kRadix];           // global position of the tile's first key of a digit - its first local slot
    uint32_t keys[kTile];
    uint32_t scan[kSortWarps];
    double diff[kSortWarps];
};

// DIFF (only with LAST): the keys are not written; every key meets the element of `ref` that sits at its final position
// and sum |ref - key| is accumulated instead (sliced_wasserstein.py:47-50: sort both projections, mean |A - B|).
template <bool FIRST, bool LAST, bool FULL, bool MATCH, bool DIFF>
__device__ __forceinline__ void scatter_tile(ScatterSmem& sm, const uint32_t* __restrict__ keys_in, uint32_t* __restrict__ keys_out,
                                             int64_t rows, int tiles, int shift, const uint32_t* __restrict__ offsets, const DiffArgs& da) {
    auto& s_cnt = sm.cnt;
    auto& s_gdelta = sm.gdelta;
    auto& s_keys = sm.keys;
    auto& s_scan = sm.scan;
    const int tile = blockIdx.x, seg = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (int i = tid; i < kSortWarps * kRadix / 4; i += kSortThreads) reinterpret_cast<uint4*>(&s_cnt[0][0])[i] = make_uint4(0, 0, 0, 0);
    const uint32_t gbase = offsets[(static_cast<size_t>(seg) * tiles + tile) * kRadix + tid];
    __syncthreads();

    const int64_t tile_base = static_cast<int64_t>(tile) * kTile;
    const uint32_t* src = keys_in + static_cast<size_t>(seg) * rows + tile_base;
    uint32_t* dst = keys_out + static_cast<size_t>(seg) * rows;
    const int valid = FULL ? kTile : static_cast<int>(rows - tile_base);  // (FULL == false: 1 .. kTile - 1)
    const int warp_base = warp * (kItems * 32);  // each warp owns 512 consecutive keys -> stability
    uint32_t* my_cnt = s_cnt[warp];

    uint32_t key[kItems];
    uint32_t rank2[kItems / 2];  // two 16-bit ranks per register (a rank is < 4096)
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        const int local = warp_base + i * 32 + lane;
        const bool ok = FULL || local < valid;
        key[i] = ok ? __ldg(src + local) : (FIRST ? 0x7fffffffu : 0xffffffffu);  // (padding: the largest key either way)
    }
    pin_loads(key);
    if constexpr (FIRST) {
#pragma unroll
        for (int i = 0; i < kItems; ++i) key[i] = float_to_ordered(key[i]);
    }
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        const bool ok = FULL || warp_base + i * 32 + lane < valid;
        // lanes holding padding keys vote with digit 256 so that they never disturb a real digit
        const uint32_t d = ok ? (key[i] >> shift) & 0xff : 256u;
        const uint32_t peers = same_digit_lanes<MATCH, !FULL>(d);
        const int leader = __ffs(peers) - 1;
        // the lowest lane of every group bumps the warp's counter of that digit (groups hold different digits, so plain
        // read + write; the __syncwarp makes it visible to whichever lane leads that digit in the next row)
        uint32_t before = 0;
        if (lane == leader && ok) {
            before = my_cnt[d];
            my_cnt[d] = before + __popc(peers);
        }
        __syncwarp();
        before = __shfl_sync(0xffffffffu, before, leader);
        const uint32_t r = before + __popc(peers & lt_mask);
        rank2[i >> 1] = (i & 1) ? (rank2[i >> 1] | (r << 16)) : r;
    }
    __syncthreads();

    // per digit (thread = digit): exclusive prefix over warps, tile total, then the exclusive scan of the totals
    uint32_t digit_total = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) digit_total += s_cnt[w][tid];
    uint32_t incl = digit_total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    uint32_t tile_start = incl - digit_total;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w)
        if (w < warp) tile_start += s_scan[w];
    s_gdelta[tid] = gbase - tile_start;
    {
        uint32_t run = tile_start;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            const uint32_t c = s_cnt[w][tid];
            s_cnt[w][tid] = run;
            run += c;
        }
    }
    __syncthreads();

    // regroup by digit in shared memory
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        if (FULL || warp_base + i * 32 + lane < valid)
            s_keys[my_cnt[(key[i] >> shift) & 0xff] + ((rank2[i >> 1] >> (16 * (i & 1))) & 0xffff)] = key[i];
    }
    __syncthreads();

    // runs of consecutive addresses out to global
    if constexpr (DIFF) {
        const float* ref = da.ref + static_cast<size_t>(seg) * rows;
        uint32_t kk[kItems];
        float rr[kItems];
#pragma unroll
        for (int i = 0; i < kItems; ++i) {
            const int slot = i * kSortThreads + tid;
            const bool ok = FULL || slot < valid;
            kk[i] = ok ? s_keys[slot] : 0u;
            rr[i] = ok ? __ldg(ref + s_gdelta[(kk[i] >> shift) & 0xff] + slot) : 0.0f;
        }
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < kItems; ++i) {
            const int slot = i * kSortThreads + tid;
            if (FULL || slot < valid) acc += static_cast<double>(fabsf(rr[i] - __uint_as_float(ordered_to_float(kk[i]))));
        }
        acc = warp_sum(acc);
        if (lane == 0) sm.diff[warp] = acc;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < kSortWarps; ++w) t += sm.diff[w];
            atomicAdd(da.partial + static_cast<size_t>((seg + da.seg_offset) / da.segs_per_sum) * kDiffSlots + (tile % kDiffSlots), t);
        }
    } else {
#pragma unroll
        for (int i = 0; i < kItems; ++i) {
            const int slot = i * kSortThreads + tid;
            if (FULL || slot < valid) {
                const uint32_t k = s_keys[slot];
                dst[s_gdelta[(k >> shift) & 0xff] + slot] = LAST ? ordered_to_float(k) : k;
            }
        }
    }
}

template <bool FIRST, bool LAST, bool MATCH, bool DIFF>
__global__ void __launch_bounds__(kSortThreads, 4) radix_scatter_kernel(const uint32_t* __restrict__ keys_in, uint32_t* __restrict__ keys_out,
                                                                     int64_t rows, int tiles, int shift,
                                                                     const uint32_t* __restrict__ offsets, const DiffArgs da) {
    __shared__ ScatterSmem sm;
    // complete tiles run without any bounds logic; only the last tile of a segment can be ragged
    if ((static_cast<int64_t>(blockIdx.x) + 1) * kTile <= rows)
        scatter_tile<FIRST, LAST, true, MATCH, DIFF>(sm, keys_in, keys_out, rows, tiles, shift, offsets, da);
    else
        scatter_tile<FIRST, LAST, false, MATCH, DIFF>(sm, keys_in, keys_out, rows, tiles, shift, offsets, da);
}

// sums[s] += the kDiffSlots partial sums of sum s (one warp per sum)
__global__ void diff_finish_kernel(const double* __restrict__ partial, int n_sums, double* __restrict__ sums) {
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (s >= n_sums) return;
    double t = 0.0;
    for (int i = lane; i < kDiffSlots; i += 32) t += partial[static_cast<size_t>(s) * kDiffSlots + i];
    t = warp_sum(t);
    if (lane == 0) sums[s] += t;
}

}  // namespace
}  // namespace bgan

using namespace bgan;

static int scan_chunks(int64_t tiles) {
    return static_cast<int>(std::min<int64_t>(kMaxScanChunks, std::max<int64_t>(1, ceil_div64(tiles, kScanChunkTiles))));
}

static size_t lsd_workspace_bytes(int64_t rows, int segments) {
    const int64_t tiles = ceil_div64(rows, kTile);
    const size_t keys = static_cast<size_t>(rows) * segments * sizeof(uint32_t);
    const size_t counts = static_cast<size_t>(segments) * kRadix * tiles * sizeof(uint32_t);
    const size_t part = static_cast<size_t>(segments) * kRadix * scan_chunks(tiles) * sizeof(uint32_t);
    return ((keys + 255) & ~size_t(255)) + 2 * counts + part;  // keys | counts | offsets | chunk partials
}

// Which sort runs: the value-bucket sort (sort_sample.cu) for segments it holds, else the radix passes below.
// BGAN_SORT_ALGO=lsd forces the radix sort (A/B measurements).
static bool use_sample_sort(int64_t rows) {
    const char* algo = getenv("BGAN_SORT_ALGO");
    if (algo && algo[0] == 'l') return false;
    return sample_sort_applicable(rows);
}

extern "C" size_t bgan_swd_sort_workspace_bytes(int64_t rows, int segments) {
    if (rows <= 0 || segments <= 0) return 0;
    // the radix sort is the fall-back of the sample sort and runs in the same workspace
    return std::max(lsd_workspace_bytes(rows, segments), sample_sort_applicable(rows) ? sample_sort_workspace_bytes(rows, segments) : size_t(0));
}

// The four passes over `segments` runs of `rows` keys at keys_device; with `da.ref` the last pass does not write the keys
// but accumulates sum |ref - key| (see DiffArgs).
static int sort_passes(float* keys_device, int64_t rows, int segments, void* ws, cudaStream_t stream, const DiffArgs& da) {
    const int tiles = static_cast<int>(ceil_div64(rows, kTile));
    const size_t keys_bytes = (static_cast<size_t>(rows) * segments * sizeof(uint32_t) + 255) & ~size_t(255);
    uint32_t* a = reinterpret_cast<uint32_t*>(keys_device);
    uint32_t* b = static_cast<uint32_t*>(ws);
    uint32_t* counts = reinterpret_cast<uint32_t*>(static_cast<char*>(ws) + keys_bytes);
    const int chunks = scan_chunks(tiles);
    uint32_t* offsets = counts + static_cast<size_t>(segments) * kRadix * tiles;
    uint32_t* part = offsets + static_cast<size_t>(segments) * kRadix * tiles;
    dim3 scan_grid(chunks, segments);
    auto scan = [&]() {
        radix_scan_partial_kernel<<<scan_grid, kRadix, 0, stream>>>(counts, tiles, chunks, part);
        radix_scan_offsets_kernel<<<scan_grid, kRadix, 0, stream>>>(counts, tiles, chunks, part, offsets);
    };
    dim3 grid(tiles, segments);
    const DiffArgs none{nullptr, nullptr, 0, 1};

    // pass 0: a -> b (float -> ordered uint on load)
    radix_hist_kernel<true><<<grid, kSortThreads, 0, stream>>>(a, rows, tiles, 0, counts);
    scan();
    radix_scatter_kernel<true, false, false, false><<<grid, kSortThreads, 0, stream>>>(a, b, rows, tiles, 0, offsets, none);
    // pass 1: b -> a
    radix_hist_kernel<false><<<grid, kSortThreads, 0, stream>>>(b, rows, tiles, 8, counts);
    scan();
    radix_scatter_kernel<false, false, false, false><<<grid, kSortThreads, 0, stream>>>(b, a, rows, tiles, 8, offsets, none);
    // pass 2: a -> b
    radix_hist_kernel<false><<<grid, kSortThreads, 0, stream>>>(a, rows, tiles, 16, counts);
    scan();
    radix_scatter_kernel<false, false, false, false><<<grid, kSortThreads, 0, stream>>>(a, b, rows, tiles, 16, offsets, none);
    // pass 3: b -> a (ordered uint -> float on store), or b -> sum |ref - key|
    radix_hist_kernel<false><<<grid, kSortThreads, 0, stream>>>(b, rows, tiles, 24, counts);
    scan();
    if (da.ref != nullptr)
        radix_scatter_kernel<false, true, true, true><<<grid, kSortThreads, 0, stream>>>(b, a, rows, tiles, 24, offsets, da);
    else
        radix_scatter_kernel<false, true, true, false><<<grid, kSortThreads, 0, stream>>>(b, a, rows, tiles, 24, offsets, none);
    BGAN_LAUNCH_OK();
    return 0;
}

// One array of segments: the sample sort when it applies, with the radix passes as its fall-back.  The sample sort reports a
// bucket that does not fit shared memory through a device flag and then has written nothing, so the fall-back starts from
// the untouched keys; reading the flag costs one 4-byte device -> host copy (and a stream synchronisation) per call for
// segments of more than 8192 keys.  A stream that is being captured cannot be synchronised: it takes the radix passes.
static thread_local int g_sort_path = 0;
extern "C" int bgan_swd_sort_last_path(void) { return g_sort_path; }

static int sort_one(float* keys_device, int64_t rows, int segments, void* ws, cudaStream_t stream, const DiffArgs& da) {
    g_sort_path = 1;
    if (use_sample_sort(rows)) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        BGAN_CUDA_OK(cudaStreamIsCapturing(stream, &cap));
        const uint32_t* flag_device = nullptr;
        if (cap == cudaStreamCaptureStatusNone || rows <= 8192) {  // (<= 8192 keys: one CTA per segment, no flag, no read-back)
            const int rc = sample_sort_launch(keys_device, rows, segments, ws, stream, da, &flag_device);
            if (rc) return rc;
            g_sort_path = 2;
            if (flag_device == nullptr) return 0;
            uint32_t flag = 0;
            BGAN_CUDA_OK(cudaMemcpyAsync(&flag, flag_device, sizeof(flag), cudaMemcpyDeviceToHost, stream));
            BGAN_CUDA_OK(cudaStreamSynchronize(stream));
            if (flag == 0) return 0;
            g_sort_path = 3;
        }
    }
    return sort_passes(keys_device, rows, segments, ws, stream, da);
}

extern "C" int bgan_swd_sort_f32(float* keys_device, int64_t rows, int segments, void* ws, size_t ws_bytes, void* stream_) {
    BGAN_REQUIRE(keys_device && rows > 0 && segments > 0, BGAN_E_BADARG, "bgan_swd_sort_f32: bad arguments");
    BGAN_REQUIRE(rows < (int64_t(1) << 31), BGAN_E_BADARG, "bgan_swd_sort_f32: at most 2^31-1 rows per segment");
    BGAN_REQUIRE(segments <= 65535, BGAN_E_BADARG, "bgan_swd_sort_f32: at most 65535 segments per call");
    BGAN_REQUIRE(ws && ws_bytes >= bgan_swd_sort_workspace_bytes(rows, segments), BGAN_E_WORKSPACE,
                 "bgan_swd_sort_f32: workspace of %zu bytes required, got %zu", bgan_swd_sort_workspace_bytes(rows, segments), ws_bytes);
    return sort_one(keys_device, rows, segments, ws, static_cast<cudaStream_t>(stream_), DiffArgs{nullptr, nullptr, 0, 1});
}

static size_t diff_partial_bytes(int n_sums) { return static_cast<size_t>(n_sums) * kDiffSlots * sizeof(double); }

extern "C" size_t bgan_swd_sort_diff_workspace_bytes(int64_t rows, int segments, int n_sums) {
    if (rows <= 0 || segments <= 0 || n_sums <= 0) return 0;
    return ((bgan_swd_sort_workspace_bytes(rows, segments) + 255) & ~size_t(255)) + diff_partial_bytes(n_sums);
}

extern "C" int bgan_swd_sort_diff_f32(float* a_device, float* b_device, int64_t rows, int segments, int seg_offset, int segs_per_sum, int n_sums,
                                      double* sums_device, void* ws, size_t ws_bytes, void* stream_) {
    BGAN_REQUIRE(a_device && b_device && sums_device && rows > 0 && segments > 0, BGAN_E_BADARG, "bgan_swd_sort_diff_f32: bad arguments");
    BGAN_REQUIRE(rows < (int64_t(1) << 31) && segments <= 65535, BGAN_E_BADARG, "bgan_swd_sort_diff_f32: at most 2^31-1 rows, 65535 segments");
    BGAN_REQUIRE(seg_offset >= 0 && segs_per_sum > 0 && n_sums > 0 && (seg_offset + segments - 1) / segs_per_sum < n_sums, BGAN_E_BADARG,
                 "bgan_swd_sort_diff_f32: segments %d .. %d do not map into %d sums of %d segments", seg_offset, seg_offset + segments - 1, n_sums,
                 segs_per_sum);
    BGAN_REQUIRE(ws && ws_bytes >= bgan_swd_sort_diff_workspace_bytes(rows, segments, n_sums), BGAN_E_WORKSPACE,
                 "bgan_swd_sort_diff_f32: workspace of %zu bytes required, got %zu", bgan_swd_sort_diff_workspace_bytes(rows, segments, n_sums),
                 ws_bytes);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    double* partial = reinterpret_cast<double*>(static_cast<char*>(ws) + ((bgan_swd_sort_workspace_bytes(rows, segments) + 255) & ~size_t(255)));
    int rc = sort_one(a_device, rows, segments, ws, stream, DiffArgs{nullptr, nullptr, 0, 1});  // A: sorted in place
    if (rc) return rc;
    BGAN_CUDA_OK(cudaMemsetAsync(partial, 0, diff_partial_bytes(n_sums), stream));
    rc = sort_one(b_device, rows, segments, ws, stream, DiffArgs{a_device, partial, seg_offset, segs_per_sum});  // B: never written back
    if (rc) return rc;
    diff_finish_kernel<<<ceil_div(n_sums, 4), 128, 0, stream>>>(partial, n_sums, sums_device);
    BGAN_LAUNCH_OK();
    return 0;
}
