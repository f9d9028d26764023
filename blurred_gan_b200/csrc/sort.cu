// Segmented LSD radix sort of fp32 keys for sm_100a -- the np.sort(proj, axis=0) of the reference's
// sliced_wasserstein.py:47-48, one segment per projection direction (direction-major layout).
//
// 4 passes of 8-bit digits over order-preserving uint32 images of the floats.  Per pass:
//   hist    : per-(segment, tile) digit counts                       -> counts[seg][digit][tile]
//   scan    : exclusive prefix over (digit-major, tile-minor)        -> same array, in place
//   scatter : stable rank inside the tile (warp match_any + per-warp counters), keys regrouped by
//             digit in shared memory, then written as coalesced runs to their global positions.
// Tile = 256 threads x 16 keys.  The float<->uint transform is fused into the first load and the
// last store, so the sort runs in place from the caller's point of view (ping-pong through ws).
#include "common.cuh"

namespace bgan {
namespace {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kItems = 16;
constexpr int kTile = kSortThreads * kItems;  // 4096 keys
constexpr int kRadix = 256;

__device__ __forceinline__ uint32_t float_to_ordered(uint32_t b) { return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u); }
__device__ __forceinline__ uint32_t ordered_to_float(uint32_t u) { return u ^ ((u >> 31) ? 0x80000000u : 0xffffffffu); }

template <bool FIRST>
__device__ __forceinline__ uint32_t load_key(const uint32_t* __restrict__ p) {
    const uint32_t v = __ldg(p);
    return FIRST ? float_to_ordered(v) : v;
}

// counts[(seg * 256 + digit) * tiles + tile]
template <bool FIRST>
__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const uint32_t* __restrict__ keys, int64_t rows, int tiles, int shift,
                                                                  uint32_t* __restrict__ counts) {
    __shared__ uint32_t s_cnt[kSortWarps][kRadix];
    const int tile = blockIdx.x, seg = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < kSortWarps * kRadix; i += kSortThreads) (&s_cnt[0][0])[i] = 0;
    __syncthreads();
    const uint32_t* src = keys + static_cast<size_t>(seg) * rows;
    const int64_t base = static_cast<int64_t>(tile) * kTile;
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        const int64_t idx = base + i * kSortThreads + tid;
        if (idx < rows) atomicAdd(&s_cnt[warp][(load_key<FIRST>(src + idx) >> shift) & 0xff], 1u);
    }
    __syncthreads();
    if (tid < kRadix) {
        uint32_t total = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) total += s_cnt[w][tid];
        counts[(static_cast<size_t>(seg) * kRadix + tid) * tiles + tile] = total;
    }
}

// one CTA per segment: exclusive scan of counts[seg][0 .. 256*tiles)
__global__ void __launch_bounds__(1024) radix_scan_kernel(uint32_t* __restrict__ counts, int tiles) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    uint32_t* c = counts + static_cast<size_t>(blockIdx.x) * kRadix * tiles;
    const int n = kRadix * tiles;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (n + 1023) / 1024;
    const int lo = tid * per, hi = min(n, lo + per);
    uint32_t local = 0;
    for (int i = lo; i < hi; ++i) local += c[i];
    // block-wide exclusive scan of `local`
    uint32_t incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t v = s_warp[lane], iv = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, iv, o);
            if (lane >= o) iv += t;
        }
        s_warp[lane] = iv - v;
    }
    __syncthreads();
    uint32_t run = s_warp[warp] + incl - local;
    for (int i = lo; i < hi; ++i) {
        const uint32_t v = c[i];
        c[i] = run;
        run += v;
    }
    (void)s_carry;
}

template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(const uint32_t* __restrict__ keys_in, uint32_t* __restrict__ keys_out,
                                                                     int64_t rows, int tiles, int shift,
                                                                     const uint32_t* __restrict__ offsets) {
    __shared__ uint32_t s_cnt[kSortWarps][kRadix];   // per-warp digit counts, then per-warp exclusive prefixes
    __shared__ uint32_t s_tile_start[kRadix];        // first local slot of each digit inside the tile
    __shared__ uint32_t s_gbase[kRadix];             // global position of the tile's first key of each digit
    __shared__ uint32_t s_keys[kTile];
    __shared__ uint32_t s_scan[kSortWarps];

    const int tile = blockIdx.x, seg = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (int i = tid; i < kSortWarps * kRadix; i += kSortThreads) (&s_cnt[0][0])[i] = 0;
    if (tid < kRadix) s_gbase[tid] = offsets[(static_cast<size_t>(seg) * kRadix + tid) * tiles + tile];
    __syncthreads();

    const uint32_t* src = keys_in + static_cast<size_t>(seg) * rows;
    uint32_t* dst = keys_out + static_cast<size_t>(seg) * rows;
    const int64_t tile_base = static_cast<int64_t>(tile) * kTile;
    const int valid = (rows - tile_base < kTile) ? static_cast<int>(rows - tile_base) : kTile;
    const int warp_base = warp * (kItems * 32);  // each warp owns 512 consecutive keys -> stability

    uint32_t key[kItems];
    uint16_t rank[kItems];
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        const int local = warp_base + i * 32 + lane;
        const bool ok = local < valid;
        key[i] = ok ? load_key<FIRST>(src + tile_base + local) : 0xffffffffu;
        const uint32_t d = (key[i] >> shift) & 0xff;
        // lanes holding padding keys vote with digit 256 so that they never disturb a real digit
        const uint32_t peers = __match_any_sync(0xffffffffu, ok ? d : 256u);
        uint32_t before = 0;
        if (ok) before = s_cnt[warp][d];
        __syncwarp();
        if (ok && (peers & lt_mask) == 0) s_cnt[warp][d] = before + __popc(peers);
        __syncwarp();
        rank[i] = static_cast<uint16_t>(before + __popc(peers & lt_mask));
    }
    __syncthreads();

    // per digit: exclusive prefix over warps, tile total
    uint32_t digit_total = 0;
    if (tid < kRadix) {
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            const uint32_t c = s_cnt[w][tid];
            s_cnt[w][tid] = digit_total;
            digit_total += c;
        }
    }
    // exclusive scan of digit totals over the 256 digits (threads 0..255 = all threads)
    uint32_t incl = digit_total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    uint32_t warp_off = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w)
        if (w < warp) warp_off += s_scan[w];
    if (tid < kRadix) s_tile_start[tid] = warp_off + incl - digit_total;
    __syncthreads();

    // regroup by digit in shared memory
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        const int local = warp_base + i * 32 + lane;
        if (local < valid) {
            const uint32_t d = (key[i] >> shift) & 0xff;
            s_keys[s_tile_start[d] + s_cnt[warp][d] + rank[i]] = key[i];
        }
    }
    __syncthreads();

    // coalesced runs out to global
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        const int slot = i * kSortThreads + tid;
        if (slot < valid) {
            const uint32_t k = s_keys[slot];
            const uint32_t d = (k >> shift) & 0xff;
            dst[s_gbase[d] + (slot - s_tile_start[d])] = LAST ? ordered_to_float(k) : k;
        }
    }
}

}  // namespace
}  // namespace bgan

using namespace bgan;

extern "C" size_t bgan_swd_sort_workspace_bytes(int64_t rows, int segments) {
    if (rows <= 0 || segments <= 0) return 0;
    const int64_t tiles = ceil_div64(rows, kTile);
    const size_t keys = static_cast<size_t>(rows) * segments * sizeof(uint32_t);
    const size_t counts = static_cast<size_t>(segments) * kRadix * tiles * sizeof(uint32_t);
    return ((keys + 255) & ~size_t(255)) + counts;
}

extern "C" int bgan_swd_sort_f32(float* keys_device, int64_t rows, int segments, void* ws, size_t ws_bytes, void* stream_) {
    BGAN_REQUIRE(keys_device && rows > 0 && segments > 0, BGAN_E_BADARG, "bgan_swd_sort_f32: bad arguments");
    BGAN_REQUIRE(rows < (int64_t(1) << 31), BGAN_E_BADARG, "bgan_swd_sort_f32: at most 2^31-1 rows per segment");
    BGAN_REQUIRE(segments <= 65535, BGAN_E_BADARG, "bgan_swd_sort_f32: at most 65535 segments per call");
    BGAN_REQUIRE(ws && ws_bytes >= bgan_swd_sort_workspace_bytes(rows, segments), BGAN_E_WORKSPACE,
                 "bgan_swd_sort_f32: workspace of %zu bytes required, got %zu", bgan_swd_sort_workspace_bytes(rows, segments), ws_bytes);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int tiles = static_cast<int>(ceil_div64(rows, kTile));
    const size_t keys_bytes = (static_cast<size_t>(rows) * segments * sizeof(uint32_t) + 255) & ~size_t(255);
    uint32_t* a = reinterpret_cast<uint32_t*>(keys_device);
    uint32_t* b = static_cast<uint32_t*>(ws);
    uint32_t* counts = reinterpret_cast<uint32_t*>(static_cast<char*>(ws) + keys_bytes);
    dim3 grid(tiles, segments);

    // pass 0: a -> b (float -> ordered uint on load)
    radix_hist_kernel<true><<<grid, kSortThreads, 0, stream>>>(a, rows, tiles, 0, counts);
    radix_scan_kernel<<<segments, 1024, 0, stream>>>(counts, tiles);
    radix_scatter_kernel<true, false><<<grid, kSortThreads, 0, stream>>>(a, b, rows, tiles, 0, counts);
    // pass 1: b -> a
    radix_hist_kernel<false><<<grid, kSortThreads, 0, stream>>>(b, rows, tiles, 8, counts);
    radix_scan_kernel<<<segments, 1024, 0, stream>>>(counts, tiles);
    radix_scatter_kernel<false, false><<<grid, kSortThreads, 0, stream>>>(b, a, rows, tiles, 8, counts);
    // pass 2: a -> b
    radix_hist_kernel<false><<<grid, kSortThreads, 0, stream>>>(a, rows, tiles, 16, counts);
    radix_scan_kernel<<<segments, 1024, 0, stream>>>(counts, tiles);
    radix_scatter_kernel<false, false><<<grid, kSortThreads, 0, stream>>>(a, b, rows, tiles, 16, counts);
    // pass 3: b -> a (ordered uint -> float on store)
    radix_hist_kernel<false><<<grid, kSortThreads, 0, stream>>>(b, rows, tiles, 24, counts);
    radix_scan_kernel<<<segments, 1024, 0, stream>>>(counts, tiles);
    radix_scatter_kernel<false, true><<<grid, kSortThreads, 0, stream>>>(b, a, rows, tiles, 24, counts);
    BGAN_LAUNCH_OK();
    return 0;
}
