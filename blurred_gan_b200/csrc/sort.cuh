// Shared between the two segmented sorts of fp32 keys (sort.cu: LSD radix; sort_sample.cu: value buckets sorted in shared memory).
#pragma once

#include "common.cuh"

namespace bgan {

// The |a - b| partial sums of a fused "sort B and meet sorted A" go to kDiffSlots accumulators per sum, added up by a
// one-warp-per-sum kernel afterwards: thousands of fp64 atomics on a handful of addresses would serialise in L2.
constexpr int kDiffSlots = 64;
struct DiffArgs {
    const float* ref;  // sorted reference segments (same layout as the keys), or nullptr
    double* partial;   // [sums][kDiffSlots]
    int seg_offset;    // sum index of segment s = (s + seg_offset) / segs_per_sum
    int segs_per_sum;
};

#ifdef __CUDACC__
// order-preserving image of a float's bits (and back): negative floats reversed below the positive ones
__device__ __forceinline__ uint32_t float_to_ordered(uint32_t b) { return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u); }
__device__ __forceinline__ uint32_t ordered_to_float(uint32_t u) { return u ^ ((u >> 31) ? 0x80000000u : 0xffffffffu); }
#endif

// sort_sample.cu
bool sample_sort_applicable(int64_t rows);
size_t sample_sort_workspace_bytes(int64_t rows, int segments);
// Enqueues the sort of `segments` runs of `rows` keys (in place, or -- with da.ref -- sum |ref - sorted keys| into da.partial
// with the keys left untouched).  *flag_device: a device word that is non-zero after the kernels have run if a bucket did not fit
// shared memory; in that case NOTHING was written (neither keys nor partial sums) and the caller sorts with the radix path.
// nullptr: this shape cannot overflow (a segment fits one CTA).
int sample_sort_launch(float* keys_device, int64_t rows, int segments, void* ws, cudaStream_t stream, const DiffArgs& da,
                       const uint32_t** flag_device);

}  // namespace bgan
