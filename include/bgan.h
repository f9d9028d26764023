/*
 * bgan.h -- C ABI of libbgan_b200.so: the sm_100a kernels behind blurred-GAN's hot path
 * (sigma-annealed Gaussian blur fwd/bwd + sliced-Wasserstein evaluation).
 *
 * The reference (lebrice/blurred-GAN) has no FFI of its own: its boundary is the Python
 * surface of three files.  Each entry point below names the reference lines whose
 * arithmetic it replaces (paths relative to the reference repository root).
 *
 * Conventions
 *   - every pointer called "device" is a raw CUDA device pointer owned by the caller;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous w.r.t. the host;
 *   - no call allocates persistent device memory: scratch is passed in (`ws`, `ws_bytes`),
 *     sized by the matching *_workspace_bytes query;
 *   - return value: 0 = OK, > 0 = cudaError_t, < 0 = argument error (BGAN_E_*);
 *     bgan_last_error() returns a thread-local message for the last non-zero return;
 *   - functions are re-entrant across streams and devices (the current device must be
 *     the device that owns the pointers).
 */
#ifndef BGAN_H_
#define BGAN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BGAN_ABI_VERSION 4 /* 2: + bgan_blur2d_multi_f32, bgan_lerp_rows_f32, bgan_swd_feed_f32; 3: + bgan_swd_sort_diff_f32; 4: + bgan_swd_sort_last_path */

/* tap rules */
#define BGAN_RULE_TORCH 0    /* pytorch/gaussianblur.py:9-14 (normaliser inside the exp)  */
#define BGAN_RULE_TF 1       /* gaussian_blur.py:83-88 (true Gaussian)                   */
#define BGAN_RULE_EXPLICIT 2 /* caller supplies 2*radius+1 taps on the device             */

/* memory layouts */
#define BGAN_LAYOUT_NCHW 0
#define BGAN_LAYOUT_NHWC 1

/* argument errors */
#define BGAN_E_BADARG (-1)
#define BGAN_E_WORKSPACE (-2)
#define BGAN_E_UNSUPPORTED (-3)

int bgan_version(void);
const char* bgan_last_error(void);

/* ------------------------------------------------------------------------------------ */
/* Gaussian blur                                                                          */
/* ------------------------------------------------------------------------------------ */

/* Host-only.  Kernel-size rules:
 *   BGAN_RULE_TORCH: pytorch/gaussianblur.py:17-30,43-57,61  k = max(3, odd_up(ceil(5 std))),
 *                    clipped to odd_up(min(H,W)) - 2;            std_eff = std
 *   BGAN_RULE_TF:    gaussian_blur.py:21-31,58-59,65-72          k = floor(6 std) + 1, clipped
 *                    to [3, max(H,W)];  std_eff = max((k-1)/6, 0.01); taps span [-(k/2), k/2]
 * Outputs: *kernel_size = k, *radius = k / 2, *std_eff = the std that enters the exponential. */
int bgan_blur_plan(int rule, float std, int height, int width, int* kernel_size, int* radius, float* std_eff);

/* Writes the 2*radius+1 normalised taps exactly as the blur kernels compute them
 * (pytorch/gaussianblur.py:9-14 / gaussian_blur.py:83-88) to device memory. */
int bgan_blur_taps_f32(int rule, float std_eff, int radius, float* taps_out_device, void* stream);

/* Scratch needed by bgan_blur2d_f32 for this problem (0 when the fused single-kernel path applies). */
size_t bgan_blur2d_workspace_bytes(int n, int c, int height, int width, int layout, int radius);

/* y = B x: separable depthwise blur, zero "same" padding, taps generated in-kernel from
 * (rule, std_eff, radius) or read from taps_device (rule == BGAN_RULE_EXPLICIT).
 * Replaces GaussianBlur.forward (pytorch/gaussianblur.py:59-71), gaussian_blur()
 * (gaussian_blur.py:91-132) and -- B being self-adjoint -- their autograd backward and
 * double-backward (call sites wgan.py:140,166,244): backward is the same call on grad_output.
 * x and y are device pointers to N*C*H*W floats in `layout`; x == y is NOT allowed.
 * Non-finite inputs: an Inf / NaN pixel contaminates its tap neighbourhood on the FFMA paths (1, 2, 3, 5), like the reference's
 * convolution, and its whole plane on the tensor-core paths (4, 6, 7: 0 x NaN inside the MMAs); other planes are unaffected. */
int bgan_blur2d_f32(const float* x_device, float* y_device, int n, int c, int height, int width, int layout,
                    int rule, float std_eff, int radius, const float* taps_device,
                    void* ws_device, size_t ws_bytes, void* stream);

/* The same blur applied to `count` (1 .. 4) equally shaped tensors: ys[t] = B xs[t].  xs / ys are HOST arrays of device
 * pointers.  128 x 128 planes are served by ONE launch over all tensors (fuller persistent grid, one pipeline fill and
 * drain instead of `count`); other shapes run one launch sequence per tensor on `stream`, sharing the scratch (size: the
 * single-tensor query).  Call sites this serves: the critic step blurs the generated and the real batch with the same
 * sigma back to back (wgan.py:136-137 through blurred_gan.py:30-35), SURVEY section 8f row N3.  No output may alias any
 * other tensor of the call. */
int bgan_blur2d_multi_f32(const float* const* xs_device, float* const* ys_device, int count, int n, int c, int height,
                          int width, int layout, int rule, float std_eff, int radius, const float* taps_device,
                          void* ws_device, size_t ws_bytes, void* stream);

/* out[i][:] = u[i][:] + weights[i] * (v[i][:] - u[i][:]) for `rows` rows of `per_row` floats: the gradient-penalty
 * interpolates x_hat = reals + a (fakes - reals) of wgan.py:237-239.  Applied to two *blurred* batches it yields
 * B x_hat without a third forward blur (B is linear). */
int bgan_lerp_rows_f32(const float* u_device, const float* v_device, const float* weights_device, float* out_device,
                       int64_t rows, int64_t per_row, void* stream);

/* How the last bgan_blur2d_f32 / bgan_blur2d_multi_f32 call on this thread was executed: 1 = fused tile kernel with TMA
 * staging, 2 = fused tile kernel with plain loads, 3 = two launches (row pass, column pass), 4 = tcgen05 tensor-core kernel (128 x 128 planes),
 * 5 = direct FFMA kernel for 128 x 128 images with radius <= 5 NCHW / <= 4 NHWC (NCHW radius 6 .. 8 only with caller-supplied taps), 6 = two-pass tcgen05 kernel for planes of
 * 256 x 256 and more with radius 12 .. 256 (needs the workspace), 7 = the same for NHWC images with 2 .. 4 channels:
 * interleaved -> planar copy, path 6 on the channel planes, planar -> interleaved (needs the workspace). */
int bgan_blur2d_last_path(void);

/* ------------------------------------------------------------------------------------ */
/* Sliced Wasserstein distance (sliced_wasserstein.py)                                    */
/* ------------------------------------------------------------------------------------ */

/* sliced_wasserstein.py:72-74  pyr_down: 5x5 binomial/256, mirror border, keep even rows/cols.
 * x: [planes, H, W] -> y: [planes, (H+1)/2, (W+1)/2].  fp64 accumulate, fp32 store. */
int bgan_pyr_down_f32(const float* x_device, float* y_device, int planes, int height, int width, void* stream);

/* sliced_wasserstein.py:76-81  pyr_up: zero-stuff to 2x, 5x5 binomial * 4/256, mirror border.
 * x: [planes, h, w] -> y: [planes, 2h, 2w].  If minuend_device != NULL writes
 * y = minuend - pyr_up(x) instead (the Laplacian residual of :87). */
int bgan_pyr_up_f32(const float* x_device, const float* minuend_device, float* y_device, int planes, int h, int w,
                    void* stream);

/* sliced_wasserstein.py:13-23  gather nhoods_per_image patches of nhood x nhood x C per image at
 * the caller-drawn centres (px = column, py = row; int32 device arrays of n*nhoods_per_image)
 * into desc_out[m, c, i, j] = level[m / per_image, c, py[m]+j-half, px[m]+i-half], and add the
 * per-channel sum and sum of squares of everything written to sums_device[0..C) / [C..2C)
 * (fp64, caller zeroes them) -- the statistics half of finalize_descriptors (:31-32). */
int bgan_swd_extract_f32(const float* level_device, int n, int c, int h, int w, const int32_t* px_device,
                         const int32_t* py_device, int nhoods_per_image, int nhood, float* desc_out_device,
                         double* sums_device, void* stream);

/* sliced_wasserstein.py:119-123  one API.feed as ONE host call: Laplacian pyramid of the minibatch [n, c, h, w] (h, w
 * divisible by 2^(levels-1)) and the descriptor gather of every level -- the kernels of bgan_pyr_down_f32,
 * bgan_pyr_up_f32 (residual form) and bgan_swd_extract_f32 launched back to back on `stream`; the minibatch is not
 * modified.  positions_device: int32 [levels][x | y][n * nhoods_per_image] (the layout bgan_swd_positions_i32 writes);
 * desc_out / sums: HOST arrays of `levels` device pointers (per level: [n * nhoods_per_image, c, nhood, nhood] floats;
 * 2c doubles accumulated into, as in bgan_swd_extract_f32).  Scratch: bgan_swd_feed_workspace_bytes. */
size_t bgan_swd_feed_workspace_bytes(int n, int c, int h, int w, int levels);
int bgan_swd_feed_f32(const float* minibatch_device, int n, int c, int h, int w, int levels, const int32_t* positions_device,
                      int nhoods_per_image, int nhood, float* const* desc_out_device, double* const* sums_device,
                      void* ws_device, size_t ws_bytes, void* stream);

/* sliced_wasserstein.py:31-32  desc = (desc - mean_c) / std_c in place, with mean/std derived on
 * the device from sums_device (as filled by bgan_swd_extract_f32, possibly all-reduced across
 * ranks) and `count` = number of values per channel.  rows x (C*per_channel) floats. */
int bgan_swd_normalise_f32(float* desc_device, int64_t rows, int c, int per_channel, const double* sums_device,
                           double count, void* stream);

/* sliced_wasserstein.py:45-46  proj = desc @ dirs, written direction-major:
 * desc [rows, k] row-major, dirs [k, d] row-major -> proj_out [d, rows]. */
int bgan_swd_project_f32(const float* desc_device, int64_t rows, int k, const float* dirs_device, int d,
                         float* proj_out_device, void* stream);
/* How the last bgan_swd_project_f32 call on this thread was executed: 1 = FP32 FFMA kernel, 2 = tcgen05 kernel
 * (error-compensated fp16 split, k <= 192). */
int bgan_swd_project_last_path(void);

/* sliced_wasserstein.py:47-48  ascending sort of each of the `segments` contiguous runs of `rows` floats, in place (ws
 * from the query below).  The order is that of the floats' order-preserving uint32 images: -0 before +0, NaNs with the sign
 * bit clear last (as np.sort), NaNs with the sign bit set first.
 * Segments of up to 2^21 keys are sorted by value buckets in shared memory (two passes over HBM: csrc/sort_sample.cu);
 * longer segments, and inputs whose values are so unevenly spread that a bucket does not fit a CTA (outliers hundreds of
 * standard deviations out; up to 16 heavily repeated values per segment are recognised and cost nothing), by four LSD radix
 * passes (csrc/sort.cu).  The second
 * case is detected on the device: for segments of more than 8192 keys the call therefore reads one 4-byte flag back and
 * SYNCHRONISES THE STREAM before it returns (a stream that is being captured takes the radix passes, which never do). */
size_t bgan_swd_sort_workspace_bytes(int64_t rows, int segments);
int bgan_swd_sort_f32(float* keys_device, int64_t rows, int segments, void* ws_device, size_t ws_bytes, void* stream);
/* How the last sort (bgan_swd_sort_f32, or the second sort of bgan_swd_sort_diff_f32) on this thread was executed:
 * 1 = LSD radix passes, 2 = value buckets, 3 = value buckets gave up (flag) and the radix passes did the work. */
int bgan_swd_sort_last_path(void);

/* sliced_wasserstein.py:49-50  *sum_out_device += sum_i |a[i] - b[i]| (fp64; caller zeroes it). */
int bgan_swd_absdiff_sum_f32(const float* a_device, const float* b_device, int64_t count, double* sum_out_device,
                             void* stream);

/* sliced_wasserstein.py:47-50 as ONE call: sort every segment of a_device in place (ascending), sort the segments of
 * b_device, and add sum_i |sort(a)[i] - sort(b)[i]| of segment s to sums_device[(s + seg_offset) / segs_per_sum] (fp64,
 * accumulated into; n_sums entries).  The last step of the sort of b does not write the keys: every key meets the element of
 * the sorted a that sits at its final position and only the distance is accumulated -- the sorted b is never stored and the
 * separate |a - b| pass disappears.  b_device is clobbered (unspecified contents on return).  a and b: `segments` contiguous
 * runs of `rows` floats.  seg_offset / segs_per_sum: a rank that owns directions [seg_offset, seg_offset + segments) of the
 * dir_repeats x dirs_per_repeat matrix passes segs_per_sum = dirs_per_repeat and gets one sum per repeat. */
size_t bgan_swd_sort_diff_workspace_bytes(int64_t rows, int segments, int n_sums);
int bgan_swd_sort_diff_f32(float* a_device, float* b_device, int64_t rows, int segments, int seg_offset, int segs_per_sum,
                           int n_sums, double* sums_device, void* ws_device, size_t ws_bytes, void* stream);

/* Host-only.  Bit-exact replay of numpy's legacy np.random.randint(low, high, size=count) -- the call behind the patch
 * positions of sliced_wasserstein.py:20-21 -- on a caller-held MT19937 state: key624 / *pos_inout are the `key` and `pos`
 * fields of np.random.get_state(), updated in place so that np.random.set_state() continues the same stream.
 * out_host (count int32, host memory) may be NULL: the stream is advanced and nothing is stored (the draws of other
 * ranks' minibatches in a multi-rank evaluation).  Ranges are limited to what numpy serves with 32-bit masked rejection
 * (high - low <= 2^32 - 1) and to int32 results. */
int bgan_mt19937_randint_i32(uint32_t* key624, int* pos_inout, int64_t low, int64_t high, int64_t count, int32_t* out_host);

/* Host-only.  All patch-position draws of one `feed` round (sliced_wasserstein.py:20-21 as called from :121-123), in the
 * order a single process feeding the `world` minibatches of this round one after the other would make them: for every
 * minibatch, for every pyramid level (finest first), `count` columns in [nhood/2, width - nhood/2) then `count` rows in
 * [nhood/2, height - nhood/2).  The draws of minibatch `rank` are stored as out_host[level][x | y][count] (int32, host
 * memory, e.g. a pinned staging buffer); the others only advance the stream.  State as in bgan_mt19937_randint_i32. */
int bgan_swd_positions_i32(uint32_t* key624, int* pos_inout, int levels, const int* heights, const int* widths, int nhood,
                           int64_t count, int world, int rank, int32_t* out_host);

#ifdef __cplusplus
}
#endif
#endif /* BGAN_H_ */
