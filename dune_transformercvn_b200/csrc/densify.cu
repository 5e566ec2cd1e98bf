// K1 ingest: Minkowski-format COO hit lists -> dense NCHW fp32 pixel maps, one pass over the output.
//
// Replaces transformercvn/network/trainers/neutrino_full_dense_trainer.py:15-24 (sparse_to_dense:
// zeros(N,H,W,C); out[img,y,x] += v; permute(0,3,1,2).contiguous()) and the "/ 255.0" of
// preprocess_pixels (:59-60).  The reference makes three passes over 1.344 MB/image (fill,
// index_put_, permute copy) plus a .item() sync; here each CTA owns a band of rows of one image,
// stores its zeros with 16-byte writes, barriers, and drops that band's hits on top while the lines
// are still in L2, so HBM sees each output byte once.  HBM-bound: algorithmic bytes =
// nnz*(12 + C*sizeof(value)) read + N*C*H*W*4 written.
#include "common.cuh"

namespace tcvn {

__device__ __forceinline__ int64_t lower_bound_image(const int32_t* __restrict__ coords, int64_t nnz, int image) {
  int64_t lo = 0, hi = nnz;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (__ldg(coords + 3 * mid) < image) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// standard normal from a counter hash of (seed, element): two 24-bit uniforms -> Box-Muller
__device__ __forceinline__ float hash_normal(unsigned long long seed, unsigned long long idx) {
  unsigned long long z = seed * 0x100000001b3ull + idx + 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  z ^= z >> 31;
  const float u1 = ((float)(unsigned)(z >> 40) + 1.0f) * (1.0f / 16777216.0f);        // (0, 1]
  const float u2 = (float)(unsigned)((z >> 16) & 0xffffffu) * (1.0f / 16777216.0f);   // [0, 1)
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

template <typename V>
__global__ void __launch_bounds__(256) densify_nchw_kernel(const int32_t* __restrict__ coords, const V* __restrict__ values,
                                                           int64_t nnz, int channels, int height, int width,
                                                           int rows_per_band, float divisor, float noise_std,
                                                           unsigned long long seed0, const unsigned long long* seed_off,
                                                           float* __restrict__ out) {
  const unsigned long long seed = seed_with_offset(seed0, seed_off);
  const int image = blockIdx.y;
  const int y0 = blockIdx.x * rows_per_band;
  const int y1 = min(height, y0 + rows_per_band);
  __shared__ int64_t range[2];
  if (threadIdx.x == 0) range[0] = lower_bound_image(coords, nnz, image);
  if (threadIdx.x == 32) range[1] = lower_bound_image(coords, nnz, image + 1);
  const size_t plane = (size_t)height * width;
  float* img = out + (size_t)image * channels * plane;
  // zeros: every channel plane's band is one contiguous run of floats
  const size_t band = (size_t)(y1 - y0) * width;
  for (int c = 0; c < channels; ++c) {
    float* dst = img + c * plane + (size_t)y0 * width;
    // 16-byte stores where the run is aligned (width % 4 == 0 for the 400x280 maps), scalar edges otherwise
    size_t head = ((16 - ((uintptr_t)dst & 15)) & 15) >> 2;
    if (head > band) head = band;
    for (size_t i = threadIdx.x; i < head; i += blockDim.x) dst[i] = 0.f;
    const size_t n4 = (band - head) >> 2;
    float4* d4 = reinterpret_cast<float4*>(dst + head);
    for (size_t i = threadIdx.x; i < n4; i += blockDim.x) d4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (size_t i = head + (n4 << 2) + threadIdx.x; i < band; i += blockDim.x) dst[i] = 0.f;
  }
  __syncthreads();  // orders the zero stores of this CTA before its hit stores (same CTA owns the band)
  const int64_t lo = range[0], hi = range[1];
  for (int64_t h = lo + threadIdx.x; h < hi; h += blockDim.x) {
    const int y = __ldg(coords + 3 * h + 1);
    const int x = __ldg(coords + 3 * h + 2);
    if (y < y0 || y >= y1 || x < 0 || x >= width) continue;
    float* dst = img + (size_t)y * width + x;
    for (int c = 0; c < channels; ++c) {
      float v = static_cast<float>(values[h * channels + c]);
      if (divisor != 0.f) v = __fdiv_rn(v, divisor);  // IEEE division, same bits as torch's v / 255.0
      // training-time pixel noise of preprocess_pixels (:62-65): v *= 1 + randn * std, one draw per stored value
      if (noise_std != 0.f) v *= 1.0f + hash_normal(seed, (unsigned long long)h * channels + c) * noise_std;
      dst[c * plane] = v;
    }
  }
}

}  // namespace tcvn

extern "C" int tcvn_densify(const int32_t* coords, const void* values, tcvn_value_dtype value_dtype, int64_t nnz,
                            int channels, int n_images, int height, int width, float divisor, float* out,
                            tcvn_dense_layout layout, tcvn_stream_t stream) {
  return tcvn_densify_noise(coords, values, value_dtype, nnz, channels, n_images, height, width, divisor, 0.f, 0, out, layout,
                            stream);
}

extern "C" int tcvn_densify_noise(const int32_t* coords, const void* values, tcvn_value_dtype value_dtype, int64_t nnz,
                                  int channels, int n_images, int height, int width, float divisor, float noise_std,
                                  uint64_t seed, float* out, tcvn_dense_layout layout, tcvn_stream_t stream) {
  using namespace tcvn;
  TCVN_CHECK_ARG(noise_std >= 0.f, "densify: negative noise_std");
  TCVN_CHECK_ARG(layout == TCVN_NCHW_F32, "densify: unknown layout %d", (int)layout);
  TCVN_CHECK_ARG(n_images >= 0 && channels > 0 && height > 0 && width > 0 && nnz >= 0, "densify: bad sizes");
  TCVN_CHECK_ARG(n_images <= 65535, "densify: more than 65535 images in one call");
  if (n_images == 0) return TCVN_OK;
  TCVN_CHECK_ARG(out != nullptr && (nnz == 0 || (coords != nullptr && values != nullptr)), "densify: null pointer");
  // enough CTAs for >= 4 waves of 148 SMs x 8 resident CTAs when the batch is small; bands of >= 8 rows
  int bands = ceil_div(148 * 8 * 4, n_images);
  if (bands < 1) bands = 1;
  int rows = ceil_div(height, bands);
  if (rows < 8) rows = 8;
  bands = ceil_div(height, rows);
  dim3 grid(bands, n_images);
  if (value_dtype == TCVN_VAL_F32)
    densify_nchw_kernel<float><<<grid, 256, 0, stream>>>(coords, static_cast<const float*>(values), nnz, channels,
                                                         height, width, rows, divisor, noise_std, seed, seed_offset_ptr(), out);
  else if (value_dtype == TCVN_VAL_U8)
    densify_nchw_kernel<uint8_t><<<grid, 256, 0, stream>>>(coords, static_cast<const uint8_t*>(values), nnz, channels,
                                                           height, width, rows, divisor, noise_std, seed, seed_offset_ptr(), out);
  else
    return fail(TCVN_ERR_ARG, "densify: unknown value dtype %d", (int)value_dtype);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}

// ------------------------------------------------------------------------------------------------
// Collate: per-event COO hit lists (image index local to the event) -> one batch-global hit list.
// Replaces MinkowskiCollection.collate_sparse (transformercvn/dataset/minkowski_dataset.py:34-47): a Python loop that
// adds the running image count to column 0 of every event's coordinates, with one `.item()` host sync per event.
// Here: an exclusive scan of the per-event image counts (one CTA), then one thread per hit finds its event by binary
// search in the hit-count prefix and adds the offset.  Integer work, HBM-bound: 12 B read + 12 B written per hit.
// ------------------------------------------------------------------------------------------------
namespace tcvn {

// prefix[0] = 0, prefix[e+1] = prefix[e] + counts[e]   (two arrays at once; B is a few thousand at most)
__global__ void collate_scan_kernel(const int32_t* __restrict__ hits_per_event, const int32_t* __restrict__ images_per_event,
                                    const uint8_t* __restrict__ masks, int n_events, int max_slots,
                                    int64_t* __restrict__ hit_prefix, int32_t* __restrict__ image_prefix) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int64_t h = 0;
  int32_t im = 0;
  for (int e = 0; e < n_events; ++e) {
    hit_prefix[e] = h;
    image_prefix[e] = im;
    h += hits_per_event[e];
    if (images_per_event) im += images_per_event[e];
    else for (int l = 0; l < max_slots; ++l) im += masks[(size_t)e * max_slots + l] != 0;
  }
  hit_prefix[n_events] = h;
  image_prefix[n_events] = im;
}

__global__ void collate_offset_kernel(const int32_t* __restrict__ coords_in, int64_t nnz, const int64_t* __restrict__ hit_prefix,
                                      const int32_t* __restrict__ image_prefix, int n_events, int32_t* __restrict__ coords_out) {
  const int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= nnz) return;
  int lo = 0, hi = n_events;   // last event e with hit_prefix[e] <= h
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (hit_prefix[mid] <= h) lo = mid; else hi = mid;
  }
  coords_out[3 * h] = coords_in[3 * h] + image_prefix[lo];
  coords_out[3 * h + 1] = coords_in[3 * h + 1];
  coords_out[3 * h + 2] = coords_in[3 * h + 2];
}

}  // namespace tcvn

extern "C" size_t tcvn_collate_workspace_bytes(int n_events) {
  return ((size_t)(n_events > 0 ? n_events : 0) + 1) * (sizeof(int64_t) + sizeof(int32_t)) + 64;
}

extern "C" int tcvn_collate_coords(const int32_t* coords_in, int64_t nnz, const int32_t* hits_per_event,
                                   const int32_t* images_per_event, const uint8_t* masks, int n_events, int max_slots,
                                   int32_t* coords_out, void* workspace, size_t workspace_bytes, tcvn_stream_t stream) {
  using namespace tcvn;
  TCVN_CHECK_ARG(n_events >= 0 && nnz >= 0 && max_slots >= 0, "collate: negative size");
  if (n_events == 0 || nnz == 0) return TCVN_OK;
  TCVN_CHECK_ARG(coords_in && coords_out && hits_per_event && workspace && (images_per_event || masks), "collate: null pointer");
  if (workspace_bytes < tcvn_collate_workspace_bytes(n_events))
    return fail(TCVN_ERR_WORKSPACE, "collate: workspace %zu < %zu bytes", workspace_bytes, tcvn_collate_workspace_bytes(n_events));
  int64_t* hit_prefix = static_cast<int64_t*>(workspace);
  int32_t* image_prefix = reinterpret_cast<int32_t*>(hit_prefix + n_events + 1);
  collate_scan_kernel<<<1, 32, 0, stream>>>(hits_per_event, images_per_event, masks, n_events, max_slots, hit_prefix, image_prefix);
  TCVN_LAUNCH_CHECK();
  collate_offset_kernel<<<(unsigned)ceil_div_ll(nnz, 256), 256, 0, stream>>>(coords_in, nnz, hit_prefix, image_prefix, n_events,
                                                                            coords_out);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}
