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

template <typename V>
__global__ void __launch_bounds__(256) densify_nchw_kernel(const int32_t* __restrict__ coords, const V* __restrict__ values,
                                                           int64_t nnz, int channels, int height, int width,
                                                           int rows_per_band, float divisor, float* __restrict__ out) {
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
      dst[c * plane] = v;
    }
  }
}

}  // namespace tcvn

extern "C" int tcvn_densify(const int32_t* coords, const void* values, tcvn_value_dtype value_dtype, int64_t nnz,
                            int channels, int n_images, int height, int width, float divisor, float* out,
                            tcvn_dense_layout layout, tcvn_stream_t stream) {
  using namespace tcvn;
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
                                                         height, width, rows, divisor, out);
  else if (value_dtype == TCVN_VAL_U8)
    densify_nchw_kernel<uint8_t><<<grid, 256, 0, stream>>>(coords, static_cast<const uint8_t*>(values), nnz, channels,
                                                           height, width, rows, divisor, out);
  else
    return fail(TCVN_ERR_ARG, "densify: unknown value dtype %d", (int)value_dtype);
  TCVN_LAUNCH_CHECK();
  return TCVN_OK;
}
