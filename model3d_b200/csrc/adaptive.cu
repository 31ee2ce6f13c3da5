// Adaptive per-pixel sampling rounds (see adaptive.h); restates the early-stop rule of
// rayRenderer.estimateColor / Converged (render3d/ray_renderer.go:112-173) on the device.
#include <algorithm>

#include "adaptive.h"

namespace m3d {

namespace {

__global__ void __launch_bounds__(256)
adaptive_flush_kernel(PathBatch b, const float4 *__restrict__ accum, AdaptiveState st, AdaptiveParams ap) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= b.nP) return;
  const int idx = batch_pixel(b, p) - st.pix_begin;
  if (st.divisor[idx] != 0) return;
  double s0 = st.csum[3 * idx], s1 = st.csum[3 * idx + 1], s2 = st.csum[3 * idx + 2];
  double q0 = st.csq[3 * idx], q1 = st.csq[3 * idx + 1], q2 = st.csq[3 * idx + 2];
  int divisor = 0;
  for (int s = 0; s < b.S; s++) {
    const int num = (int)b.sample0 + s;  // the reference's loop index of this sample
    const float4 a = __ldcs(accum + (size_t)s * b.nP + p);
    s0 += a.x;
    s1 += a.y;
    s2 += a.z;
    q0 += (double)a.x * a.x;
    q1 += (double)a.y * a.y;
    q2 += (double)a.z * a.z;
    if (num < ap.min_samples || num < 2) continue;
    // ray_renderer.go:134-146: statistics over `num` (== count - 1), population rescale
    const double n = (double)num, inv = 1.0 / n;
    const double resc = sqrt(n) / (n - 1.0);
    const double m[3] = {s0 * inv, s1 * inv, s2 * inv};
    const double v[3] = {fmax(q0 * inv - m[0] * m[0], 0.0), fmax(q1 * inv - m[1] * m[1], 0.0),
                         fmax(q2 * inv - m[2] * m[2], 0.0)};
    bool conv = true;
    for (int k = 0; k < 3; k++) {  // Converged (ray_renderer.go:157-173)
      const double sd = sqrt(v[k]) * resc;
      if (sd < ap.max_stddev) continue;
      if (ap.oversaturated_stddevs != 0.0 && m[k] - ap.oversaturated_stddevs * sd > 1.0) continue;
      conv = false;
    }
    if (conv) {
      divisor = num;  // `break` leaves numSamples at the loop index
      break;
    }
  }
  if (divisor == 0 && (int)b.sample0 + b.S >= ap.num_samples) divisor = ap.num_samples;
  st.csum[3 * idx] = s0;
  st.csum[3 * idx + 1] = s1;
  st.csum[3 * idx + 2] = s2;
  st.csq[3 * idx] = q0;
  st.csq[3 * idx + 1] = q1;
  st.csq[3 * idx + 2] = q2;
  st.divisor[idx] = divisor;
}

__global__ void __launch_bounds__(256)
adaptive_compact_kernel(const int32_t *__restrict__ pixels_in, int32_t pix0, int32_t n, AdaptiveState st,
                        int32_t *__restrict__ pixels_out, int *count_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned lane = threadIdx.x & 31u;
  bool active = false;
  int pix = 0;
  if (i < n) {
    pix = pixels_in ? pixels_in[i] : pix0 + i;
    active = st.divisor[pix - st.pix_begin] == 0;
  }
  const unsigned m = __ballot_sync(0xffffffffu, active);
  if (!m) return;
  int base = 0;
  if (lane == (unsigned)(__ffs(m) - 1)) base = atomicAdd(count_out, __popc(m));
  base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
  if (active) pixels_out[base + __popc(m & ((1u << lane) - 1u))] = pix;
}

__global__ void __launch_bounds__(256)
adaptive_finalize_kernel(int32_t npix, AdaptiveState st, AdaptiveParams ap, float *__restrict__ rgb_sum,
                         unsigned long long *samples_total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long taken = 0;
  if (i < npix) {
    const int d = st.divisor[i];
    // samples actually taken by the reference: the loop index on break is count - 1
    taken = (unsigned long long)(d == ap.num_samples ? d : d + 1);
    const double scale = (double)ap.num_samples / (double)d;  // colorSum.Scale(1/numSamples) * NumSamples
    const size_t o = (size_t)(st.pix_begin + i) * 3;
    rgb_sum[o] += (float)(st.csum[3 * i] * scale);
    rgb_sum[o + 1] += (float)(st.csum[3 * i + 1] * scale);
    rgb_sum[o + 2] += (float)(st.csum[3 * i + 2] * scale);
  }
  for (int off = 16; off > 0; off >>= 1) taken += __shfl_down_sync(0xffffffffu, taken, off);
  if ((threadIdx.x & 31u) == 0 && taken) atomicAdd(samples_total, taken);
}

}  // namespace

void launch_adaptive_flush(const PathBatch &b, const float4 *accum, const AdaptiveState &st,
                           const AdaptiveParams &ap, cudaStream_t stream) {
  if (b.nP <= 0) return;
  adaptive_flush_kernel<<<(unsigned)((b.nP + 255) / 256), 256, 0, stream>>>(b, accum, st, ap);
}

void launch_adaptive_compact(const int32_t *pixels_in, int32_t pix0, int32_t n, const AdaptiveState &st,
                             int32_t *pixels_out, int *count_out, cudaStream_t stream) {
  if (n <= 0) return;
  adaptive_compact_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(pixels_in, pix0, n, st, pixels_out,
                                                                           count_out);
}

void launch_adaptive_finalize(int32_t npix, const AdaptiveState &st, const AdaptiveParams &ap, float *rgb_sum,
                              unsigned long long *samples_total, cudaStream_t stream) {
  if (npix <= 0) return;
  adaptive_finalize_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, stream>>>(npix, st, ap, rgb_sum,
                                                                               samples_total);
}

int32_t run_adaptive(m3d_ctx *ctx, cudaStream_t s, int32_t width, int32_t pix_begin, int32_t npix, int64_t cap,
                     const AdaptiveParams &ap, const float4 *accum, float *d_rgb_sum,
                     const std::function<int32_t(const PathBatch &)> &run_batch, int64_t *samples_out) {
  const size_t n = (size_t)npix;
  const size_t o_csum = 0, o_csq = o_csum + n * 24, o_div = o_csq + n * 24, o_list0 = o_div + n * 4,
               o_list1 = o_list0 + n * 4, o_cnt = o_list1 + n * 4, total = o_cnt + 64;
  M3D_CUDA(ctx->scratch[8].reserve(total));
  char *p = ctx->scratch[8].as<char>();
  AdaptiveState st;
  st.pix_begin = pix_begin;
  st.csum = (double *)(p + o_csum);
  st.csq = (double *)(p + o_csq);
  st.divisor = (int32_t *)(p + o_div);
  int32_t *lists[2] = {(int32_t *)(p + o_list0), (int32_t *)(p + o_list1)};
  int *d_count = (int *)(p + o_cnt);
  unsigned long long *d_samples = (unsigned long long *)(p + o_cnt + 16);
  M3D_CUDA(cudaMemsetAsync(p, 0, o_list0, s));
  M3D_CUDA(cudaMemsetAsync(d_count, 0, 64, s));

  int32_t n_active = npix;
  const int32_t *cur_list = nullptr;  // nullptr: the full range
  int which = 0;
  int64_t done_samples = 0;  // samples every active pixel has consumed so far
  while (n_active > 0 && done_samples < ap.num_samples) {
    // round size: the first round reaches the first possible test, later rounds add 50 %
    int64_t want = done_samples == 0 ? std::max<int64_t>((int64_t)ap.min_samples + 1, 3)
                                     : std::max<int64_t>(16, done_samples / 2);
    want = std::min<int64_t>(want, ap.num_samples - done_samples);
    const int64_t chunk_pixels = std::min<int64_t>(n_active, cap);
    const int64_t S = std::max<int64_t>(1, std::min<int64_t>(want, cap / chunk_pixels));
    for (int64_t c0 = 0; c0 < n_active; c0 += chunk_pixels) {
      PathBatch b;
      b.W = width;
      b.nP = (int32_t)std::min<int64_t>(chunk_pixels, n_active - c0);
      b.S = (int32_t)S;
      b.sample0 = (uint32_t)done_samples;
      b.pix0 = pix_begin + (int32_t)c0;
      b.pixels = cur_list ? cur_list + c0 : nullptr;
      if (int32_t rc = run_batch(b)) return rc;
      launch_adaptive_flush(b, accum, st, ap, s);
    }
    done_samples += S;
    // compact the pixels that are still active
    M3D_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int), s));
    launch_adaptive_compact(cur_list, pix_begin, n_active, st, lists[which], d_count, s);
    int h_count = 0;
    M3D_CUDA(cudaMemcpyAsync(&h_count, d_count, sizeof(int), cudaMemcpyDeviceToHost, s));
    M3D_CUDA(cudaStreamSynchronize(s));
    n_active = h_count;
    cur_list = lists[which];
    which ^= 1;
  }
  launch_adaptive_finalize(npix, st, ap, d_rgb_sum, d_samples, s);
  unsigned long long h_samples = 0;
  M3D_CUDA(cudaMemcpyAsync(&h_samples, d_samples, sizeof(h_samples), cudaMemcpyDeviceToHost, s));
  M3D_CUDA(cudaStreamSynchronize(s));
  M3D_CUDA(cudaGetLastError());
  if (samples_out) *samples_out = (int64_t)h_samples;
  return M3D_OK;
}

}  // namespace m3d
