// Per-ray training samples: utils/data_sampler.py:18-260 of the reference builds, for
// every measured point, the end point itself, `ns` Gaussian samples around it, `nf` uniform samples in front and `nb`
// behind, with ~35 torch ops over repeated / concatenated [n * total] tensors and a final transpose into ray-major
// order.  The random draws stay torch calls (same generator, same order, same sizes as the reference: the RNG
// stream is part of the parity contract); everything else is this one kernel, thread per output sample, written
// directly in ray-major order.
#include <algorithm>

#include "common.cuh"

namespace pinb {

struct SampleParams {
  const float* points;  // [n,3] sensor frame
  const float* colors;  // [n,cc] or null
  const float* z_surf;  // [ns*n] standard normal draws   (index j*n + i: the reference repeats whole arrays)
  const float* u_front; // [nf*n] uniform draws
  const float* u_behind;// [nb*n]
  long long n;
  int ns, nf, nb, cc;
  float sigma, begin_ratio, end_dist, max_range, dist_weight_scale;
  int dist_weight_on, behind_dropoff_on;
  float* coord;   // [n*total,3]
  float* label;   // [n*total]
  float* weight;  // [n*total]
  float* color;   // [n*total,cc] or null
};

__global__ void __launch_bounds__(256) ray_sample_kernel(const SampleParams p) {
  const int total = 1 + p.ns + p.nf + p.nb;
  const long long m = p.n * total;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < m; e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / total;
    const int s = (int)(e - i * total);
    const float x = p.points[3 * i], y = p.points[3 * i + 1], z = p.points[3 * i + 2];
    const float d = sqrtf(x * x + y * y + z * z);
    float ratio = 1.f, disp = 0.f;
    if (s >= 1 && s <= p.ns) {  // around the surface
      disp = p.z_surf[(long long)(s - 1) * p.n + i] * p.sigma;
      ratio = disp / d + 1.f;
    } else if (s > p.ns && s <= p.ns + p.nf) {  // free space in front
      const float hi = 1.f - 2.f * p.sigma / d;
      ratio = p.u_front[(long long)(s - 1 - p.ns) * p.n + i] * (hi - p.begin_ratio) + p.begin_ratio;
      disp = (ratio - 1.f) * d;
    } else if (s > p.ns + p.nf) {  // behind the surface
      const float lo = 1.f + 2.f * p.sigma / d;
      ratio = p.u_behind[(long long)(s - 1 - p.ns - p.nf) * p.n + i] * (p.end_dist / d + 1.f - lo) + lo;
      disp = (ratio - 1.f) * d;
    }
    float w = 1.f;
    if (p.dist_weight_on && s <= p.ns) w = 1.f + p.dist_weight_scale * 0.5f - (d / p.max_range) * p.dist_weight_scale;
    if (p.behind_dropoff_on) {
      const float lo = 0.2f * p.end_dist, hi = p.end_dist;
      w *= fminf(fmaxf((hi - disp) / (hi - lo), 0.f), 1.f) * 0.8f + 0.2f;
    }
    if (s > p.ns) w = -w;  // the sign marks free-space samples (data_sampler.py:168)
    p.coord[3 * e] = x * ratio;
    p.coord[3 * e + 1] = y * ratio;
    p.coord[3 * e + 2] = z * ratio;
    p.label[e] = -disp;
    p.weight[e] = w;
    if (p.color)
      for (int c = 0; c < p.cc; ++c) p.color[e * p.cc + c] = s <= p.ns ? p.colors[i * p.cc + c] : 0.f;
  }
}

}  // namespace pinb

using namespace pinb;

extern "C" int pinb200_ray_samples(const float* points, const float* colors, int32_t color_channels, int64_t n,
                                   const float* z_surf, const float* u_front, const float* u_behind, int32_t n_surface,
                                   int32_t n_front, int32_t n_behind, float sigma, float free_begin_ratio, float free_end_dist,
                                   float max_range, int32_t dist_weight_on, float dist_weight_scale, int32_t behind_dropoff_on,
                                   float* coord, float* label, float* weight, float* color, void* stream) {
  if (!points || !coord || !label || !weight || n < 0 || n_surface < 0 || n_front < 0 || n_behind < 0 ||
      (n_surface && !z_surf) || (n_front && !u_front) || (n_behind && !u_behind) || (color && (!colors || color_channels < 1))) {
    set_error("ray_samples: bad argument");
    return PINB200_ERR_BAD_ARG;
  }
  if (n == 0) return PINB200_OK;
  SampleParams p{points, colors, z_surf, u_front, u_behind, n, n_surface, n_front, n_behind, color_channels, sigma,
                 free_begin_ratio, free_end_dist, max_range, dist_weight_scale, dist_weight_on, behind_dropoff_on, coord, label,
                 weight, color};
  const long long m = n * (1 + n_surface + n_front + n_behind);
  const int grid = (int)std::min<long long>((m + 255) / 256, (long long)sm_count() * 16);
  ray_sample_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  return check_launch("ray_sample_kernel");
}
