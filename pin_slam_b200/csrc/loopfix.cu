// Loop-closure map surgery: after a pose-graph optimisation the reference moves every
// neural point -- and every replay-pool sample -- by the pose correction of the frame it belongs to:
//   model/neural_points.py:791-822 (adjust_map: points, orientations)   utils/mapper.py:527-531 (transform_data_pool)
// as gather + bmm + quaternion product over O(Mg) / O(pool) rows in ~10 torch ops.  One streaming kernel here:
// rows are independent, 12 + 16 bytes in / out each, the per-frame corrections (a few hundred 3x4 matrices and
// quaternions) stay in L1/L2.
#include <algorithm>

#include "common.cuh"

namespace pinb {

__global__ void __launch_bounds__(256) frame_transform_kernel(float* __restrict__ xyz, float* __restrict__ quat,
                                                              const int32_t* __restrict__ ts_a, const int32_t* __restrict__ ts_b,
                                                              const float* __restrict__ tf, const float* __restrict__ dq,
                                                              long long n, int n_ts) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    // frame of the row: creation stamp, or the (truncated) mean of creation and last update (config.use_mid_ts)
    int t = ts_a[i];
    if (ts_b) t = (int)(((float)t + (float)ts_b[i]) / 2.f);
    if (t < 0) t += n_ts;  // torch index wrap
    const float* T = tf + 12 * (size_t)t;
    const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    xyz[3 * i + 0] = fmaf(T[2], z, fmaf(T[1], y, T[0] * x)) + T[3];
    xyz[3 * i + 1] = fmaf(T[6], z, fmaf(T[5], y, T[4] * x)) + T[7];
    xyz[3 * i + 2] = fmaf(T[10], z, fmaf(T[9], y, T[8] * x)) + T[11];
    if (quat) {  // q <- dq (x) q, wxyz (utils/tools.py quat_multiply)
      const float4 a = __ldg(reinterpret_cast<const float4*>(dq) + t);
      const float4 b = reinterpret_cast<const float4*>(quat)[i];
      float4 o;
      o.x = a.x * b.x - a.y * b.y - a.z * b.z - a.w * b.w;
      o.y = a.x * b.y + a.y * b.x + a.z * b.w - a.w * b.z;
      o.z = a.x * b.z - a.y * b.w + a.z * b.x + a.w * b.y;
      o.w = a.x * b.w + a.y * b.z - a.z * b.y + a.w * b.x;
      reinterpret_cast<float4*>(quat)[i] = o;
    }
  }
}

}  // namespace pinb

using namespace pinb;

extern "C" int pinb200_frame_transform(float* xyz, float* quat, const int32_t* ts_a, const int32_t* ts_b, const float* tf3x4,
                                       const float* dquat, int64_t n, int32_t n_ts, void* stream) {
  if (!xyz || !ts_a || !tf3x4 || n < 0 || n_ts <= 0 || (quat && !dquat)) {
    set_error("frame_transform: null argument");
    return PINB200_ERR_BAD_ARG;
  }
  if (n == 0) return PINB200_OK;
  const int grid = (int)std::min<long long>((n + 255) / 256, (long long)sm_count() * 8);
  frame_transform_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(xyz, quat, ts_a, ts_b, tf3x4, dquat, n, n_ts);
  return check_launch("frame_transform_kernel");
}
