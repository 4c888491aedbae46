// Map growth: model/neural_points.py:311-422 (NeuralPoints.update) of the reference
// tests every voxel-down-sampled scan point against the hash table -- empty slot, owner farther than sqrt(3) voxels,
// or owner not refreshed within the local travel distance -- and appends the ones that pass to seven tensors with
// torch.cat, then rewrites the table with an index_put that has duplicate indices.  Here:
//   * the point arrays live in caller-owned fixed-capacity arenas; new points are written in place behind the current
//     count, in candidate order (block counts -> scan -> ordered scatter), so the map equals the reference's row for row;
//   * the table update is deterministic: of all candidates that hash to the same slot the LAST one in candidate order
//     decides the slot's owner, which is what the reference's (sequential, CPU) index_put does; a CUDA index_put with
//     duplicates picks an arbitrary one.  Two passes: every candidate tags its slot with atomicMin(-(i + 2)), then the
//     candidate that finds its own tag writes the owner (its new id, or the old owner it read before).
#include <algorithm>

#include "scan.cuh"

namespace pinb {

constexpr int MG_TPB = 256;

struct GrowParams {
  const float* cand;
  long long n;
  int32_t* table;
  long long buffer_size;
  float resolution;
  float* points;
  float* orient;
  int32_t* ts_create;
  int32_t* ts_update;
  float* certainty;
  long long n_points;
  const float* travel_dist;
  int cur_ts, grow_all, temporal_on;
  float diff_travel, far2;
  int* flag;   // [n] 1 = becomes a new neural point
  int* slot;   // [n]
  int* owner;  // [n] table[slot] before this frame
  int* bsum;   // [blocks]
  long long* n_new;
};

__device__ __forceinline__ bool grow_test(const GrowParams& p, long long i, int& slot, int& owner) {
  const float x = p.cand[3 * i], y = p.cand[3 * i + 1], z = p.cand[3 * i + 2];
  // fp32 true division then floor (model/neural_points.py:334), int64 prime hash, floor-mod (:340-343)
  const long long cx = (long long)floorf(__fdiv_rn(x, p.resolution)), cy = (long long)floorf(__fdiv_rn(y, p.resolution)),
                  cz = (long long)floorf(__fdiv_rn(z, p.resolution));
  slot = (int)floormod_u32(cx * PRIME0 + cy * PRIME1 + cz * PRIME2, p.buffer_size);
  owner = p.table[slot];
  if (p.grow_all || owner < 0) return true;
  const float dx = __fsub_rn(p.points[3 * (size_t)owner], x), dy = __fsub_rn(p.points[3 * (size_t)owner + 1], y),
              dz = __fsub_rn(p.points[3 * (size_t)owner + 2], z);
  const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  if (d2 > p.far2) return true;  // the slot belongs to another voxel (hash collision) or a far point (:352)
  if (p.temporal_on) {           // not refreshed within the local travel distance (:355-359)
    const float age = __fsub_rn(p.travel_dist[p.cur_ts], p.travel_dist[p.ts_update[owner]]);
    if (age > p.diff_travel) return true;
  }
  return false;
}

__global__ void __launch_bounds__(MG_TPB) grow_flag_kernel(const GrowParams p) {
  __shared__ int s_warp[64];
  const long long i = (long long)blockIdx.x * MG_TPB + threadIdx.x;
  int f = 0;
  if (i < p.n) {
    int slot, owner;
    f = grow_test(p, i, slot, owner) ? 1 : 0;
    p.flag[i] = f;
    p.slot[i] = slot;
    p.owner[i] = owner;
  }
  int total;
  block_exclusive_scan(f, s_warp, total);
  if (threadIdx.x == 0) p.bsum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) grow_scan_kernel(const GrowParams p, int n_blocks) {
  __shared__ int s_warp[64];
  int carry = 0;
  for (int base = 0; base < n_blocks; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < n_blocks ? p.bsum[i] : 0;
    int total;
    const int ex = block_exclusive_scan(v, s_warp, total);
    if (i < n_blocks) p.bsum[i] = carry + ex;
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) p.n_new[0] = carry;
}

// append the new points in candidate order; tag every candidate's slot (the largest candidate index wins)
__global__ void __launch_bounds__(MG_TPB) grow_append_kernel(const GrowParams p) {
  __shared__ int s_warp[64];
  const long long i = (long long)blockIdx.x * MG_TPB + threadIdx.x;
  const int f = i < p.n ? p.flag[i] : 0;
  int total;
  const long long rank = (long long)p.bsum[blockIdx.x] + block_exclusive_scan(f, s_warp, total);
  if (i >= p.n) return;
  if (f) {
    const long long id = p.n_points + rank;
    p.points[3 * id] = p.cand[3 * i];
    p.points[3 * id + 1] = p.cand[3 * i + 1];
    p.points[3 * id + 2] = p.cand[3 * i + 2];
    reinterpret_cast<float4*>(p.orient)[id] = make_float4(1.f, 0.f, 0.f, 0.f);
    p.ts_create[id] = p.cur_ts;
    p.ts_update[id] = p.cur_ts;
    p.certainty[id] = 0.f;
    p.flag[i] = (int)id + 1;  // from here on: new id + 1 (0 = not growing)
  }
  atomicMin(p.table + p.slot[i], -(int)(i + 2));
}

__global__ void __launch_bounds__(MG_TPB) grow_table_kernel(const GrowParams p) {
  const long long i = (long long)blockIdx.x * MG_TPB + threadIdx.x;
  if (i >= p.n) return;
  const int s = p.slot[i];
  if (p.table[s] == -(int)(i + 2)) p.table[s] = p.flag[i] ? p.flag[i] - 1 : p.owner[i];
}

}  // namespace pinb

using namespace pinb;

extern "C" int64_t pinb200_map_grow_scratch(int64_t n) {  // int32 elements
  return n <= 0 ? 0 : 3 * n + (n + MG_TPB - 1) / MG_TPB + 8;
}

extern "C" int pinb200_map_grow(const float* cand, int64_t n, int32_t* table, int64_t buffer_size, float resolution,
                                float* points, float* orient, int32_t* ts_create, int32_t* ts_update, float* certainty,
                                int64_t n_points, int64_t capacity, const float* travel_dist, int32_t cur_ts, int32_t grow_all,
                                int32_t temporal_on, float diff_travel, float far2, int32_t* scratch, int64_t* n_new, void* stream) {
  if (!cand || !table || !points || !orient || !ts_create || !ts_update || !certainty || !scratch || !n_new || n < 0 ||
      buffer_size <= 0 || buffer_size >= (1LL << 31) || (temporal_on && !grow_all && !travel_dist)) {
    set_error("map_grow: bad argument");
    return PINB200_ERR_BAD_ARG;
  }
  if (n_points + n > capacity || n_points + n >= (1LL << 30)) {
    set_error("map_grow: arenas of %lld rows cannot take %lld + %lld points", (long long)capacity, (long long)n_points, (long long)n);
    return PINB200_ERR_BAD_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    cudaMemsetAsync(n_new, 0, 8, st);
    return PINB200_OK;
  }
  const int n_blocks = (int)((n + MG_TPB - 1) / MG_TPB);
  GrowParams p{};
  p.cand = cand;
  p.n = n;
  p.table = table;
  p.buffer_size = buffer_size;
  p.resolution = resolution;
  p.points = points;
  p.orient = orient;
  p.ts_create = ts_create;
  p.ts_update = ts_update;
  p.certainty = certainty;
  p.n_points = n_points;
  p.travel_dist = travel_dist;
  p.cur_ts = cur_ts;
  p.grow_all = grow_all;
  p.temporal_on = temporal_on;
  p.diff_travel = diff_travel;
  p.far2 = far2;  // `3 * res**2` evaluated by the caller in double and cast to fp32, like the reference's python scalar
  p.flag = scratch;
  p.slot = scratch + n;
  p.owner = scratch + 2 * n;
  p.bsum = scratch + 3 * n;
  p.n_new = reinterpret_cast<long long*>(n_new);
  grow_flag_kernel<<<n_blocks, MG_TPB, 0, st>>>(p);
  grow_scan_kernel<<<1, 1024, 0, st>>>(p, n_blocks);
  grow_append_kernel<<<n_blocks, MG_TPB, 0, st>>>(p);
  grow_table_kernel<<<n_blocks, MG_TPB, 0, st>>>(p);
  return check_launch("map_grow");
}
