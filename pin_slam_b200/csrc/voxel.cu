// Voxel down-sampling: utils/tools.py:583-626 voxel_down_sample_torch and :629-668
// voxel_down_sample_min_value_torch of the reference, used three times per frame (dataset/slam_dataset.py:434,480,
// model/neural_points.py:331) and once per rehash.
//
// The reference (and round 1 of this repo) runs `torch.unique(return_inverse=True)` over one key per point: a full
// radix sort of the frame plus several host synchronisations -- 12 ms of host time per call on the benchmark scans,
// by far the largest item of the per-frame "mapping preparation".  Here the per-voxel winner is found with a
// lock-free open-addressing hash set keyed by the voxel key (atomicCAS on the key, atomicMin on the packed
// (quantised value, point index) pair), the occupied entries are compacted, and only the few thousand (key, winner)
// pairs that survive are sorted by key (the reference returns the winners in ascending key order, and the order
// matters: it becomes the id order of new neural points).  Selection rule, quantisation and output order are the
// reference's, bit for bit (IEEE fp32 operations in the same order).
#include <algorithm>

#include "common.cuh"

namespace pinb {

struct VoxelScalars {        // device scratch, zero-initialised by the call
  int gmin[3];               // per-axis minimum of floor(p / voxel) (origin of the key grid)
  int gmax[3];               // per-axis maximum
  unsigned int vmax_bits;    // max of the (non-negative) selection value, float bits
  int count;                 // number of occupied voxels (output)
};

__device__ __forceinline__ void cell_of(const float* __restrict__ p, float voxel, int& gx, int& gy, int& gz) {
  gx = (int)floorf(__fdiv_rn(p[0], voxel));
  gy = (int)floorf(__fdiv_rn(p[1], voxel));
  gz = (int)floorf(__fdiv_rn(p[2], voxel));
}

// distance of the point to the centre of its voxel, reference arithmetic (tools.py:598-600)
__device__ __forceinline__ float centre_dist(const float* __restrict__ p, float voxel) {
  float s = 0.f;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float g = floorf(__fdiv_rn(p[a], voxel));
    const float c = __fmul_rn(__fadd_rn(g, 0.5f), voxel);
    const float d = __fsub_rn(p[a], c);
    s = a == 0 ? __fmul_rn(d, d) : __fadd_rn(s, __fmul_rn(d, d));
  }
  return __fsqrt_rn(s);
}

__global__ void voxel_bounds_kernel(const float* __restrict__ pts, const float* __restrict__ value, long long n,
                                    float voxel, VoxelScalars* sc) {
  int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
  unsigned int vm = 0u;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int g[3];
    cell_of(pts + 3 * i, voxel, g[0], g[1], g[2]);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      lo[a] = min(lo[a], g[a]);
      hi[a] = max(hi[a], g[a]);
    }
    const float v = value ? value[i] : centre_dist(pts + 3 * i, voxel);
    vm = max(vm, __float_as_uint(v));  // v >= 0: unsigned order == float order
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    lo[a] = __reduce_min_sync(FULL, lo[a]);
    hi[a] = __reduce_max_sync(FULL, hi[a]);
  }
  vm = __reduce_max_sync(FULL, vm);
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      atomicMin(&sc->gmin[a], lo[a]);
      atomicMax(&sc->gmax[a], hi[a]);
    }
    atomicMax(&sc->vmax_bits, vm);
  }
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}

__global__ void voxel_insert_kernel(const float* __restrict__ pts, const float* __restrict__ value, long long n,
                                    float voxel, const VoxelScalars* __restrict__ sc, long long* keys,
                                    unsigned long long* best, unsigned long long mask) {
  const long long ox = sc->gmin[0], oy = sc->gmin[1], oz = sc->gmin[2];
  // v = g.max() over all three axes of the offset grid (tools.py:606-608)
  const long long v = max(max((long long)sc->gmax[0] - ox, (long long)sc->gmax[1] - oy), (long long)sc->gmax[2] - oz);
  const float vmax = __uint_as_float(sc->vmax_bits);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int gx, gy, gz;
    cell_of(pts + 3 * i, voxel, gx, gy, gz);
    const long long key = (gx - ox) + (gy - oy) * v + (gz - oz) * v * v;
    const float val = value ? value[i] : centre_dist(pts + 3 * i, voxel);
    const long long q = (long long)__fmul_rn(__fdiv_rn(val, vmax), 999.f);  // quantised to 1000 levels
    const unsigned long long packed = (unsigned long long)q * (unsigned long long)n + (unsigned long long)i;
    unsigned long long s = mix64((unsigned long long)key) & mask;
    while (true) {
      const long long prev = (long long)atomicCAS((unsigned long long*)&keys[s], (unsigned long long)-1LL, (unsigned long long)key);
      if (prev == -1LL || prev == key) {
        atomicMin(&best[s], packed);
        break;
      }
      s = (s + 1) & mask;
    }
  }
}

__global__ void voxel_compact_kernel(const long long* __restrict__ keys, const unsigned long long* __restrict__ best,
                                     long long table, long long n, VoxelScalars* sc, long long* out_key,
                                     long long* out_idx) {
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < table; s += (long long)gridDim.x * blockDim.x) {
    const long long k = keys[s];
    if (k != -1LL) {
      const int pos = atomicAdd(&sc->count, 1);
      out_key[pos] = k;
      out_idx[pos] = (long long)(best[s] % (unsigned long long)n);  // the winner's point index
    }
  }
}

}  // namespace pinb

using namespace pinb;

extern "C" int64_t pinb200_voxel_table_size(int64_t n) {
  int64_t t = 1024;
  while (t < 2 * n) t <<= 1;
  return t;
}

extern "C" int pinb200_voxel_downsample(const float* points, int64_t n, float voxel_size, const float* value,
                                        int64_t* ws_keys, uint64_t* ws_best, int64_t table_size, int32_t* scalars,
                                        int64_t* out_key, int64_t* out_idx, void* stream) {
  if (n <= 0) return PINB200_OK;
  if (!points || !ws_keys || !ws_best || !scalars || !out_key || !out_idx || voxel_size <= 0.f ||
      table_size < 2 * n || (table_size & (table_size - 1)) != 0) {
    set_error("voxel_downsample: bad argument (table_size must be a power of two >= 2 n)");
    return PINB200_ERR_BAD_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  VoxelScalars init;
  for (int a = 0; a < 3; ++a) {
    init.gmin[a] = INT_MAX;
    init.gmax[a] = INT_MIN;
  }
  init.vmax_bits = 0u;
  init.count = 0;
  cudaError_t e = cudaMemcpyAsync(scalars, &init, sizeof(init), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(ws_keys, 0xff, (size_t)table_size * 8, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(ws_best, 0xff, (size_t)table_size * 8, st);
  if (e != cudaSuccess) {
    set_error("voxel_downsample: %s", cudaGetErrorString(e));
    return PINB200_ERR_CUDA;
  }
  VoxelScalars* sc = reinterpret_cast<VoxelScalars*>(scalars);
  const int grid = (int)std::min<long long>((n + 255) / 256, (long long)sm_count() * 8);
  voxel_bounds_kernel<<<grid, 256, 0, st>>>(points, value, n, voxel_size, sc);
  voxel_insert_kernel<<<grid, 256, 0, st>>>(points, value, n, voxel_size, sc, (long long*)ws_keys,
                                            (unsigned long long*)ws_best, (unsigned long long)(table_size - 1));
  const int grid_t = (int)std::min<long long>((table_size + 255) / 256, (long long)sm_count() * 8);
  voxel_compact_kernel<<<grid_t, 256, 0, st>>>((const long long*)ws_keys, (const unsigned long long*)ws_best, table_size,
                                               n, sc, (long long*)out_key, (long long*)out_idx);
  return check_launch("voxel_downsample");
}
