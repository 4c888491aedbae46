// Probe index of a map view (pinb200_map_view.probe_words / probe_rec / probe_gid): a succinct rank structure over
// the reference's 5e7-slot hash table that K1 probes instead of buffer_pt_index.
//
// The reference's table is 400 MB of int64 with ~1e5..1e6 live slots; a probe reads slot -> point -> timestamp ->
// travel distance -> global2local (model/neural_points.py:963-999,573).  Random 4..16-byte reads into a table that is
// several times the 126 MB L2 are served by HBM at ~1.3 TB/s of 32-byte sectors (measured, scripts/micro/gather_bw.cu:
// 0.14 sectors/cycle/SM vs 0.8 for an L2-resident table) and thrash the 256 MB TLB reach.  The index keeps the
// table's ANSWERS in L2-resident form:
//   probe_words [ceil(B/32)][2] u32 : {occupancy bits of 32 consecutive slots, number of set bits before this word}
//                                     (12.5 MB for B = 5e7) -- bit s is set iff slot s is owned by a neural point this
//                                     view can return (id >= 0, inside the travel-distance window when time_filter)
//   probe_rec   [n_rec][4] f32      : {x, y, z, bit-cast id | REMAP} of the owners, ordered by slot (rank order)
//   probe_gid   [n_rec] i32         : their global ids
// lookup(slot): w = probe_words[slot >> 5]; hit = w.bits >> (slot & 31) & 1; rank = w.prefix + popc(w.bits & below).
// Built by five small launches whenever the map view changes (NeuralPoints.update / reset_local_map): mark, three-step
// prefix sum over the words, scatter.  Identical semantics to probing the reference table: the key is the reference
// slot id, so hash collisions, overwritten owners and the per-probe validity tests are preserved bit for bit.
#include <algorithm>

#include "common.cuh"
#include "scan.cuh"

namespace pinb {

constexpr int SCAN_TPB = 256;
constexpr int SCAN_WPT = 4;  // words per thread
constexpr int SCAN_WPBLK = SCAN_TPB * SCAN_WPT;

// id in the queried index space (with the REMAP flag) of global point i, or -1 if this view cannot return it
__device__ __forceinline__ int view_id_of(const pinb200_map_view& m, long long i, float x, float y, float z, float td_cur) {
  int id = m.global2local ? m.global2local[i] : (int)i;
  if (id < 0) return -1;
  if (m.time_filter && !(fabsf(td_cur - m.travel_dist[m.ts_create[i]]) < m.diff_travel_dist_local)) return -1;
  if (m.global2local && m.nb_points) {
    const float* lp = m.nb_points + 3 * (size_t)id;
    if (!(lp[0] == x && lp[1] == y && lp[2] == z)) id |= PINB200_REC_REMAP;
  }
  return id;
}

__global__ void probe_mark_kernel(const __grid_constant__ pinb200_map_view m, uint32_t* __restrict__ words) {
  const float td_cur = m.time_filter ? m.travel_dist[m.cur_ts] : 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m.n_global;
       i += (long long)gridDim.x * blockDim.x) {
    const float x = m.points[3 * i], y = m.points[3 * i + 1], z = m.points[3 * i + 2];
    const uint32_t slot = base_slot(m, x, y, z);
    if (m.slot_table[slot] != (int)i) continue;  // overwritten by a colliding point: unreachable, like in the reference
    if (view_id_of(m, i, x, y, z, td_cur) < 0) continue;
    atomicOr(words + 2 * (size_t)(slot >> 5), 1u << (slot & 31));
  }
}

__global__ void __launch_bounds__(SCAN_TPB) probe_blocksum_kernel(const uint32_t* __restrict__ words, long long n_words,
                                                                  int* __restrict__ bsum) {
  __shared__ int s_warp[64];
  const long long w0 = (long long)blockIdx.x * SCAN_WPBLK + threadIdx.x * SCAN_WPT;
  int c = 0;
#pragma unroll
  for (int j = 0; j < SCAN_WPT; ++j)
    if (w0 + j < n_words) c += __popc(words[2 * (w0 + j)]);
  int total;
  block_exclusive_scan(c, s_warp, total);
  if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) probe_scan_blocksums_kernel(int* __restrict__ bsum, int n_blocks) {
  __shared__ int s_warp[64];
  int carry = 0;
  for (int base = 0; base < n_blocks; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < n_blocks ? bsum[i] : 0;
    int total;
    const int ex = block_exclusive_scan(v, s_warp, total);
    if (i < n_blocks) bsum[i] = carry + ex;
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) bsum[n_blocks] = carry;  // number of records
}

__global__ void __launch_bounds__(SCAN_TPB) probe_prefix_kernel(uint32_t* __restrict__ words, long long n_words,
                                                                const int* __restrict__ bsum) {
  __shared__ int s_warp[64];
  const long long w0 = (long long)blockIdx.x * SCAN_WPBLK + threadIdx.x * SCAN_WPT;
  int pc[SCAN_WPT], c = 0;
#pragma unroll
  for (int j = 0; j < SCAN_WPT; ++j) {
    pc[j] = w0 + j < n_words ? __popc(words[2 * (w0 + j)]) : 0;
    c += pc[j];
  }
  int total;
  int run = bsum[blockIdx.x] + block_exclusive_scan(c, s_warp, total);
#pragma unroll
  for (int j = 0; j < SCAN_WPT; ++j) {
    if (w0 + j < n_words) words[2 * (w0 + j) + 1] = (uint32_t)run;
    run += pc[j];
  }
}

__global__ void probe_scatter_kernel(const __grid_constant__ pinb200_map_view m, const uint32_t* __restrict__ words,
                                     float4* __restrict__ rec, int32_t* __restrict__ gid) {
  const float td_cur = m.time_filter ? m.travel_dist[m.cur_ts] : 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m.n_global;
       i += (long long)gridDim.x * blockDim.x) {
    const float x = m.points[3 * i], y = m.points[3 * i + 1], z = m.points[3 * i + 2];
    const uint32_t slot = base_slot(m, x, y, z);
    if (m.slot_table[slot] != (int)i) continue;
    const int id = view_id_of(m, i, x, y, z, td_cur);
    if (id < 0) continue;
    const uint2 w = *reinterpret_cast<const uint2*>(words + 2 * (size_t)(slot >> 5));
    const uint32_t rank = w.y + __popc(w.x & ((1u << (slot & 31)) - 1u));
    rec[rank] = make_float4(x, y, z, __int_as_float(id));
    gid[rank] = (int)i;
  }
}

}  // namespace pinb

using namespace pinb;

extern "C" int64_t pinb200_probe_index_words(int64_t buffer_size) { return (buffer_size + 31) / 32; }

extern "C" int64_t pinb200_probe_index_scratch(int64_t buffer_size) {
  const int64_t n_words = (buffer_size + 31) / 32;
  return (n_words + SCAN_WPBLK - 1) / SCAN_WPBLK + 1;
}

extern "C" int pinb200_build_probe_index(const pinb200_map_view* map, uint32_t* probe_words, float* probe_rec,
                                         int32_t* probe_gid, int32_t* scratch, void* stream) {
  if (!map || !probe_words || !scratch || !map->slot_table || (map->n_global > 0 && (!map->points || !probe_rec || !probe_gid))) {
    set_error("build_probe_index: null argument");
    return PINB200_ERR_BAD_ARG;
  }
  if (map->buffer_size <= 0 || map->buffer_size >= (1LL << 31)) {
    set_error("build_probe_index: buffer_size %lld out of range (0, 2^31)", (long long)map->buffer_size);
    return PINB200_ERR_BAD_ARG;
  }
  if (map->time_filter && (!map->ts_create || !map->travel_dist)) {
    set_error("build_probe_index: time_filter needs ts_create and travel_dist");
    return PINB200_ERR_BAD_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const long long n_words = (map->buffer_size + 31) / 32;
  const int n_blocks = (int)((n_words + SCAN_WPBLK - 1) / SCAN_WPBLK);
  cudaError_t e = cudaMemsetAsync(probe_words, 0, (size_t)n_words * 8, st);
  if (e != cudaSuccess) {
    set_error("build_probe_index: cudaMemsetAsync: %s", cudaGetErrorString(e));
    return PINB200_ERR_CUDA;
  }
  const int grid_pts = (int)std::min<long long>(std::max<long long>(1, (map->n_global + 255) / 256), (long long)sm_count() * 16);
  if (map->n_global > 0) probe_mark_kernel<<<grid_pts, 256, 0, st>>>(*map, probe_words);
  probe_blocksum_kernel<<<n_blocks, SCAN_TPB, 0, st>>>(probe_words, n_words, scratch);
  probe_scan_blocksums_kernel<<<1, 1024, 0, st>>>(scratch, n_blocks);
  probe_prefix_kernel<<<n_blocks, SCAN_TPB, 0, st>>>(probe_words, n_words, scratch);
  if (map->n_global > 0)
    probe_scatter_kernel<<<grid_pts, 256, 0, st>>>(*map, probe_words, reinterpret_cast<float4*>(probe_rec), probe_gid);
  return check_launch("build_probe_index");
}
