// MoE routing / permutation kernels:
//   * moe_align_sort  : (expert-major, block-aligned) token sort -- sorted_token_ids, per-m-tile expert ids,
//                       per-expert offsets, padded total.  One CTA; counting sort in shared memory.
//   * gather_rows     : x_sorted[i] = x[sorted_token_ids[i] / topk_div]   (pad rows -> zeros)
//   * scatter_rows    : y[sorted_token_ids[i]] = y_sorted[i]
//   * topk_reduce     : out[t] = sum_j w[t, j] * y[t * topk + j]          (fp32 accumulate)
//   * bincount        : per-expert histogram
//
// Reference: csrc/lib/moe_utils.cu:65-315 (legacy, not compiled there), kernels/nvidia/moe_utils.py:145-508
// (calc_gather_scatter_index, reduce_topk), allgather_group_gemm.py:86-166 (calc_sorted_gather_index_kernel),
// threadblock_swizzle_ag_moe*.{py,cu} (tile order by arrival stage -- see `stage_of_rank` below).
#include "td/ptx.cuh"
#include "runtime/driver.h"

using namespace td;

namespace {

// sorted layout: for each expert e, its (token,k) pairs in increasing flat index order (optionally grouped by the
// all-gather arrival stage of the token's source rank first), padded with `pad_id` up to a multiple of block_m.
//   topk_ids        : [n] flat (token * topk + k) -> expert
//   sorted_ids      : [capacity]  flat pair index or pad_id
//   tile_expert     : [capacity / block_m] expert of each m-tile (or -1 for unused tail tiles)
//   expert_offsets  : [E + 1] start of each expert's padded segment
//   total_padded    : [1]
//   tokens_per_rank > 0: pairs of an expert are ordered by stage = (src_rank - my_rank) mod world  (AG arrival order)
// Stable multi-CTA counting sort in three small launches (a single-CTA version with a per-bucket insertion sort took 13 ms
// for 16 K pairs -- longer than the grouped GEMM it feeds):
//   hist  (one CTA per 1024 pairs): pad-fill a slice of sorted_ids, per-CTA bucket histogram -> cnt[bucket][cta]
//   scan  (one CTA)               : exclusive scan over (bucket, cta), padded expert segments, tile -> expert map
//   place (one CTA per 1024 pairs): in-CTA stable rank (earlier pairs of the same bucket) + scanned base -> sorted_ids
// bucket = expert * stages + stage; the result is deterministic (increasing flat index inside every bucket).
constexpr int kSortChunk = 1024;
constexpr int kSortMaxBuckets = 4096;

__device__ __forceinline__ int sort_bucket(int e, int flat, int E, int stages, int topk, int tokens_per_rank, int my_rank, int world) {
  if (e < 0 || e >= E) return -1;
  if (stages == 1) return e;
  const int src = (flat / topk) / tokens_per_rank;
  return e * stages + (src - my_rank + world) % world;
}

__global__ void __launch_bounds__(kSortChunk) moe_sort_hist_kernel(const int* __restrict__ topk_ids, int n, int E, int stages, int capacity,
                                                                    int pad_id, int* __restrict__ sorted_ids, int* __restrict__ cnt,
                                                                    int topk, int tokens_per_rank, int my_rank, int world) {
  __shared__ int hist[kSortMaxBuckets];
  const int B = E * stages, nblk = gridDim.x;
  for (int i = threadIdx.x; i < B; i += blockDim.x) hist[i] = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < capacity; i += nblk * blockDim.x) sorted_ids[i] = pad_id;
  __syncthreads();
  const int i = blockIdx.x * kSortChunk + threadIdx.x;
  if (i < n) {
    const int b = sort_bucket(topk_ids[i], i, E, stages, topk, tokens_per_rank, my_rank, world);
    if (b >= 0) atomicAdd(&hist[b], 1);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) cnt[static_cast<size_t>(b) * nblk + blockIdx.x] = hist[b];
}

__global__ void __launch_bounds__(1024) moe_sort_scan_kernel(int* __restrict__ cnt, int nblk, int E, int stages, int block_m, int capacity,
                                                             int* __restrict__ bucket_start, int* __restrict__ tile_expert,
                                                             int* __restrict__ expert_offsets, int* __restrict__ total_padded) {
  __shared__ int tot[kSortMaxBuckets];
  __shared__ int offs[kSortMaxBuckets + 1];     // padded start of every expert (E + 1 entries used)
  const int B = E * stages;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {      // exclusive prefix over the CTAs, in place
    int run = 0;
    int* row = cnt + static_cast<size_t>(b) * nblk;
    for (int j = 0; j < nblk; ++j) { const int c = row[j]; row[j] = run; run += c; }
    tot[b] = run;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int off = 0;
    for (int e = 0; e < E; ++e) {
      offs[e] = off;
      int c = 0;
      for (int s = 0; s < stages; ++s) { bucket_start[e * stages + s] = off + c; c += tot[e * stages + s]; }
      off += (c + block_m - 1) / block_m * block_m;
    }
    offs[E] = off;
    total_padded[0] = off;
  }
  __syncthreads();
  for (int e = threadIdx.x; e <= E; e += blockDim.x) expert_offsets[e] = offs[e];
  for (int t = threadIdx.x; t < capacity / block_m; t += blockDim.x) {
    const int row = t * block_m;
    int e = -1;
    if (row < offs[E]) {            // binary search: last expert whose segment starts at or before `row` and is non-empty there
      int lo = 0, hi = E;
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (offs[mid] <= row) lo = mid; else hi = mid; }
      e = lo;
    }
    tile_expert[t] = e;
  }
}

__global__ void __launch_bounds__(kSortChunk) moe_sort_place_kernel(const int* __restrict__ topk_ids, int n, int E, int stages,
                                                                     const int* __restrict__ pre, const int* __restrict__ bucket_start,
                                                                     int* __restrict__ sorted_ids, int topk, int tokens_per_rank,
                                                                     int my_rank, int world) {
  __shared__ int sb[kSortChunk];
  const int i = blockIdx.x * kSortChunk + threadIdx.x;
  const int b = (i < n) ? sort_bucket(topk_ids[i], i, E, stages, topk, tokens_per_rank, my_rank, world) : -1;
  sb[threadIdx.x] = b;
  __syncthreads();
  if (b < 0) return;
  int rank = 0;
  for (int j = 0; j < static_cast<int>(threadIdx.x); ++j) rank += (sb[j] == b);
  sorted_ids[bucket_start[b] + pre[static_cast<size_t>(b) * gridDim.x + blockIdx.x] + rank] = i;
}

// dst[i, :] = (ids[i] == pad) ? 0 : src[ids[i] / div, :]    (rows of `row_bytes`, multiple of 16)
__global__ void gather_rows_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, const int* __restrict__ ids,
                                   const int* __restrict__ n_rows_ptr, int n_rows_max, int vec_per_row, int div, int pad_id) {
  const int n_rows = n_rows_ptr ? min(*n_rows_ptr, n_rows_max) : n_rows_max;
  const long long total = static_cast<long long>(n_rows) * vec_per_row;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int r = static_cast<int>(t / vec_per_row), v = static_cast<int>(t % vec_per_row);
    const int id = ids[r];
    dst[t] = (id == pad_id) ? make_uint4(0, 0, 0, 0) : src[static_cast<long long>(id / div) * vec_per_row + v];
  }
}
// dst[ids[i], :] = src[i, :]   for ids[i] != pad
__global__ void scatter_rows_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, const int* __restrict__ ids,
                                    const int* __restrict__ n_rows_ptr, int n_rows_max, int vec_per_row, int pad_id) {
  const int n_rows = n_rows_ptr ? min(*n_rows_ptr, n_rows_max) : n_rows_max;
  const long long total = static_cast<long long>(n_rows) * vec_per_row;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int r = static_cast<int>(t / vec_per_row), v = static_cast<int>(t % vec_per_row);
    const int id = ids[r];
    if (id != pad_id) dst[static_cast<long long>(id) * vec_per_row + v] = src[t];
  }
}

// dst[c, i] = ids[i] >= 0 ? src[ids[i], c] : 0   (16-bit elements): the token-major ("MN-major") operands of a weight-gradient
// product, re-laid out reduction-major so the K-major tcgen05 GEMM can consume them; pad positions become zero columns.
// 64 x 64 tile per CTA through shared memory: 16-byte global loads along src rows, 16-byte global stores along dst rows.
__global__ void __launch_bounds__(256) transpose_gather_kernel(uint16_t* __restrict__ dst, const uint16_t* __restrict__ src,
                                                              const int* __restrict__ ids, int n_out, int cols, long long ld_src) {
  __shared__ uint16_t tile[64][64 + 8];          // [c][i], row pitch 144 B keeps the 16-byte reads aligned
  const int i0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  for (int q = threadIdx.x; q < 64 * 8; q += 256) {
    const int r = q >> 3, ch = q & 7;            // row of the tile (position i0 + r), 8-column chunk
    const int i = i0 + r, c = c0 + ch * 8;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (i < n_out && c < cols) {
      const int id = ids ? ids[i] : i;
      if (id >= 0) v = *reinterpret_cast<const uint4*>(src + static_cast<size_t>(id) * ld_src + c);
    }
    const uint16_t* e = reinterpret_cast<const uint16_t*>(&v);
#pragma unroll
    for (int k = 0; k < 8; ++k) tile[ch * 8 + k][r] = e[k];
  }
  __syncthreads();
  for (int q = threadIdx.x; q < 64 * 8; q += 256) {
    const int cr = q >> 3, ch = q & 7;           // dst row c0 + cr, 8 consecutive positions
    const int c = c0 + cr, i = i0 + ch * 8;
    if (c < cols && i < n_out) *reinterpret_cast<uint4*>(dst + static_cast<size_t>(c) * n_out + i) = *reinterpret_cast<const uint4*>(&tile[cr][ch * 8]);
  }
}

template <bool kBF16>
__global__ void topk_reduce_kernel(uint4* __restrict__ out, const uint4* __restrict__ y, const float* __restrict__ w, int T,
                                   int topk, int vec_per_row) {
  const long long total = static_cast<long long>(T) * vec_per_row;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int tok = static_cast<int>(t / vec_per_row), v = static_cast<int>(t % vec_per_row);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < topk; ++j) {
      const uint4 x = y[(static_cast<long long>(tok) * topk + j) * vec_per_row + v];
      const float wj = w ? w[tok * topk + j] : 1.f;
      const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if constexpr (kBF16) { acc[2 * i] += wj * ptx::bf16_lo(xs[i]); acc[2 * i + 1] += wj * ptx::bf16_hi(xs[i]); }
        else { const __half2 h = *reinterpret_cast<const __half2*>(&xs[i]); acc[2 * i] += wj * __low2float(h); acc[2 * i + 1] += wj * __high2float(h); }
      }
    }
    uint4 o;
    if constexpr (kBF16) { o.x = ptx::pack_bf16x2(acc[0], acc[1]); o.y = ptx::pack_bf16x2(acc[2], acc[3]); o.z = ptx::pack_bf16x2(acc[4], acc[5]); o.w = ptx::pack_bf16x2(acc[6], acc[7]); }
    else { o.x = ptx::pack_f16x2(acc[0], acc[1]); o.y = ptx::pack_f16x2(acc[2], acc[3]); o.z = ptx::pack_f16x2(acc[4], acc[5]); o.w = ptx::pack_f16x2(acc[6], acc[7]); }
    out[t] = o;
  }
}

__global__ void bincount_kernel(const int* __restrict__ ids, int n, int* __restrict__ counts, int E) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int e = ids[i];
    if (e >= 0 && e < E) atomicAdd(counts + e, 1);
  }
}

}  // namespace

TD_API int td_moe_align_sort(const void* topk_ids, int n, int E, int block_m, int capacity, int pad_id, void* sorted_ids,
                             void* tile_expert, void* expert_offsets, void* total_padded, int topk, int tokens_per_rank,
                             int my_rank, int world, void* ws, long long ws_bytes, void* stream) {
  if (capacity % block_m) { td::drv::set_error("moe_align_sort: capacity must be a multiple of block_m"); return -1; }
  const int stages = tokens_per_rank > 0 ? world : 1;
  const int B = E * stages;
  if (B > kSortMaxBuckets) { td::drv::set_error("moe_align_sort: experts x ranks exceeds 4096 buckets"); return -1; }
  const int nblk = n > 0 ? (n + kSortChunk - 1) / kSortChunk : 1;
  const long long need = (static_cast<long long>(B) * nblk + B) * 4;
  if (!ws || ws_bytes < need) { td::drv::set_error("moe_align_sort: workspace too small"); return -1; }
  int* cnt = reinterpret_cast<int*>(ws);
  int* bucket_start = cnt + static_cast<size_t>(B) * nblk;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  moe_sort_hist_kernel<<<nblk, kSortChunk, 0, s>>>((const int*)topk_ids, n, E, stages, capacity, pad_id, (int*)sorted_ids, cnt, topk,
                                                   tokens_per_rank, my_rank, world);
  moe_sort_scan_kernel<<<1, 1024, 0, s>>>(cnt, nblk, E, stages, block_m, capacity, bucket_start, (int*)tile_expert,
                                          (int*)expert_offsets, (int*)total_padded);
  moe_sort_place_kernel<<<nblk, kSortChunk, 0, s>>>((const int*)topk_ids, n, E, stages, cnt, bucket_start, (int*)sorted_ids, topk,
                                                    tokens_per_rank, my_rank, world);
  TD_CUDA_CHECK(cudaGetLastError());
  return 0;
}

TD_API int td_gather_rows(void* dst, const void* src, const void* ids, const void* n_rows_ptr, int n_rows_max, long long row_bytes,
                          int div, int pad_id, void* stream) {
  if (row_bytes % 16) { td::drv::set_error("gather_rows: row size must be a multiple of 16 bytes"); return -1; }
  if (n_rows_max == 0) return 0;
  const int vpr = (int)(row_bytes / 16);
  const long long total = (long long)n_rows_max * vpr;
  const int grid = (int)min((long long)148 * 8, (total + 255) / 256);
  gather_rows_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>((uint4*)dst, (const uint4*)src, (const int*)ids,
                                                                                (const int*)n_rows_ptr, n_rows_max, vpr, div, pad_id);
  TD_CUDA_CHECK(cudaGetLastError());
  return 0;
}

TD_API int td_scatter_rows(void* dst, const void* src, const void* ids, const void* n_rows_ptr, int n_rows_max, long long row_bytes,
                           int pad_id, void* stream) {
  if (row_bytes % 16) { td::drv::set_error("scatter_rows: row size must be a multiple of 16 bytes"); return -1; }
  if (n_rows_max == 0) return 0;
  const int vpr = (int)(row_bytes / 16);
  const long long total = (long long)n_rows_max * vpr;
  const int grid = (int)min((long long)148 * 8, (total + 255) / 256);
  scatter_rows_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>((uint4*)dst, (const uint4*)src, (const int*)ids,
                                                                                 (const int*)n_rows_ptr, n_rows_max, vpr, pad_id);
  TD_CUDA_CHECK(cudaGetLastError());
  return 0;
}

TD_API int td_transpose_gather(void* dst, const void* src, const void* ids, int n_out, int cols, long long ld_src, void* stream) {
  if (n_out % 8 || cols % 8 || ld_src % 8) { td::drv::set_error("transpose_gather: sizes must be multiples of 8 elements"); return -1; }
  if (n_out == 0 || cols == 0) return 0;
  dim3 grid((n_out + 63) / 64, (cols + 63) / 64);
  transpose_gather_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>((uint16_t*)dst, (const uint16_t*)src, (const int*)ids,
                                                                                   n_out, cols, ld_src);
  TD_CUDA_CHECK(cudaGetLastError());
  return 0;
}

TD_API int td_topk_reduce(void* out, const void* y, const void* w, int T, int topk, int H, int is_bf16, void* stream) {
  if (H % 8) { td::drv::set_error("topk_reduce: H must be a multiple of 8"); return -1; }
  if (T == 0) return 0;
  const int vpr = H / 8;
  const long long total = (long long)T * vpr;
  const int grid = (int)min((long long)148 * 8, (total + 255) / 256);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (is_bf16) topk_reduce_kernel<true><<<grid, 256, 0, s>>>((uint4*)out, (const uint4*)y, (const float*)w, T, topk, vpr);
  else topk_reduce_kernel<false><<<grid, 256, 0, s>>>((uint4*)out, (const uint4*)y, (const float*)w, T, topk, vpr);
  TD_CUDA_CHECK(cudaGetLastError());
  return 0;
}

TD_API int td_bincount(const void* ids, int n, void* counts, int E, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  TD_CUDA_CHECK(cudaMemsetAsync(counts, 0, sizeof(int) * E, s));
  if (n) bincount_kernel<<<min(148, (n + 255) / 256), 256, 0, s>>>((const int*)ids, n, (int*)counts, E);
  TD_CUDA_CHECK(cudaGetLastError());
  return 0;
}
