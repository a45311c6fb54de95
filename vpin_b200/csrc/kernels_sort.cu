// SPARK memory-checking timestamps on the device (Spartan/src/sparse_mlpoly.rs:232-265, AddrTimestamps::new).
//
// The reference replays the 3N operations (matrices A, B, C in that order, entries in COO order, padding entries address 0)
// sequentially against one counter per memory cell: read_ts[g] = counter[addr[g]]++ and audit_ts = the final counters.
// Equivalently read_ts[g] is the number of EARLIER operations with the same address, i.e. the rank of g inside its address
// group when the operations are stably sorted by address. That is what runs here: a least-significant-digit radix sort
// (8 bits per pass, stable) of (address, g), one pass marking where each address group starts, one pass writing
// rank = position - group start back to slot g; audit_ts is a plain histogram. No host replay, no host<->device copies.
#include "launch_count.hpp"
#include <atomic>

#include "kernels_poly.cuh"

namespace vpin {


namespace {

const int kSortThreads = 256;                          // 8 warps
const int kSortItems = 16;                             // elements per thread
const int kSortTile = kSortThreads * kSortItems;       // 4096 elements per block
const int kSortWarpChunk = 32 * kSortItems;            // a warp owns 512 consecutive elements of the tile

// keys[g] = address of operation g (0 for padding), vals[g] = g. (The audit counts come out of the sorted order in
// k_ts_ranks: a histogram with atomics here serialises on the hot addresses — the constant-1 column is touched by a third
// of all operations, address 0 by every padding slot — and took 0.6 ms per side at 3 x 2^20 operations.)
__global__ void __launch_bounds__(256) k_ts_init(const uint32_t *a0, const uint32_t *a1, const uint32_t *a2, size_t n0, size_t n1, size_t n2,
                                                 size_t N, uint32_t *keys, uint32_t *vals) {
  size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= 3 * N) return;
  size_t k = g / N, i = g - k * N;
  const uint32_t *a = k == 0 ? a0 : (k == 1 ? a1 : a2);
  size_t n = k == 0 ? n0 : (k == 1 ? n1 : n2);
  uint32_t key = i < n ? a[i] : 0u;
  keys[g] = key;
  vals[g] = (uint32_t)g;
}
// hist[digit * tiles + tile] = number of keys of the tile with that digit
__global__ void __launch_bounds__(kSortThreads) k_sort_hist(const uint32_t *keys, size_t n, int shift, uint32_t *hist, size_t tiles) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  size_t base = (size_t)blockIdx.x * kSortTile;
  for (int it = 0; it < kSortItems; it++) {
    size_t p = base + (size_t)it * kSortThreads + threadIdx.x;
    if (p < n) atomicAdd(&h[(keys[p] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * tiles + blockIdx.x] = h[threadIdx.x];
}
// ---- in-place exclusive prefix sum of n 32-bit counters, three launches: (1) every block scans its 4096-element chunk and
// records the chunk total, (2) one block scans the chunk totals, (3) every block adds its chunk's offset ----
const int kScanChunk = 4096;  // 1024 threads x 4
__device__ __forceinline__ uint32_t block_exclusive_scan4(uint32_t (&v)[4], uint32_t *total) {
  __shared__ uint32_t warp_sums[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t t = v[0] + v[1] + v[2] + v[3], incl = t;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    uint32_t o = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += o;
  }
  __syncthreads();  // warp_sums may still be read by a previous call
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = warp_sums[lane], wi = w;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      uint32_t o = __shfl_up_sync(0xffffffffu, wi, off);
      if (lane >= off) wi += o;
    }
    warp_sums[lane] = wi - w;  // exclusive
    if (lane == 31 && total) *total = wi;
  }
  __syncthreads();
  return warp_sums[warp] + incl - t;  // exclusive prefix of this thread's first element
}
__global__ void __launch_bounds__(1024) k_scan_chunks(uint32_t *a, size_t n, uint32_t *chunk_totals) {
  size_t i = (size_t)blockIdx.x * kScanChunk + (size_t)threadIdx.x * 4;
  uint32_t v[4];
#pragma unroll
  for (int k = 0; k < 4; k++) v[k] = i + k < n ? a[i + k] : 0u;
  __shared__ uint32_t total;
  uint32_t excl = block_exclusive_scan4(v, &total);
#pragma unroll
  for (int k = 0; k < 4; k++) { if (i + k < n) a[i + k] = excl; excl += v[k]; }
  if (threadIdx.x == 0) chunk_totals[blockIdx.x] = total;
}
// one block: exclusive scan of m chunk totals (m is at most a few thousand per pass; loops with a carry beyond 4096)
__global__ void __launch_bounds__(1024) k_scan_totals(uint32_t *t, size_t m) {
  __shared__ uint32_t total, carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (size_t base = 0; base < m; base += kScanChunk) {
    size_t i = base + (size_t)threadIdx.x * 4;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = i + k < m ? t[i + k] : 0u;
    uint32_t excl = carry + block_exclusive_scan4(v, &total);
#pragma unroll
    for (int k = 0; k < 4; k++) { if (i + k < m) t[i + k] = excl; excl += v[k]; }
    __syncthreads();
    if (threadIdx.x == 0) carry += total;
    __syncthreads();
  }
}
__global__ void __launch_bounds__(1024) k_scan_add(uint32_t *a, size_t n, const uint32_t *chunk_offsets) {
  size_t i = (size_t)blockIdx.x * kScanChunk + (size_t)threadIdx.x * 4;
  uint32_t off = chunk_offsets[blockIdx.x];
#pragma unroll
  for (int k = 0; k < 4; k++)
    if (i + k < n) a[i + k] += off;
}
// stable scatter of one tile: a warp owns 512 consecutive elements and walks them 32 at a time, so the order inside
// the tile is (warp, iteration, lane) = the original order
__global__ void __launch_bounds__(kSortThreads) k_sort_scatter(const uint32_t *keys, const uint32_t *vals, size_t n, int shift,
                                                               const uint32_t *hist, size_t tiles, uint32_t *keys_out, uint32_t *vals_out) {
  __shared__ uint32_t cnt[8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 8 * 256; i += kSortThreads) (&cnt[0][0])[i] = 0;
  __syncthreads();
  size_t base = (size_t)blockIdx.x * kSortTile + (size_t)warp * kSortWarpChunk;
  uint32_t k[kSortItems], v[kSortItems];
#pragma unroll
  for (int it = 0; it < kSortItems; it++) {
    size_t p = base + (size_t)it * 32 + lane;
    bool ok = p < n;
    k[it] = ok ? keys[p] : 0xffffffffu;
    v[it] = ok ? vals[p] : 0u;
    uint32_t d = (k[it] >> shift) & 255u;
    uint32_t m = __match_any_sync(0xffffffffu, ok ? d : 256u + 0u);
    if (ok && lane == __ffs(m) - 1) cnt[warp][d] += __popc(m);
    __syncwarp();
  }
  __syncthreads();
  {  // thread d: turn the per-warp counts of digit d into starting offsets
    uint32_t d = threadIdx.x, run = hist[(size_t)d * tiles + blockIdx.x];
#pragma unroll
    for (int w = 0; w < 8; w++) { uint32_t c = cnt[w][d]; cnt[w][d] = run; run += c; }
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < kSortItems; it++) {
    size_t p = base + (size_t)it * 32 + lane;
    bool ok = p < n;
    uint32_t d = (k[it] >> shift) & 255u;
    uint32_t m = __match_any_sync(0xffffffffu, ok ? d : 256u);
    uint32_t pos = 0;
    if (ok) pos = cnt[warp][d] + __popc(m & ((1u << lane) - 1u));
    __syncwarp();
    if (ok && lane == __ffs(m) - 1) cnt[warp][d] += __popc(m);
    __syncwarp();
    if (ok) { keys_out[pos] = k[it]; vals_out[pos] = v[it]; }
  }
}
__global__ void __launch_bounds__(256) k_ts_starts(const uint32_t *keys, size_t n, uint32_t *start) {
  size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  uint32_t key = keys[p];
  if (p == 0 || keys[p - 1] != key) start[key] = (uint32_t)p;
}
// read_ts = rank inside the address group; audit_ts[address] = size of the group (written by the group's last element;
// addresses nobody touches keep the zero of the memset)
__global__ void __launch_bounds__(256) k_ts_ranks(const uint32_t *keys, const uint32_t *vals, size_t n, const uint32_t *start, uint32_t *addr_out,
                                                  uint32_t *read_ts, uint32_t *audit) {
  size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  uint32_t key = keys[p], g = vals[p];
  uint32_t rank = (uint32_t)p - start[key];
  addr_out[g] = key;
  read_ts[g] = rank;
  if (p + 1 == n || keys[p + 1] != key) audit[key] = rank + 1u;
}

}  // namespace

size_t exclusive_scan_scratch_words(size_t n) { return (n + kScanChunk - 1) / kScanChunk + 1; }
void launch_exclusive_scan_u32(uint32_t *d, size_t n, uint32_t *d_scratch, cudaStream_t st) {
  if (!n) return;
  size_t chunks = (n + kScanChunk - 1) / kScanChunk;
  ++g_kernel_launches, k_scan_chunks<<<(unsigned)chunks, 1024, 0, st>>>(d, n, d_scratch);
  if (chunks > 1) {
    ++g_kernel_launches, k_scan_totals<<<1, 1024, 0, st>>>(d_scratch, chunks);
    ++g_kernel_launches, k_scan_add<<<(unsigned)chunks, 1024, 0, st>>>(d, n, d_scratch);
  }
}
size_t spark_timestamps_scratch_words(size_t N, size_t M) {
  size_t n = 3 * N, tiles = (n + kSortTile - 1) / kSortTile;
  return 4 * n + 256 * tiles + M + exclusive_scan_scratch_words(256 * tiles);
}
void launch_spark_timestamps(const uint32_t *const addr[3], const size_t nnz[3], size_t N, size_t M, uint32_t *d_addr_out, uint32_t *d_read_ts,
                             uint32_t *d_audit_ts, uint32_t *d_scratch, cudaStream_t st) {
  const size_t n = 3 * N, tiles = (n + kSortTile - 1) / kSortTile;
  uint32_t *keys = d_scratch, *vals = keys + n, *keys2 = vals + n, *vals2 = keys2 + n, *hist = vals2 + n, *start = hist + 256 * tiles;
  uint32_t *scan_scratch = start + M;
  cudaMemsetAsync(d_audit_ts, 0, M * sizeof(uint32_t), st);
  unsigned eb = (unsigned)((n + 255) / 256);
  ++g_kernel_launches, k_ts_init<<<eb, 256, 0, st>>>(addr[0], addr[1], addr[2], nnz[0], nnz[1], nnz[2], N, keys, vals);
  int bits = 0;
  while (((size_t)1 << bits) < M) bits++;
  for (int shift = 0; shift < bits; shift += 8) {
    ++g_kernel_launches, k_sort_hist<<<(unsigned)tiles, kSortThreads, 0, st>>>(keys, n, shift, hist, tiles);
    launch_exclusive_scan_u32(hist, 256 * tiles, scan_scratch, st);
    ++g_kernel_launches, k_sort_scatter<<<(unsigned)tiles, kSortThreads, 0, st>>>(keys, vals, n, shift, hist, tiles, keys2, vals2);
    uint32_t *t = keys; keys = keys2; keys2 = t;
    t = vals; vals = vals2; vals2 = t;
  }
  ++g_kernel_launches, k_ts_starts<<<eb, 256, 0, st>>>(keys, n, start);
  ++g_kernel_launches, k_ts_ranks<<<eb, 256, 0, st>>>(keys, vals, n, start, d_addr_out, d_read_ts, d_audit_ts);
}

}  // namespace vpin
