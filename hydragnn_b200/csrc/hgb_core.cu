// libhgb.so -- error plumbing, prefix scan, CSR construction.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "hgb_common.cuh"

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void hgb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void hgb_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" int hgb_version(void) { return 100; }
extern "C" const char* hgb_last_error(void) { return g_err; }
extern "C" int64_t hgb_launch_count(void) { return g_launches.load(); }

// ------------------------------------------------------------------------------------------------
// exclusive scan (three small kernels; n <= 2^31).  Block = 1024 items.
// ------------------------------------------------------------------------------------------------
#define SCAN_B 1024

__device__ __forceinline__ int block_exclusive_scan_1024(int v, int* total) {
  __shared__ int warp_tot[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int w = warp_tot[lane];
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    warp_tot[lane] = winc - w;  // exclusive over warps
    if (lane == 31) *total = winc;
  }
  __syncthreads();
  return inc - v + warp_tot[wid];
}

__global__ void scan_block_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out, int64_t n,
                                  int32_t* __restrict__ block_sums) {
  __shared__ int total;
  int64_t i = (int64_t)blockIdx.x * SCAN_B + threadIdx.x;
  int v = i < n ? in[i] : 0;
  int ex = block_exclusive_scan_1024(v, &total);
  if (i < n) out[i] = ex;
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block: exclusive scan of block sums in place (nb arbitrary: serial over chunks of 1024)
__global__ void scan_sums_kernel(int32_t* __restrict__ sums, int nb, int32_t* __restrict__ grand_total) {
  __shared__ int total;
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += SCAN_B) {
    int i = base + threadIdx.x;
    int v = i < nb ? sums[i] : 0;
    int ex = block_exclusive_scan_1024(v, &total);
    int c = carry;
    if (i < nb) sums[i] = ex + c;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *grand_total = carry;
}

__global__ void scan_add_kernel(int32_t* __restrict__ out, int64_t n, const int32_t* __restrict__ sums,
                                const int32_t* __restrict__ grand_total) {
  int64_t i = (int64_t)blockIdx.x * SCAN_B + threadIdx.x;
  if (i < n) out[i] += sums[blockIdx.x];
  if (i == 0) out[n] = *grand_total;
}

extern "C" int64_t hgb_exclusive_scan_workspace_bytes(int64_t n) { return 4 * ((n + SCAN_B - 1) / SCAN_B + 2); }

extern "C" int hgb_exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, void* workspace,
                                      hgb_stream_t stream) {
  HGB_REQUIRE(n >= 0 && out && workspace, "exclusive_scan: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  int nb = (int)((n + SCAN_B - 1) / SCAN_B);
  int32_t* sums = (int32_t*)workspace;
  int32_t* total = sums + nb + 1;
  if (nb == 0) {
    cudaMemsetAsync(out, 0, sizeof(int32_t), st);
    return HGB_OK;
  }
  scan_block_kernel<<<nb, SCAN_B, 0, st>>>(in, out, n, sums);
  HGB_LAUNCH_CHECK("scan_block");
  scan_sums_kernel<<<1, SCAN_B, 0, st>>>(sums, nb, total);
  HGB_LAUNCH_CHECK("scan_sums");
  scan_add_kernel<<<nb, SCAN_B, 0, st>>>(out, n, sums, total);
  HGB_LAUNCH_CHECK("scan_add");
  return HGB_OK;
}

// ------------------------------------------------------------------------------------------------
// CSR view of an index vector (stable)
// ------------------------------------------------------------------------------------------------
__global__ void csr_hist_kernel(const int64_t* __restrict__ idx, int64_t e, int32_t n, int32_t* __restrict__ idx32,
                                int32_t* __restrict__ count, int32_t* __restrict__ bad) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < e; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t k = idx[i];
    if (k < 0 || k >= n) {
      *bad = 1;
      k = 0;
    }
    idx32[i] = (int32_t)k;
    atomicAdd(&count[k], 1);
  }
}

__global__ void csr_fill_kernel(const int32_t* __restrict__ idx32, int64_t e, const int32_t* __restrict__ rowptr,
                                int32_t* __restrict__ cursor, int32_t* __restrict__ perm) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < e; i += (int64_t)gridDim.x * blockDim.x) {
    int k = idx32[i];
    int slot = atomicAdd(&cursor[k], 1);
    perm[rowptr[k] + slot] = (int32_t)i;
  }
}

// restore ascending edge id inside every segment (segments are short: insertion sort per thread)
__global__ void csr_sort_kernel(const int32_t* __restrict__ rowptr, int32_t n, int32_t* __restrict__ perm) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    int lo = rowptr[k], hi = rowptr[k + 1];
    for (int a = lo + 1; a < hi; ++a) {
      int v = perm[a];
      int b = a - 1;
      while (b >= lo && perm[b] > v) {
        perm[b + 1] = perm[b];
        --b;
      }
      perm[b + 1] = v;
    }
  }
}

extern "C" int64_t hgb_csr_workspace_bytes(int64_t e, int32_t n) {
  return 4 * ((int64_t)n + 8) * 2 + hgb_exclusive_scan_workspace_bytes(n) + 64;
}

extern "C" int hgb_csr_build(const int64_t* idx, int64_t e, int32_t n, int32_t* idx32, int32_t* rowptr,
                             int32_t* perm, void* workspace, hgb_stream_t stream) {
  HGB_REQUIRE(e >= 0 && n >= 0 && rowptr && workspace, "csr_build: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  int32_t* count = (int32_t*)workspace;           // n + 8 (last slot: bad flag)
  int32_t* cursor = count + n + 8;                // n + 8
  void* scan_ws = (void*)(cursor + n + 8);
  cudaMemsetAsync(count, 0, 4 * ((int64_t)n + 8) * 2, st);
  if (e > 0) {
    csr_hist_kernel<<<hgb_grid_for(e, 256), 256, 0, st>>>(idx, e, n, idx32, count, count + n);
    HGB_LAUNCH_CHECK("csr_hist");
  }
  int rc = hgb_exclusive_scan_i32(count, rowptr, n, scan_ws, stream);
  if (rc) return rc;
  if (e > 0) {
    csr_fill_kernel<<<hgb_grid_for(e, 256), 256, 0, st>>>(idx32, e, rowptr, cursor, perm);
    HGB_LAUNCH_CHECK("csr_fill");
    csr_sort_kernel<<<hgb_grid_for(n, 128), 128, 0, st>>>(rowptr, n, perm);
    HGB_LAUNCH_CHECK("csr_sort");
  }
  return HGB_OK;
}

// out[p] = idx[perm[p]]  (the neighbour of every CSR slot, so the fused kernels do one dependent index load less)
__global__ void gather_i32_kernel(const int32_t* __restrict__ idx, const int32_t* __restrict__ perm, int64_t e, int32_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < e; i += (int64_t)gridDim.x * blockDim.x) out[i] = idx[perm[i]];
}
extern "C" int hgb_gather_i32(const int32_t* idx, const int32_t* perm, int64_t e, int32_t* out, hgb_stream_t stream) {
  if (e == 0) return HGB_OK;
  HGB_REQUIRE(idx && perm && out, "gather_i32: null pointer");
  gather_i32_kernel<<<hgb_grid_for(e, 256), 256, 0, (cudaStream_t)stream>>>(idx, perm, e, out);
  HGB_LAUNCH_CHECK("gather_i32");
  return HGB_OK;
}

// ------------------------------------------------------------------------------------------------
// device-side collate (replaces the index bookkeeping of PyG Batch.from_data_list, hydragnn/preprocess/load_data.py:157-164):
// batch[i] = graph of node i, and edge_index with per-graph node offsets added, both from [G+1] offset vectors.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int ptr_search(const int32_t* __restrict__ ptr, int g, int64_t i) {   // largest k with ptr[k] <= i
  int lo = 0, hi = g;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (ptr[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void collate_batch_vector_kernel(const int32_t* __restrict__ ptr, int g, int64_t n, int64_t* __restrict__ batch) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) batch[i] = ptr_search(ptr, g, i);
}

__global__ void collate_offset_edges_kernel(const int64_t* __restrict__ local, const int32_t* __restrict__ eptr,
                                            const int32_t* __restrict__ nptr, int g, int64_t e, int64_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < e; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t off = nptr[ptr_search(eptr, g, i)];
    out[i] = local[i] + off;
    out[e + i] = local[e + i] + off;
  }
}

extern "C" int hgb_collate_batch_vector(const int32_t* ptr, int32_t g, int64_t n, int64_t* batch, hgb_stream_t stream) {
  HGB_REQUIRE(g >= 0 && n >= 0 && ptr && (n == 0 || batch), "collate_batch_vector: bad arguments");
  if (n == 0) return HGB_OK;
  collate_batch_vector_kernel<<<hgb_grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(ptr, g, n, batch);
  HGB_LAUNCH_CHECK("collate_batch_vector");
  return HGB_OK;
}

extern "C" int hgb_collate_offset_edges(const int64_t* edge_index_local, const int32_t* edge_ptr, const int32_t* node_ptr, int32_t g,
                                        int64_t e, int64_t* edge_index, hgb_stream_t stream) {
  HGB_REQUIRE(g >= 0 && e >= 0 && edge_ptr && node_ptr && (e == 0 || (edge_index_local && edge_index)), "collate_offset_edges: bad arguments");
  if (e == 0) return HGB_OK;
  collate_offset_edges_kernel<<<hgb_grid_for(e, 256), 256, 0, (cudaStream_t)stream>>>(edge_index_local, edge_ptr, node_ptr, g, e, edge_index);
  HGB_LAUNCH_CHECK("collate_offset_edges");
  return HGB_OK;
}
