// libhgb.so -- row gather, atomics-free segmented sum over a CSR view, graph pooling.
//
// All three are HBM-bound.  Thread mapping: a group of LANES = min(32, pow2 >= C/4) threads owns one
// output row and walks the row with float4 loads (16 B per thread, consecutive lanes -> consecutive
// 16 B, i.e. fully coalesced 128 B lines when C*4 >= 128), so narrow rows (C = 1..16) do not waste
// a whole warp per row.  Summation order inside a segment is the CSR order: deterministic.
#include "hgb_common.cuh"

template <int VEC>
struct VecT;
template <>
struct VecT<4> { using type = float4; };
template <>
struct VecT<1> { using type = float; };

__device__ __forceinline__ void vadd(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
__device__ __forceinline__ void vadd(float& a, const float& b) { a += b; }
__device__ __forceinline__ void vzero(float4& a) { a = make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void vzero(float& a) { a = 0.f; }

// picks the sub-warp group width for rows of `cv` vector elements
static inline int group_lanes(int cv) {
  int l = 1;
  while (l < cv && l < 32) l <<= 1;
  return l;
}

// ---- gather -----------------------------------------------------------------------------------
template <int VEC>
__global__ void gather_rows_kernel(const float* __restrict__ x, const int32_t* __restrict__ idx, int64_t e, int cv,
                                   int lanes, float* __restrict__ out) {
  using V = typename VecT<VEC>::type;
  const int gpb = blockDim.x / lanes;
  const int sub = threadIdx.x % lanes;
  for (int64_t row = (int64_t)blockIdx.x * gpb + threadIdx.x / lanes; row < e; row += (int64_t)gridDim.x * gpb) {
    const V* src = reinterpret_cast<const V*>(x) + (int64_t)idx[row] * cv;
    V* dst = reinterpret_cast<V*>(out) + row * cv;
    for (int c = sub; c < cv; c += lanes) dst[c] = __ldg(src + c);
  }
}

extern "C" int hgb_gather_rows(const float* x, const int32_t* idx, int64_t e, int32_t c, float* out, hgb_stream_t stream) {
  HGB_REQUIRE(e >= 0 && c > 0 && x && idx && out, "gather_rows: bad arguments");
  if (e == 0) return HGB_OK;
  const bool v4 = (c % 4 == 0) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)out % 16 == 0);
  const int cv = v4 ? c / 4 : c;
  const int lanes = group_lanes(cv);
  const int gpb = 256 / lanes;
  const int grid = hgb_grid_for(e, gpb);
  if (v4) gather_rows_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(x, idx, e, cv, lanes, out);
  else gather_rows_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(x, idx, e, cv, lanes, out);
  HGB_LAUNCH_CHECK("gather_rows");
  return HGB_OK;
}

// ---- segmented sum ------------------------------------------------------------------------------
template <int VEC>
__global__ void segment_sum_kernel(const float* __restrict__ m, const int32_t* __restrict__ rowptr,
                                   const int32_t* __restrict__ perm, int n, int cv, int lanes, float* __restrict__ out,
                                   int ldo_v /* output row stride in V units */) {
  using V = typename VecT<VEC>::type;
  const int gpb = blockDim.x / lanes;
  const int sub = threadIdx.x % lanes;
  for (int row = blockIdx.x * gpb + threadIdx.x / lanes; row < n; row += gridDim.x * gpb) {
    const int lo = rowptr[row], hi = rowptr[row + 1];
    for (int c = sub; c < cv; c += lanes) {
      V acc;
      vzero(acc);
      int p = lo;
      // two independent loads in flight per lane
      for (; p + 1 < hi; p += 2) {
        const int e0 = perm ? perm[p] : p, e1 = perm ? perm[p + 1] : p + 1;
        V a = __ldg(reinterpret_cast<const V*>(m) + (int64_t)e0 * cv + c);
        V b = __ldg(reinterpret_cast<const V*>(m) + (int64_t)e1 * cv + c);
        vadd(acc, a);
        vadd(acc, b);
      }
      if (p < hi) {
        const int e0 = perm ? perm[p] : p;
        vadd(acc, __ldg(reinterpret_cast<const V*>(m) + (int64_t)e0 * cv + c));
      }
      reinterpret_cast<V*>(out)[(int64_t)row * ldo_v + c] = acc;
    }
  }
}

extern "C" int hgb_segment_sum_strided(const float* m, const int32_t* rowptr, const int32_t* perm, int32_t n, int32_t c,
                                       float* out, int32_t ldo, hgb_stream_t stream) {
  HGB_REQUIRE(n >= 0 && c > 0 && rowptr && out && ldo >= c, "segment_sum: bad arguments");
  if (n == 0) return HGB_OK;
  const bool v4 = (c % 4 == 0) && (ldo % 4 == 0) && ((uintptr_t)m % 16 == 0) && ((uintptr_t)out % 16 == 0);
  const int cv = v4 ? c / 4 : c;
  const int lanes = group_lanes(cv);
  const int gpb = 256 / lanes;
  const int grid = hgb_grid_for(n, gpb);
  if (v4) segment_sum_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(m, rowptr, perm, n, cv, lanes, out, ldo / 4);
  else segment_sum_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(m, rowptr, perm, n, cv, lanes, out, ldo);
  HGB_LAUNCH_CHECK("segment_sum");
  return HGB_OK;
}

extern "C" int hgb_segment_sum(const float* m, const int32_t* rowptr, const int32_t* perm, int32_t n, int32_t c,
                               float* out, hgb_stream_t stream) {
  return hgb_segment_sum_strided(m, rowptr, perm, n, c, out, c, stream);
}

// ---- graph pooling (batch is sorted: a graph is a contiguous run of rows) -------------------------
__global__ void pool_fwd_kernel(const float* __restrict__ x, const int32_t* __restrict__ gptr, int g, int c, int mode,
                                float* __restrict__ out, int32_t* __restrict__ argmax) {
  // one warp per graph, lanes stride over channels (rows of a graph are contiguous -> coalesced)
  const int wpb = blockDim.x >> 5, lane = threadIdx.x & 31;
  for (int k = blockIdx.x * wpb + (threadIdx.x >> 5); k < g; k += gridDim.x * wpb) {
    const int lo = gptr[k], hi = gptr[k + 1];
    for (int ch = lane; ch < c; ch += 32) {
      if (mode == HGB_POOL_MAX) {
        float best = -INFINITY;
        int arg = -1;
        for (int i = lo; i < hi; ++i) {
          float v = x[(int64_t)i * c + ch];
          if (v > best || arg < 0) { if (v > best || arg < 0) { best = v; arg = i; } }
        }
        out[(int64_t)k * c + ch] = arg < 0 ? 0.f : best;
        if (argmax) argmax[(int64_t)k * c + ch] = arg;
      } else {
        float acc = 0.f;
        for (int i = lo; i < hi; ++i) acc += x[(int64_t)i * c + ch];
        if (mode == HGB_POOL_MEAN) acc /= (float)max(hi - lo, 1);
        out[(int64_t)k * c + ch] = acc;
      }
    }
  }
}

__global__ void pool_bwd_kernel(const float* __restrict__ gout, const int32_t* __restrict__ gptr,
                                const int32_t* __restrict__ argmax, int n, int g, int c, int mode, float* __restrict__ gx) {
  const int wpb = blockDim.x >> 5, lane = threadIdx.x & 31;
  for (int k = blockIdx.x * wpb + (threadIdx.x >> 5); k < g; k += gridDim.x * wpb) {
    const int lo = gptr[k], hi = gptr[k + 1];
    const float scale = mode == HGB_POOL_MEAN ? 1.f / (float)max(hi - lo, 1) : 1.f;
    for (int ch = lane; ch < c; ch += 32) {
      const float gv = gout[(int64_t)k * c + ch] * scale;
      if (mode == HGB_POOL_MAX) {
        const int arg = argmax[(int64_t)k * c + ch];
        for (int i = lo; i < hi; ++i) gx[(int64_t)i * c + ch] = (i == arg) ? gv : 0.f;
      } else {
        for (int i = lo; i < hi; ++i) gx[(int64_t)i * c + ch] = gv;
      }
    }
  }
}

extern "C" int hgb_pool_fwd(const float* x, const int32_t* graph_ptr, int32_t g, int32_t c, int32_t mode, float* out,
                            int32_t* argmax, hgb_stream_t stream) {
  HGB_REQUIRE(g >= 0 && c > 0 && graph_ptr && out && mode >= 0 && mode <= 2, "pool_fwd: bad arguments");
  HGB_REQUIRE(mode != HGB_POOL_MAX || argmax, "pool_fwd: max pooling needs an argmax buffer");
  if (g == 0) return HGB_OK;
  pool_fwd_kernel<<<hgb_grid_for(g, 8), 256, 0, (cudaStream_t)stream>>>(x, graph_ptr, g, c, mode, out, argmax);
  HGB_LAUNCH_CHECK("pool_fwd");
  return HGB_OK;
}

extern "C" int hgb_pool_bwd(const float* gout, const int32_t* graph_ptr, const int32_t* argmax, int32_t n, int32_t g,
                            int32_t c, int32_t mode, float* gx, hgb_stream_t stream) {
  HGB_REQUIRE(g >= 0 && c > 0 && graph_ptr && gx && mode >= 0 && mode <= 2, "pool_bwd: bad arguments");
  HGB_REQUIRE(mode != HGB_POOL_MAX || argmax, "pool_bwd: max pooling needs the argmax buffer");
  if (g == 0) return HGB_OK;
  pool_bwd_kernel<<<hgb_grid_for(g, 8), 256, 0, (cudaStream_t)stream>>>(gout, graph_ptr, argmax, n, g, c, mode, gx);
  HGB_LAUNCH_CHECK("pool_bwd");
  return HGB_OK;
}

// ---- segmented arg-min / arg-max (PNA aggregators, hydragnn/models/PNAEqStack.py:396-400) ---------------------
// For every (node, channel): the EDGE id holding the minimum / maximum over the node's CSR segment (first one wins
// on ties, -1 for an empty segment).  Values and gradients are then plain gathers at those ids.
__global__ void segment_argminmax_kernel(const float* __restrict__ m, const int32_t* __restrict__ rowptr,
                                         const int32_t* __restrict__ perm, int n, int c, int lanes, int64_t* __restrict__ amin,
                                         int64_t* __restrict__ amax) {
  const int gpb = blockDim.x / lanes;
  const int sub = threadIdx.x % lanes;
  for (int row = blockIdx.x * gpb + threadIdx.x / lanes; row < n; row += gridDim.x * gpb) {
    const int lo = rowptr[row], hi = rowptr[row + 1];
    for (int ch = sub; ch < c; ch += lanes) {
      float vmin = INFINITY, vmax = -INFINITY;
      int64_t imin = -1, imax = -1;
      for (int p = lo; p < hi; ++p) {
        const int e = perm ? perm[p] : p;
        const float v = m[(int64_t)e * c + ch];
        if (imin < 0 || v < vmin) { vmin = v; imin = e; }
        if (imax < 0 || v > vmax) { vmax = v; imax = e; }
      }
      amin[(int64_t)row * c + ch] = imin;
      amax[(int64_t)row * c + ch] = imax;
    }
  }
}

extern "C" int hgb_segment_argminmax(const float* m, const int32_t* rowptr, const int32_t* perm, int32_t n, int32_t c,
                                     int64_t* argmin, int64_t* argmax, hgb_stream_t stream) {
  HGB_REQUIRE(n >= 0 && c > 0 && rowptr && argmin && argmax, "segment_argminmax: bad arguments");
  if (n == 0) return HGB_OK;
  const int lanes = group_lanes(c);
  segment_argminmax_kernel<<<hgb_grid_for(n, 256 / lanes), 256, 0, (cudaStream_t)stream>>>(m, rowptr, perm, n, c, lanes, argmin, argmax);
  HGB_LAUNCH_CHECK("segment_argminmax");
  return HGB_OK;
}

// ---- PNA aggregation: mean | min | max | std of every CSR segment in ONE pass (PNAEqStack.py:396-400; PyG 2.6.1
// MeanAggregation / MinAggregation / MaxAggregation / StdAggregation).  out [n, 4c]; amin / amax [n, c] hold the EDGE id
// of the first minimum / maximum (-1 for an empty segment) for the backward.  std = sqrt(clamp(E[x^2] - E[x]^2, 1e-5)),
// reported as 0 where it equals sqrt(1e-5) (PyG masks the clamped entries).
#define PNA_EPS 1e-5f
__global__ void pna_aggregate_fwd_kernel(const float* __restrict__ m, const int32_t* __restrict__ rowptr,
                                         const int32_t* __restrict__ perm, int n, int c, int lanes, float* __restrict__ out,
                                         int32_t* __restrict__ amin, int32_t* __restrict__ amax) {
  const int gpb = blockDim.x / lanes;
  const int sub = threadIdx.x % lanes;
  for (int row = blockIdx.x * gpb + threadIdx.x / lanes; row < n; row += gridDim.x * gpb) {
    const int lo = rowptr[row], hi = rowptr[row + 1];
    const float inv = 1.f / (float)max(hi - lo, 1);
    for (int ch = sub; ch < c; ch += lanes) {
      float s1 = 0.f, s2 = 0.f, vmin = 0.f, vmax = 0.f;
      int imin = -1, imax = -1;
      for (int p = lo; p < hi; ++p) {
        const int e = perm ? perm[p] : p;
        const float v = __ldg(m + (int64_t)e * c + ch);
        s1 += v;
        s2 = fmaf(v, v, s2);
        if (imin < 0 || v < vmin) { vmin = v; imin = e; }
        if (imax < 0 || v > vmax) { vmax = v; imax = e; }
      }
      const float mean = s1 * inv;
      float sd = sqrtf(fmaxf(s2 * inv - mean * mean, PNA_EPS));
      if (sd <= sqrtf(PNA_EPS)) sd = 0.f;
      float* o = out + (int64_t)row * 4 * c;
      o[ch] = mean;
      o[c + ch] = vmin;
      o[2 * c + ch] = vmax;
      o[3 * c + ch] = sd;
      amin[(int64_t)row * c + ch] = imin;
      amax[(int64_t)row * c + ch] = imax;
    }
  }
}

// g_m[e, ch] = g_mean/cnt + [e == amin] g_min + [e == amax] g_max + g_std (m - mean) / (cnt std)
__global__ void pna_aggregate_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ m, const float* __restrict__ out,
                                         const int32_t* __restrict__ idx, const int32_t* __restrict__ rowptr,
                                         const int32_t* __restrict__ amin, const int32_t* __restrict__ amax, int64_t total, int c,
                                         float* __restrict__ gm) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(t / c);
    const int ch = (int)(t - (int64_t)e * c);
    const int i = idx[e];
    const float inv = 1.f / (float)max(rowptr[i + 1] - rowptr[i], 1);
    const float* g = gout + (int64_t)i * 4 * c;
    const float* o = out + (int64_t)i * 4 * c;
    float acc = g[ch] * inv;
    if (amin[(int64_t)i * c + ch] == e) acc += g[c + ch];
    if (amax[(int64_t)i * c + ch] == e) acc += g[2 * c + ch];
    const float sd = o[3 * c + ch];
    if (sd > 0.f) acc = fmaf(g[3 * c + ch] * inv / sd, m[t] - o[ch], acc);
    gm[t] = acc;
  }
}

extern "C" int hgb_pna_aggregate_fwd(const float* m, const int32_t* rowptr, const int32_t* perm, int32_t n, int32_t c, float* out,
                                     int32_t* argmin, int32_t* argmax, hgb_stream_t stream) {
  HGB_REQUIRE(n >= 0 && c > 0 && rowptr && out && argmin && argmax, "pna_aggregate_fwd: bad arguments");
  if (n == 0) return HGB_OK;
  const int lanes = group_lanes(c);
  pna_aggregate_fwd_kernel<<<hgb_grid_for(n, 256 / lanes), 256, 0, (cudaStream_t)stream>>>(m, rowptr, perm, n, c, lanes, out, argmin, argmax);
  HGB_LAUNCH_CHECK("pna_aggregate_fwd");
  return HGB_OK;
}

extern "C" int hgb_pna_aggregate_bwd(const float* g_out, const float* m, const float* out, const int32_t* idx, const int32_t* rowptr,
                                     const int32_t* argmin, const int32_t* argmax, int64_t e, int32_t c, float* g_m, hgb_stream_t stream) {
  HGB_REQUIRE(e >= 0 && c > 0 && g_out && out && idx && rowptr && argmin && argmax && g_m, "pna_aggregate_bwd: bad arguments");
  if (e == 0) return HGB_OK;
  HGB_REQUIRE(m, "pna_aggregate_bwd: null messages");
  pna_aggregate_bwd_kernel<<<hgb_grid_for(e * c, 256), 256, 0, (cudaStream_t)stream>>>(g_out, m, out, idx, rowptr, argmin, argmax, e * c, c, g_m);
  HGB_LAUNCH_CHECK("pna_aggregate_bwd");
  return HGB_OK;
}
