// libhgb.so -- PaiNN message (fused gather -> filter -> gate -> segmented sum) and update glue.
//
// Message forward:  one warp per aggregation node i, lanes own channels (CPL per lane), the warp walks
// the CSR segment of edge[:,0] == i.  Per edge it reads 9+R scalars (broadcast), the 3F-wide phi row and
// the 3F-wide v row of the source node -- 128 B coalesced per warp load -- and keeps the 4F partial sums
// in registers.  Nothing per-edge is written; no atomics; summation order = ascending edge id.
// Message backward is the mirror image over the CSR of edge[:,1] (the gather side): the warp that owns
// source node j reads the incoming gradients of every node it sent a message to.
// Filter weights (3F x R) sit in shared memory with an odd row stride (bank-conflict free).
#include "hgb_common.cuh"

#define RMAX 8
#define WPB 8  // warps per block

// GROUP = lanes that share one node (32 for F >= 32; for narrow layers -- the reference's first layer runs at
// F = input_dim, often 1 -- a warp serves 32/GROUP nodes at once instead of idling 31 lanes)
template <int G>
__device__ __forceinline__ float hgb_group_sum(float v, unsigned mask) {   // groups of one warp may diverge: shuffle within the group only
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}

template <int CPL, bool HAS_EF, int GROUP>
__global__ void __launch_bounds__(WPB * 32)
painn_message_fwd_kernel(const float* __restrict__ phi, const float* __restrict__ s, const float* __restrict__ v,
                         const int32_t* __restrict__ rowptr, const int32_t* __restrict__ perm,
                         const int32_t* __restrict__ src, const float* __restrict__ dir, const float* __restrict__ rbfc,
                         const float* __restrict__ fc, const float* __restrict__ wf, const float* __restrict__ bf,
                         const float* __restrict__ efilt, int n, int f, int r, float* __restrict__ s_out,
                         float* __restrict__ v_out) {
  extern __shared__ float sm[];
  const int rs = r | 1;
  float* wfs = sm;               // [3f][rs]
  float* bfs = sm + 3 * f * rs;  // [3f]
  for (int t = threadIdx.x; t < 3 * f * r; t += blockDim.x) wfs[(t / r) * rs + (t % r)] = wf[t];
  for (int t = threadIdx.x; t < 3 * f; t += blockDim.x) bfs[t] = bf[t];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  constexpr int NPW = 32 / GROUP;
  const int sub = lane % GROUP, nsub = lane / GROUP;
  const int cbase = blockIdx.y * GROUP * CPL;
  int ch[CPL];
  bool ok[CPL];
#pragma unroll
  for (int t = 0; t < CPL; ++t) { ch[t] = cbase + sub + GROUP * t; ok[t] = ch[t] < f; if (!ok[t]) ch[t] = 0; }
  const int f3 = 3 * f;
  for (int i = (blockIdx.x * WPB + (threadIdx.x >> 5)) * NPW + nsub; i < n; i += gridDim.x * WPB * NPW) {
    const int lo = rowptr[i], hi = rowptr[i + 1];
    float as[CPL], av[CPL][3];
#pragma unroll
    for (int t = 0; t < CPL; ++t) { as[t] = 0.f; av[t][0] = av[t][1] = av[t][2] = 0.f; }
    for (int p = lo; p < hi; ++p) {
      const int e = perm ? perm[p] : p;
      const int j = src[e];
      const float fce = fc[e];
      const float d0 = dir[3 * (int64_t)e], d1 = dir[3 * (int64_t)e + 1], d2 = dir[3 * (int64_t)e + 2];
      float rb[RMAX];
#pragma unroll
      for (int q = 0; q < RMAX; ++q) rb[q] = q < r ? rbfc[(int64_t)e * r + q] : 0.f;
      const float* ph = phi + (int64_t)j * f3;
      const float* vj = v + (int64_t)j * f3;
#pragma unroll
      for (int t = 0; t < CPL; ++t) {
        const int c = ch[t];
        float w0 = bfs[c] * fce, w1 = bfs[f + c] * fce, w2 = bfs[2 * f + c] * fce;
#pragma unroll
        for (int q = 0; q < RMAX; ++q)
          if (q < r) {
            w0 = fmaf(wfs[c * rs + q], rb[q], w0);
            w1 = fmaf(wfs[(f + c) * rs + q], rb[q], w1);
            w2 = fmaf(wfs[(2 * f + c) * rs + q], rb[q], w2);
          }
        if (HAS_EF) {
          const float* ef = efilt + (int64_t)e * f3;
          w0 *= ef[c]; w1 *= ef[f + c]; w2 *= ef[2 * f + c];
        }
        const float gv = w0 * __ldg(ph + c), ge = w1 * __ldg(ph + f + c), ms = w2 * __ldg(ph + 2 * f + c);
        as[t] += ms;
        av[t][0] += __ldg(vj + c) * gv + ge * d0;
        av[t][1] += __ldg(vj + f + c) * gv + ge * d1;
        av[t][2] += __ldg(vj + 2 * f + c) * gv + ge * d2;
      }
    }
#pragma unroll
    for (int t = 0; t < CPL; ++t)
      if (ok[t]) {
        const int c = ch[t];
        s_out[(int64_t)i * f + c] = s[(int64_t)i * f + c] + as[t];
#pragma unroll
        for (int k = 0; k < 3; ++k) v_out[(int64_t)i * f3 + k * f + c] = v[(int64_t)i * f3 + k * f + c] + av[t][k];
      }
  }
}

extern "C" int hgb_painn_message_fwd(const float* phi, const float* s, const float* v, const int32_t* rowptr,
                                     const int32_t* perm, const int32_t* src, const float* dir, const float* rbfc,
                                     const float* fc, const float* wf, const float* bf, const float* efilt, int32_t n,
                                     int32_t f, int32_t r, float* s_out, float* v_out, hgb_stream_t stream) {
  HGB_REQUIRE(n >= 0 && f > 0 && r > 0 && r <= RMAX, "painn_message_fwd: need 0 < num_radial <= %d (got %d)", RMAX, r);
  HGB_REQUIRE(phi && s && v && rowptr && src && dir && rbfc && fc && wf && bf && s_out && v_out, "painn_message_fwd: null pointer");
  if (n == 0) return HGB_OK;
  const size_t smem = (size_t)(3 * f * ((r | 1) + 1)) * sizeof(float);
  HGB_REQUIRE(smem <= 48 * 1024, "painn_message_fwd: hidden_dim %d too large for the filter staging", f);
  const int cpl = f <= 32 ? 1 : 2;
  int group = 32;
  if (f < 32) { group = 1; while (group < f) group <<= 1; }
  dim3 grid(hgb_grid_for(n, WPB * (32 / group), HGB_NUM_SMS * 8), (f + group * cpl - 1) / (group * cpl));
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(C, E, G) painn_message_fwd_kernel<C, E, G><<<grid, WPB * 32, smem, st>>>(phi, s, v, rowptr, perm, src, dir, rbfc, fc, wf, bf, efilt, n, f, r, s_out, v_out)
#define LAUNCH_G(E)                                                  \
  switch (group) {                                                   \
    case 1: LAUNCH(1, E, 1); break;                                  \
    case 2: LAUNCH(1, E, 2); break;                                  \
    case 4: LAUNCH(1, E, 4); break;                                  \
    case 8: LAUNCH(1, E, 8); break;                                  \
    case 16: LAUNCH(1, E, 16); break;                                \
    default: if (cpl == 1) LAUNCH(1, E, 32); else LAUNCH(2, E, 32);  \
  }
  if (efilt) { LAUNCH_G(true) } else { LAUNCH_G(false) }
#undef LAUNCH_G
#undef LAUNCH
  HGB_LAUNCH_CHECK("painn_message_fwd");
  return HGB_OK;
}

// ---- backward ----------------------------------------------------------------------------------------
// workspace layout: part[gridDim.x][3f][r+1]  (column r holds the bias gradient)
template <int CPL, bool HAS_EF, bool NEED_EDGE, int GROUP>
__global__ void __launch_bounds__(WPB * 32)
painn_message_bwd_kernel(const float* __restrict__ gs_out, const float* __restrict__ gv_out, const float* __restrict__ phi,
                         const float* __restrict__ v, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ perm,
                         const int32_t* __restrict__ agg, const float* __restrict__ dir, const float* __restrict__ rbfc,
                         const float* __restrict__ fc, const float* __restrict__ wf, const float* __restrict__ bf,
                         const float* __restrict__ efilt, int n, int f, int r, float* __restrict__ gphi, float* __restrict__ gv,
                         float* __restrict__ part, float* __restrict__ g_dir, float* __restrict__ g_rbfc,
                         float* __restrict__ g_fc, float* __restrict__ g_efilt, int multi_cb) {
  extern __shared__ float sm[];
  const int rs = r | 1;
  float* wfs = sm;
  float* bfs = sm + 3 * f * rs;
  float* red = bfs + 3 * f;  // [WPB][32]
  for (int t = threadIdx.x; t < 3 * f * r; t += blockDim.x) wfs[(t / r) * rs + (t % r)] = wf[t];
  for (int t = threadIdx.x; t < 3 * f; t += blockDim.x) bfs[t] = bf[t];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NPW = 32 / GROUP;
  const int sub = lane % GROUP, nsub = lane / GROUP;
  const int cbase = blockIdx.y * GROUP * CPL;
  int ch[CPL];
  bool ok[CPL];
#pragma unroll
  for (int t = 0; t < CPL; ++t) { ch[t] = cbase + sub + GROUP * t; ok[t] = ch[t] < f; if (!ok[t]) ch[t] = 0; }
  const int f3 = 3 * f;
  const unsigned gmask = GROUP == 32 ? 0xffffffffu : (((1u << GROUP) - 1u) << (nsub * GROUP));
  float gw[CPL][3][RMAX + 1];
#pragma unroll
  for (int t = 0; t < CPL; ++t)
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int q = 0; q <= RMAX; ++q) gw[t][a][q] = 0.f;

  for (int j = (blockIdx.x * WPB + warp) * NPW + nsub; j < n; j += gridDim.x * WPB * NPW) {
    const int lo = rowptr[j], hi = rowptr[j + 1];
    float ph[CPL][3], vj[CPL][3], aphi[CPL][3], agv[CPL][3];
#pragma unroll
    for (int t = 0; t < CPL; ++t)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        ph[t][k] = phi[(int64_t)j * f3 + k * f + ch[t]];
        vj[t][k] = v[(int64_t)j * f3 + k * f + ch[t]];
        aphi[t][k] = 0.f;
        agv[t][k] = 0.f;
      }
    for (int p = lo; p < hi; ++p) {
      const int e = perm ? perm[p] : p;
      const int i = agg[e];
      const float fce = fc[e];
      const float d0 = dir[3 * (int64_t)e], d1 = dir[3 * (int64_t)e + 1], d2 = dir[3 * (int64_t)e + 2];
      float rb[RMAX];
#pragma unroll
      for (int q = 0; q < RMAX; ++q) rb[q] = q < r ? rbfc[(int64_t)e * r + q] : 0.f;
      float e_rb[RMAX], e_fc = 0.f, e_d0 = 0.f, e_d1 = 0.f, e_d2 = 0.f;  // per-edge gradients (NEED_EDGE)
#pragma unroll
      for (int q = 0; q < RMAX; ++q) e_rb[q] = 0.f;
#pragma unroll
      for (int t = 0; t < CPL; ++t) {
        const int c = ch[t];
        float w[3];
        w[0] = bfs[c] * fce; w[1] = bfs[f + c] * fce; w[2] = bfs[2 * f + c] * fce;
#pragma unroll
        for (int q = 0; q < RMAX; ++q)
          if (q < r) {
            w[0] = fmaf(wfs[c * rs + q], rb[q], w[0]);
            w[1] = fmaf(wfs[(f + c) * rs + q], rb[q], w[1]);
            w[2] = fmaf(wfs[(2 * f + c) * rs + q], rb[q], w[2]);
          }
        float ef[3] = {1.f, 1.f, 1.f};
        if (HAS_EF) {
#pragma unroll
          for (int a = 0; a < 3; ++a) ef[a] = efilt[(int64_t)e * f3 + a * f + c];
        }
        const float gsi = ok[t] ? __ldg(gs_out + (int64_t)i * f + c) : 0.f;
        float gvi[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) gvi[k] = ok[t] ? __ldg(gv_out + (int64_t)i * f3 + k * f + c) : 0.f;
        // gradients w.r.t. the three gate values f = W * ef * phi
        float gg[3];
        gg[0] = gvi[0] * vj[t][0] + gvi[1] * vj[t][1] + gvi[2] * vj[t][2];
        gg[1] = gvi[0] * d0 + gvi[1] * d1 + gvi[2] * d2;
        gg[2] = gsi;
        const float gate_v = w[0] * ef[0] * ph[t][0];
#pragma unroll
        for (int k = 0; k < 3; ++k) agv[t][k] += gvi[k] * gate_v;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          aphi[t][a] += gg[a] * w[a] * ef[a];
          const float gwe = gg[a] * ph[t][a];  // gradient w.r.t. (W * ef)
          if (HAS_EF && ok[t]) g_efilt[(int64_t)e * f3 + a * f + c] = gwe * w[a];
          const float gW = gwe * ef[a];        // gradient w.r.t. the raw filter W[a, c]
#pragma unroll
          for (int q = 0; q < RMAX; ++q)
            if (q < r) gw[t][a][q] = fmaf(gW, rb[q], gw[t][a][q]);
          gw[t][a][RMAX] = fmaf(gW, fce, gw[t][a][RMAX]);
          if (NEED_EDGE) {
#pragma unroll
            for (int q = 0; q < RMAX; ++q)
              if (q < r) e_rb[q] = fmaf(gW, wfs[(a * f + c) * rs + q], e_rb[q]);
            e_fc = fmaf(gW, bfs[a * f + c], e_fc);
          }
        }
        if (NEED_EDGE) {
          const float ge = w[1] * ef[1] * ph[t][1];
          e_d0 = fmaf(gvi[0], ge, e_d0); e_d1 = fmaf(gvi[1], ge, e_d1); e_d2 = fmaf(gvi[2], ge, e_d2);
        }
      }
      if (NEED_EDGE) {
#pragma unroll
        for (int q = 0; q < RMAX; ++q)
          if (q < r) e_rb[q] = hgb_group_sum<GROUP>(e_rb[q], gmask);
        e_fc = hgb_group_sum<GROUP>(e_fc, gmask); e_d0 = hgb_group_sum<GROUP>(e_d0, gmask); e_d1 = hgb_group_sum<GROUP>(e_d1, gmask); e_d2 = hgb_group_sum<GROUP>(e_d2, gmask);
        if (sub == 0) {
          if (multi_cb) {  // several channel blocks contribute to the same edge
#pragma unroll
            for (int q = 0; q < RMAX; ++q)
              if (q < r) atomicAdd(g_rbfc + (int64_t)e * r + q, e_rb[q]);
            atomicAdd(g_fc + e, e_fc);
            atomicAdd(g_dir + 3 * (int64_t)e, e_d0); atomicAdd(g_dir + 3 * (int64_t)e + 1, e_d1); atomicAdd(g_dir + 3 * (int64_t)e + 2, e_d2);
          } else {
#pragma unroll
            for (int q = 0; q < RMAX; ++q)
              if (q < r) g_rbfc[(int64_t)e * r + q] = e_rb[q];
            g_fc[e] = e_fc;
            g_dir[3 * (int64_t)e] = e_d0; g_dir[3 * (int64_t)e + 1] = e_d1; g_dir[3 * (int64_t)e + 2] = e_d2;
          }
        }
      }
    }
#pragma unroll
    for (int t = 0; t < CPL; ++t)
      if (ok[t]) {
        const int c = ch[t];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          gphi[(int64_t)j * f3 + k * f + c] = aphi[t][k];
          gv[(int64_t)j * f3 + k * f + c] = gv_out[(int64_t)j * f3 + k * f + c] + agv[t][k];
        }
      }
  }
  // block-level reduction of the filter-weight gradients -> part[blockIdx.x]
  float* mypart = part + (int64_t)blockIdx.x * f3 * (r + 1);
#pragma unroll
  for (int t = 0; t < CPL; ++t)
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int q = 0; q <= RMAX; ++q) {
        if (q < r || q == RMAX) {
          __syncthreads();
          red[warp * 32 + lane] = gw[t][a][q];
          __syncthreads();
          if (warp == 0 && lane < GROUP) {     // lanes with the same `sub` own the same channel
            float acc = 0.f;
#pragma unroll
            for (int w8 = 0; w8 < WPB; ++w8)
#pragma unroll
              for (int g = 0; g < NPW; ++g) acc += red[w8 * 32 + lane + GROUP * g];
            if (ok[t]) mypart[(a * f + ch[t]) * (r + 1) + (q == RMAX ? r : q)] = acc;
          }
        }
      }
}

__global__ void painn_wgrad_reduce_kernel(const float* __restrict__ part, int nblocks, int f3, int r,
                                          float* __restrict__ gwf, float* __restrict__ gbf) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= f3 * (r + 1)) return;
  float acc = 0.f;
  for (int b = 0; b < nblocks; ++b) acc += part[(int64_t)b * f3 * (r + 1) + t];
  const int row = t / (r + 1), q = t % (r + 1);
  if (q == r) gbf[row] = acc; else gwf[row * r + q] = acc;
}

static int painn_group(int f) { int g = 32; if (f < 32) { g = 1; while (g < f) g <<= 1; } return g; }
static int painn_bwd_grid(int n, int f) { return hgb_grid_for(n, WPB * (32 / painn_group(f)), HGB_NUM_SMS * 4); }

extern "C" int64_t hgb_painn_message_bwd_workspace_bytes(int32_t n, int32_t f, int32_t r) {
  return (int64_t)painn_bwd_grid(n, f) * 3 * f * (r + 1) * 4;
}

extern "C" int hgb_painn_message_bwd(const float* gs_out, const float* gv_out, const float* phi, const float* v,
                                     const int32_t* rowptr_src, const int32_t* perm_src, const int32_t* agg, const float* dir,
                                     const float* rbfc, const float* fc, const float* wf, const float* bf, const float* efilt,
                                     int32_t n, int32_t f, int32_t r, float* gphi, float* gv, float* gwf, float* gbf,
                                     float* g_dir, float* g_rbfc, float* g_fc, float* g_efilt, void* workspace,
                                     int64_t workspace_bytes, hgb_stream_t stream) {
  HGB_REQUIRE(n >= 0 && f > 0 && r > 0 && r <= RMAX, "painn_message_bwd: need 0 < num_radial <= %d (got %d)", RMAX, r);
  HGB_REQUIRE(gs_out && gv_out && phi && v && rowptr_src && agg && dir && rbfc && fc && wf && bf && gphi && gv && gwf && gbf && workspace,
              "painn_message_bwd: null pointer");
  const bool need_edge = g_dir != nullptr;
  HGB_REQUIRE((g_rbfc != nullptr) == need_edge && (g_fc != nullptr) == need_edge, "painn_message_bwd: g_dir/g_rbfc/g_fc must come together");
  HGB_REQUIRE((efilt != nullptr) == (g_efilt != nullptr), "painn_message_bwd: g_efilt iff efilt");
  HGB_REQUIRE(workspace_bytes >= hgb_painn_message_bwd_workspace_bytes(n, f, r), "painn_message_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int f3 = 3 * f;
  if (n == 0) {
    cudaMemsetAsync(gwf, 0, (size_t)f3 * r * 4, st);
    cudaMemsetAsync(gbf, 0, (size_t)f3 * 4, st);
    return HGB_OK;
  }
  const size_t smem = (size_t)(f3 * ((r | 1) + 1) + WPB * 32) * sizeof(float);
  HGB_REQUIRE(smem <= 48 * 1024, "painn_message_bwd: hidden_dim %d too large for the filter staging", f);
  const int cpl = f <= 32 ? 1 : 2;
  const int group = painn_group(f);
  const int ncb = (f + group * cpl - 1) / (group * cpl);
  dim3 grid(painn_bwd_grid(n, f), ncb);
  float* part = (float*)workspace;
#define LAUNCH(C, E, G, W) painn_message_bwd_kernel<C, E, G, W><<<grid, WPB * 32, smem, st>>>(gs_out, gv_out, phi, v, rowptr_src, perm_src, agg, dir, rbfc, fc, wf, bf, efilt, n, f, r, gphi, gv, part, g_dir, g_rbfc, g_fc, g_efilt, ncb > 1)
#define LAUNCH_G(E, G)                                                     \
  switch (group) {                                                         \
    case 1: LAUNCH(1, E, G, 1); break;                                     \
    case 2: LAUNCH(1, E, G, 2); break;                                     \
    case 4: LAUNCH(1, E, G, 4); break;                                     \
    case 8: LAUNCH(1, E, G, 8); break;                                     \
    case 16: LAUNCH(1, E, G, 16); break;                                   \
    default: if (cpl == 1) LAUNCH(1, E, G, 32); else LAUNCH(2, E, G, 32);  \
  }
  if (efilt) { if (need_edge) { LAUNCH_G(true, true) } else { LAUNCH_G(true, false) } }
  else { if (need_edge) { LAUNCH_G(false, true) } else { LAUNCH_G(false, false) } }
#undef LAUNCH_G
#undef LAUNCH
  HGB_LAUNCH_CHECK("painn_message_bwd");
  painn_wgrad_reduce_kernel<<<(f3 * (r + 1) + 127) / 128, 128, 0, st>>>(part, grid.x, f3, r, gwf, gbf);
  HGB_LAUNCH_CHECK("painn_wgrad_reduce");
  return HGB_OK;
}

// ---- update block glue --------------------------------------------------------------------------------
__global__ void painn_update_pre_fwd_kernel(const float* __restrict__ vv, const float* __restrict__ s, int64_t nf, int f,
                                            float* __restrict__ mlp_in) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nf; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / f;
    const int c = (int)(t % f);
    const float a = vv[i * 3 * f + c], b = vv[i * 3 * f + f + c], d = vv[i * 3 * f + 2 * f + c];
    mlp_in[i * 2 * f + c] = sqrtf(a * a + b * b + d * d);
    mlp_in[i * 2 * f + f + c] = s[t];
  }
}

extern "C" int hgb_painn_update_pre_fwd(const float* vv, const float* s, int32_t n, int32_t f, float* mlp_in, hgb_stream_t stream) {
  HGB_REQUIRE(n >= 0 && f > 0 && vv && s && mlp_in, "painn_update_pre_fwd: bad arguments");
  if (n == 0) return HGB_OK;
  const int64_t nf = (int64_t)n * f;
  painn_update_pre_fwd_kernel<<<hgb_grid_for(nf, 256), 256, 0, (cudaStream_t)stream>>>(vv, s, nf, f, mlp_in);
  HGB_LAUNCH_CHECK("painn_update_pre_fwd");
  return HGB_OK;
}

__global__ void painn_update_post_fwd_kernel(const float* __restrict__ a, const float* __restrict__ uv,
                                             const float* __restrict__ vv, const float* __restrict__ s,
                                             const float* __restrict__ v, int64_t nf, int f, int last, float* __restrict__ s_out,
                                             float* __restrict__ v_out) {
  const int na = last ? 2 : 3;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nf; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / f;
    const int c = (int)(t % f);
    const float* ai = a + i * na * f;
    const float a_sv = ai[(na - 2) * f + c], a_ss = ai[(na - 1) * f + c];
    float inner = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) inner += uv[i * 3 * f + k * f + c] * vv[i * 3 * f + k * f + c];
    s_out[t] = s[t] + a_sv * inner + a_ss;
    if (!last) {
      const float a_vv = ai[c];
#pragma unroll
      for (int k = 0; k < 3; ++k) v_out[i * 3 * f + k * f + c] = v[i * 3 * f + k * f + c] + a_vv * uv[i * 3 * f + k * f + c];
    }
  }
}

extern "C" int hgb_painn_update_post_fwd(const float* a, const float* uv, const float* vv, const float* s, const float* v,
                                         int32_t n, int32_t f, int32_t last, float* s_out, float* v_out, hgb_stream_t stream) {
  HGB_REQUIRE(n >= 0 && f > 0 && a && uv && vv && s && s_out && (last || (v && v_out)), "painn_update_post_fwd: bad arguments");
  if (n == 0) return HGB_OK;
  const int64_t nf = (int64_t)n * f;
  painn_update_post_fwd_kernel<<<hgb_grid_for(nf, 256), 256, 0, (cudaStream_t)stream>>>(a, uv, vv, s, v, nf, f, last, s_out, v_out);
  HGB_LAUNCH_CHECK("painn_update_post_fwd");
  return HGB_OK;
}

// ga = gradient w.r.t. the update_mlp output a (needed first: it feeds the MLP backward)
__global__ void painn_update_post_bwd_a_kernel(const float* __restrict__ gs_out, const float* __restrict__ gv_out,
                                               const float* __restrict__ uv, const float* __restrict__ vv, int64_t nf, int f,
                                               int last, float* __restrict__ ga) {
  const int na = last ? 2 : 3;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nf; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / f;
    const int c = (int)(t % f);
    float inner = 0.f, gdot = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float u = uv[i * 3 * f + k * f + c];
      inner += u * vv[i * 3 * f + k * f + c];
      if (!last) gdot += gv_out[i * 3 * f + k * f + c] * u;
    }
    float* gi = ga + i * na * f;
    const float g = gs_out[t];
    if (!last) gi[c] = gdot;
    gi[(na - 2) * f + c] = g * inner;
    gi[(na - 1) * f + c] = g;
  }
}

extern "C" int hgb_painn_update_post_bwd_a(const float* gs_out, const float* gv_out, const float* uv, const float* vv, int32_t n,
                                           int32_t f, int32_t last, float* ga, hgb_stream_t stream) {
  HGB_REQUIRE(n >= 0 && f > 0 && gs_out && uv && vv && ga && (last || gv_out), "painn_update_post_bwd_a: bad arguments");
  if (n == 0) return HGB_OK;
  const int64_t nf = (int64_t)n * f;
  painn_update_post_bwd_a_kernel<<<hgb_grid_for(nf, 256), 256, 0, (cudaStream_t)stream>>>(gs_out, gv_out, uv, vv, nf, f, last, ga);
  HGB_LAUNCH_CHECK("painn_update_post_bwd_a");
  return HGB_OK;
}

// everything else of the update backward: guv, gvv (inputs of the U / V linear backward), gs, gv
__global__ void painn_update_bwd_kernel(const float* __restrict__ gs_out, const float* __restrict__ gv_out,
                                        const float* __restrict__ g_mlp_in, const float* __restrict__ a,
                                        const float* __restrict__ uv, const float* __restrict__ vv,
                                        const float* __restrict__ mlp_in, int64_t nf, int f, int last, float* __restrict__ guv,
                                        float* __restrict__ gvv, float* __restrict__ gs, float* __restrict__ gv) {
  const int na = last ? 2 : 3;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nf; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / f;
    const int c = (int)(t % f);
    const float* ai = a + i * na * f;
    const float a_sv = ai[(na - 2) * f + c];
    const float a_vv = last ? 0.f : ai[c];
    const float g = gs_out[t];
    const float nrm = mlp_in[i * 2 * f + c];
    const float gn = g_mlp_in[i * 2 * f + c];
    const float gn_over = nrm > 0.f ? gn / nrm : 0.f;   // d|vv|/dvv = vv/|vv| (0 at the origin, as torch)
    gs[t] = g + g_mlp_in[i * 2 * f + f + c];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int64_t o = i * 3 * f + k * f + c;
      const float u = uv[o], w = vv[o];
      const float gvo = last ? 0.f : gv_out[o];
      guv[o] = gvo * a_vv + g * a_sv * w;
      gvv[o] = g * a_sv * u + gn_over * w;
      if (gv) gv[o] = gvo;   // direct path v -> v_out (the U / V linear backward adds the rest)
    }
  }
}

extern "C" int hgb_painn_update_bwd(const float* gs_out, const float* gv_out, const float* g_mlp_in, const float* a,
                                    const float* uv, const float* vv, const float* mlp_in, int32_t n, int32_t f, int32_t last,
                                    float* guv, float* gvv, float* gs, float* gv, hgb_stream_t stream) {
  HGB_REQUIRE(n >= 0 && f > 0 && gs_out && g_mlp_in && a && uv && vv && mlp_in && guv && gvv && gs && (last || gv_out),
              "painn_update_bwd: bad arguments");
  if (n == 0) return HGB_OK;
  const int64_t nf = (int64_t)n * f;
  painn_update_bwd_kernel<<<hgb_grid_for(nf, 256), 256, 0, (cudaStream_t)stream>>>(gs_out, gv_out, g_mlp_in, a, uv, vv, mlp_in,
                                                                                 nf, f, last, guv, gvv, gs, gv);
  HGB_LAUNCH_CHECK("painn_update_bwd");
  return HGB_OK;
}
