// libhgb.so -- closed (any-order differentiable) primitives for the MACE force-training path and for the shapes the fused
// first-order kernels of hgb_mace.cu do not cover (correlation 3, channel counts that are not a multiple of 32).
//
// Round 1 composed these from torch.einsum; VERDICT r1 asked for hand-written primitives that are closed under differentiation,
// like GatherRows <-> SegmentSum.  Two families:
//
// (1) per-edge coupling of one tensor-product path (blocks.py:386-397, o3.TensorProduct "uvu" with per-edge weights):
//       tp_out   o[e, k, f] = w[e, f] * sum_{i, j} C[i, j, k] a[e, i, f] y[e, j]
//       tp_y     o[e, j]    = sum_f w[e, f] sum_{i, k} C[i, j, k] a[e, i, f] g[e, k, f]
//       tp_w     o[e, f]    = sum_{i, j, k} C[i, j, k] a[e, i, f] y[e, j] g[e, k, f]
//     C = real coupling tensor of the path (<= 7 x 7 x 7).  The three forms are each other's derivatives (the derivative of
//     tp_out w.r.t. a is tp_out again with C permuted), so any order of differentiation stays inside these kernels.
//
// (2) per-node, per-channel contraction steps of the symmetric contraction (symmetric_contraction.py:217-239):
//       chan_cl  o[b, c, p]    = sum_i t[b, c, p, i] x[b, i, c]        ("contract last")
//       chan_ou  o[b, c, p, i] = g[b, c, p] x[b, i, c]                 ("outer")
//       chan_rp  o[b, i, c]    = sum_p g[b, c, p] t[b, c, p, i]        ("reduce p")
//     again mutually adjoint.  The U-matrix x weight products in front of them are plain MatMuls (closed already).
//
// All are bandwidth-trivial elementwise-style SIMT kernels: one thread per output element (tp_y: one warp per edge).
#include "hgb_common.cuh"

namespace {

constexpr int CMAX = 7 * 7 * 7;

__global__ void tp_out_kernel(const float* __restrict__ a, const float* __restrict__ y, const float* __restrict__ w,
                              const float* __restrict__ cg, int64_t e, int f, int ni, int nj, int nk, float* __restrict__ out) {
  __shared__ float sc[CMAX];
  for (int q = threadIdx.x; q < ni * nj * nk; q += blockDim.x) sc[q] = cg[q];
  __syncthreads();
  const int64_t total = e * f;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ed = idx / f;
    const int c = (int)(idx - ed * f);
    float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i = 0; i < ni; ++i) {
      const float av = a[(ed * ni + i) * f + c];
      for (int j = 0; j < nj; ++j) {
        const float ay = av * y[ed * nj + j];
        const float* cr = sc + (i * nj + j) * nk;
        for (int k = 0; k < nk; ++k) acc[k] = fmaf(cr[k], ay, acc[k]);
      }
    }
    const float wv = w[ed * f + c];
    for (int k = 0; k < nk; ++k) out[(ed * nk + k) * f + c] = wv * acc[k];
  }
}

// one warp per edge, lanes over channels
__global__ void tp_y_kernel(const float* __restrict__ a, const float* __restrict__ g, const float* __restrict__ w,
                            const float* __restrict__ cg, int64_t e, int f, int ni, int nj, int nk, float* __restrict__ out) {
  __shared__ float sc[CMAX];
  for (int q = threadIdx.x; q < ni * nj * nk; q += blockDim.x) sc[q] = cg[q];
  __syncthreads();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int64_t ed = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); ed < e; ed += (int64_t)gridDim.x * wpb) {
    float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int c = lane; c < f; c += 32) {
      const float wv = w[ed * f + c];
      for (int i = 0; i < ni; ++i) {
        const float av = a[(ed * ni + i) * f + c] * wv;
        for (int k = 0; k < nk; ++k) {
          const float ag = av * g[(ed * nk + k) * f + c];
          for (int j = 0; j < nj; ++j) acc[j] = fmaf(sc[(i * nj + j) * nk + k], ag, acc[j]);
        }
      }
    }
    for (int j = 0; j < nj; ++j) {
      const float s = hgb_warp_sum(acc[j]);
      if (lane == 0) out[ed * nj + j] = s;
    }
  }
}

__global__ void tp_w_kernel(const float* __restrict__ a, const float* __restrict__ y, const float* __restrict__ g,
                            const float* __restrict__ cg, int64_t e, int f, int ni, int nj, int nk, float* __restrict__ out) {
  __shared__ float sc[CMAX];
  for (int q = threadIdx.x; q < ni * nj * nk; q += blockDim.x) sc[q] = cg[q];
  __syncthreads();
  const int64_t total = e * f;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ed = idx / f;
    const int c = (int)(idx - ed * f);
    float gv[7];
    for (int k = 0; k < nk; ++k) gv[k] = g[(ed * nk + k) * f + c];
    float acc = 0.f;
    for (int i = 0; i < ni; ++i) {
      const float av = a[(ed * ni + i) * f + c];
      for (int j = 0; j < nj; ++j) {
        const float* cr = sc + (i * nj + j) * nk;
        float t = 0.f;
        for (int k = 0; k < nk; ++k) t = fmaf(cr[k], gv[k], t);
        acc = fmaf(av * y[ed * nj + j], t, acc);
      }
    }
    out[idx] = acc;
  }
}

__global__ void chan_cl_kernel(const float* __restrict__ t, const float* __restrict__ x, int64_t n, int f, int p, int ni,
                               float* __restrict__ out) {
  const int64_t total = n * f * p;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t bc = idx / p;
    const int64_t b = bc / f;
    const int c = (int)(bc - b * f);
    float acc = 0.f;
    for (int i = 0; i < ni; ++i) acc = fmaf(t[idx * ni + i], x[(b * ni + i) * f + c], acc);
    out[idx] = acc;
  }
}

__global__ void chan_ou_kernel(const float* __restrict__ g, const float* __restrict__ x, int64_t n, int f, int p, int ni,
                               float* __restrict__ out) {
  const int64_t total = n * f * p * ni;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx % ni);
    const int64_t bcp = idx / ni;
    const int64_t bc = bcp / p;
    const int64_t b = bc / f;
    const int c = (int)(bc - b * f);
    out[idx] = g[bcp] * x[(b * ni + i) * f + c];
  }
}

__global__ void chan_rp_kernel(const float* __restrict__ g, const float* __restrict__ t, int64_t n, int f, int p, int ni,
                               float* __restrict__ out) {
  const int64_t total = n * ni * f;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % f);
    const int64_t bi = idx / f;
    const int i = (int)(bi % ni);
    const int64_t b = bi / ni;
    const int64_t base = (b * f + c) * p;
    float acc = 0.f;
    for (int q = 0; q < p; ++q) acc = fmaf(g[base + q], t[(base + q) * ni + i], acc);
    out[idx] = acc;
  }
}

}  // namespace

#define HGB_TP_CHECK(name)                                                                                               \
  HGB_REQUIRE(e >= 0 && f >= 1 && ni >= 1 && nj >= 1 && nk >= 1 && ni <= 7 && nj <= 7 && nk <= 7 && cg && out, name ": bad arguments")

// mode 0: tp_out (p0 = a [e, ni, f], p1 = y [e, nj], p2 = w [e, f])        -> out [e, nk, f]
// mode 1: tp_y   (p0 = a [e, ni, f], p1 = g [e, nk, f], p2 = w [e, f])     -> out [e, nj]
// mode 2: tp_w   (p0 = a [e, ni, f], p1 = y [e, nj], p2 = g [e, nk, f])    -> out [e, f]
extern "C" int hgb_mace_tp_path(int32_t mode, const float* p0, const float* p1, const float* p2, const float* cg, int64_t e,
                                int32_t f, int32_t ni, int32_t nj, int32_t nk, float* out, hgb_stream_t stream) {
  HGB_TP_CHECK("mace_tp_path");
  HGB_REQUIRE(mode >= 0 && mode <= 2 && p0 && p1 && p2, "mace_tp_path: bad mode / operands");
  if (e == 0) return HGB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 0) tp_out_kernel<<<hgb_grid_for(e * f, 256), 256, 0, st>>>(p0, p1, p2, cg, e, f, ni, nj, nk, out);
  else if (mode == 1) tp_y_kernel<<<hgb_grid_for(e, 8), 256, 0, st>>>(p0, p1, p2, cg, e, f, ni, nj, nk, out);
  else tp_w_kernel<<<hgb_grid_for(e * f, 256), 256, 0, st>>>(p0, p1, p2, cg, e, f, ni, nj, nk, out);
  HGB_LAUNCH_CHECK("mace_tp_path");
  return HGB_OK;
}

// mode 0: chan_cl (p0 = t [n, f, p, ni], p1 = x [n, ni, f]) -> out [n, f, p]
// mode 1: chan_ou (p0 = g [n, f, p],     p1 = x [n, ni, f]) -> out [n, f, p, ni]
// mode 2: chan_rp (p0 = g [n, f, p],     p1 = t [n, f, p, ni]) -> out [n, ni, f]
extern "C" int hgb_mace_chan_contract(int32_t mode, const float* p0, const float* p1, int64_t n, int32_t f, int32_t p, int32_t ni,
                                      float* out, hgb_stream_t stream) {
  HGB_REQUIRE(mode >= 0 && mode <= 2 && p0 && p1 && out && n >= 0 && f >= 1 && p >= 1 && ni >= 1, "mace_chan_contract: bad arguments");
  if (n == 0) return HGB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 0) chan_cl_kernel<<<hgb_grid_for(n * f * p, 256), 256, 0, st>>>(p0, p1, n, f, p, ni, out);
  else if (mode == 1) chan_ou_kernel<<<hgb_grid_for(n * f * p * ni, 256), 256, 0, st>>>(p0, p1, n, f, p, ni, out);
  else chan_rp_kernel<<<hgb_grid_for(n * ni * f, 256), 256, 0, st>>>(p0, p1, n, f, p, ni, out);
  HGB_LAUNCH_CHECK("mace_chan_contract");
  return HGB_OK;
}
