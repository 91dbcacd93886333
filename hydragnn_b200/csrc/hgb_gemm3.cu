// libhgb.so -- fp32-ACCURATE tensor-core GEMMs for the exact-fp32 mode (configs C1 / C3 / C5 are fp32 configs):
// mma.sync.m16n8k8 TF32 with every product expanded as hi*hi + hi*lo + lo*hi + lo*lo of a TF32 split ("4xTF32": the
// operands are represented to 2^-22, accumulation is fp32 -- the error of a K = 64 dot product is ~1e-7 relative, the same
// order as an fp32 FMA chain; tests/test_gpu_round2.py pins it against fp64).  The SIMT gemm_kernel of hgb_gemm.cu ran the
// large-M Linears of those configs at 13-19 % of the HBM roofline; these are memory-bound.
//
//   gemm3_rows_kernel  C[M, N] = act(A[M, K] op(B) + bias)    M huge, N, K small: forward Linear (B = W [N, K], "NT") and
//                      data gradient (B = W [K, N], "NN").  CTA = 8 warps x 16 rows; the weight panel (<= 64 columns) lives
//                      in shared memory pre-split into hi / lo; A tiles arrive by cp.async.
//   gemm3_tn_kernel    C[Mo, No] = A[R, Mo]^T B[R, No]        R huge: weight gradient.  Row chunks of 64 stream through
//                      shared memory, every CTA keeps the whole [Mo, No] result in registers; per-CTA partials + a
//                      fixed-order reduce (deterministic).
#include "hgb_common.cuh"

namespace {

__device__ __forceinline__ float tf32_hi(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void mma8(float (&c)[4], float a0, float a1, float a2, float a3, float b0, float b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(__float_as_uint(a0)), "r"(__float_as_uint(a1)), "r"(__float_as_uint(a2)), "r"(__float_as_uint(a3)),
                 "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}
// c += A B for one 8-deep k-step, the four split products through a zeroed temporary: the tensor core adds into its C operand with
// truncation, so long accumulation chains are kept in registers with round-to-nearest FADDs (see hgb_attn_tc.cu).
__device__ __forceinline__ void mma4x(float (&c)[4], const float (&ah)[4], const float (&al)[4], float b0h, float b1h, float b0l,
                                      float b1l) {
  float d[4] = {0.f, 0.f, 0.f, 0.f};
  mma8(d, al[0], al[1], al[2], al[3], b0l, b1l);
  mma8(d, al[0], al[1], al[2], al[3], b0h, b1h);
  mma8(d, ah[0], ah[1], ah[2], ah[3], b0l, b1l);
  mma8(d, ah[0], ah[1], ah[2], ah[3], b0h, b1h);
  c[0] += d[0]; c[1] += d[1]; c[2] += d[2]; c[3] += d[3];
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

constexpr int RT = 128;        // rows per CTA tile (8 warps x 16)
constexpr int PN = 64;         // weight-panel columns
constexpr int KMAX = 128;

struct RowsParams {
  const float *a, *b, *bias, *addend;
  float *c, *z;
  int m, n, k, tb;             // tb: 1 = B given as [N, K] (x W^T), 0 = B given as [K, N]
  int64_t lda, ldb, ldc;
  int act, beta_one;
  float act_param;
};

__global__ void __launch_bounds__(256) gemm3_rows_kernel(const RowsParams p) {
  extern __shared__ __align__(16) float smem[];
  const int K = p.k, KS = K + 4;
  float* wh = smem;                    // [PN][KS]
  float* wl = wh + PN * KS;            // [PN][KS]
  float* at = wl + PN * KS;            // [RT][KS]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int ntiles = (p.m + RT - 1) / RT;
  for (int n0 = 0; n0 < p.n; n0 += PN) {
    const int pn = min(PN, p.n - n0);  // multiple of 8
    __syncthreads();
    for (int i = threadIdx.x; i < PN * K; i += 256) {
      int c, kk;
      float v = 0.f;
      if (p.tb) { c = i / K; kk = i - c * K; if (c < pn) v = p.b[(int64_t)(n0 + c) * p.ldb + kk]; }
      else { kk = i / PN; c = i - kk * PN; if (c < pn) v = p.b[(int64_t)kk * p.ldb + n0 + c]; }
      const float hi = tf32_hi(v);
      wh[c * KS + kk] = hi;
      wl[c * KS + kk] = tf32_hi(v - hi);
    }
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int row0 = tile * RT;
      __syncthreads();                 // previous tile's reads of `at` are done (and the panel is staged)
      const int chunks = K / 4;
      for (int i = threadIdx.x; i < RT * chunks; i += 256) {
        const int r = i / chunks, ch = i - r * chunks;
        float* dst = at + r * KS + 4 * ch;
        if (row0 + r < p.m) cp_async16(dst, p.a + (int64_t)(row0 + r) * p.lda + 4 * ch);
        else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      cp_async_wait_all();
      __syncthreads();
      float acc[PN / 8][4];
#pragma unroll
      for (int j = 0; j < PN / 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
      const float* arow = at + (warp * 16 + g) * KS;
      for (int k0 = 0; k0 < K; k0 += 8) {
        float a[4], ah[4], al[4];
        a[0] = arow[k0 + t]; a[1] = arow[8 * KS + k0 + t]; a[2] = arow[k0 + t + 4]; a[3] = arow[8 * KS + k0 + t + 4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { ah[i] = tf32_hi(a[i]); al[i] = tf32_hi(a[i] - ah[i]); }
#pragma unroll
        for (int j = 0; j < PN / 8; ++j) {
          if (8 * j < pn) {
            const int o = (8 * j + g) * KS + k0 + t;
            mma4x(acc[j], ah, al, wh[o], wh[o + 4], wl[o], wl[o + 4]);
          }
        }
      }
      // epilogue: rows g / g + 8 of this warp's 16, columns 2t, 2t + 1 of every 8-column block
      const int ra = row0 + warp * 16 + g, rb = ra + 8;
#pragma unroll
      for (int j = 0; j < PN / 8; ++j) {
        if (8 * j >= pn) continue;
        const int col = n0 + 8 * j + 2 * t;
        const float b0 = p.bias ? __ldg(p.bias + col) : 0.f, b1 = p.bias ? __ldg(p.bias + col + 1) : 0.f;
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
          const int r = hrow ? rb : ra;
          if (r >= p.m) continue;
          float v0 = acc[j][2 * hrow] + b0, v1 = acc[j][2 * hrow + 1] + b1;
          if (p.beta_one) {
            const float2 old = *reinterpret_cast<const float2*>(p.c + (int64_t)r * p.ldc + col);
            v0 += old.x; v1 += old.y;
          }
          if (p.z) *reinterpret_cast<float2*>(p.z + (int64_t)r * p.ldc + col) = make_float2(v0, v1);
          if (p.act) { v0 = hgb_act(v0, p.act, p.act_param); v1 = hgb_act(v1, p.act, p.act_param); }
          *reinterpret_cast<float2*>(p.c + (int64_t)r * p.ldc + col) = make_float2(v0, v1);
        }
      }
    }
  }
}

// ---- C[Mo, No] = A[R, Mo]^T B[R, No] --------------------------------------------------------------------------------
constexpr int TC_ROWS = 64;          // rows per chunk
constexpr int TN_MAXT = 16;          // output tiles (16 x 8) per warp

__global__ void __launch_bounds__(256) gemm3_tn_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t lda,
                                                       int64_t ldb, int r_total, int mo, int no, int chunks_per_cta,
                                                       float* __restrict__ partial) {
  extern __shared__ __align__(16) float smem[];
  const int AS = mo + 8, BS = no + 8;
  float* sa = smem;                    // [TC_ROWS][AS]
  float* sb = sa + TC_ROWS * AS;       // [TC_ROWS][BS]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int MT = mo / 16, NTL = no / 8, T = MT * NTL;
  float acc[TN_MAXT][4];
#pragma unroll
  for (int i = 0; i < TN_MAXT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  const int nchunks = (r_total + TC_ROWS - 1) / TC_ROWS;
  const int c_beg = blockIdx.x * chunks_per_cta, c_end = min(nchunks, c_beg + chunks_per_cta);
  for (int ch = c_beg; ch < c_end; ++ch) {
    const int row0 = ch * TC_ROWS;
    __syncthreads();
    const int ca = mo / 4, cb = no / 4;
    for (int i = threadIdx.x; i < TC_ROWS * ca; i += 256) {
      const int r = i / ca, c4 = i - r * ca;
      float* dst = sa + r * AS + 4 * c4;
      if (row0 + r < r_total) cp_async16(dst, a + (int64_t)(row0 + r) * lda + 4 * c4);
      else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int i = threadIdx.x; i < TC_ROWS * cb; i += 256) {
      const int r = i / cb, c4 = i - r * cb;
      float* dst = sb + r * BS + 4 * c4;
      if (row0 + r < r_total) cp_async16(dst, b + (int64_t)(row0 + r) * ldb + 4 * c4);
      else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cp_async_wait_all();
    __syncthreads();
#pragma unroll
    for (int i = 0; i < TN_MAXT; ++i) {
      const int id = warp + 8 * i;
      if (id >= T) break;
      const int mt = id % MT, nt = id / MT;
      for (int k0 = 0; k0 < TC_ROWS; k0 += 8) {
        // A operand = A^T: element (m, k) = sa[k][m]
        float av[4], ah[4], al[4];
        av[0] = sa[(k0 + t) * AS + mt * 16 + g];     av[1] = sa[(k0 + t) * AS + mt * 16 + g + 8];
        av[2] = sa[(k0 + t + 4) * AS + mt * 16 + g]; av[3] = sa[(k0 + t + 4) * AS + mt * 16 + g + 8];
#pragma unroll
        for (int q = 0; q < 4; ++q) { ah[q] = tf32_hi(av[q]); al[q] = tf32_hi(av[q] - ah[q]); }
        const float b0 = sb[(k0 + t) * BS + nt * 8 + g], b1 = sb[(k0 + t + 4) * BS + nt * 8 + g];
        const float b0h = tf32_hi(b0), b1h = tf32_hi(b1);
        mma4x(acc[i], ah, al, b0h, b1h, tf32_hi(b0 - b0h), tf32_hi(b1 - b1h));
      }
    }
  }
  float* out = partial + (int64_t)blockIdx.x * mo * no;
#pragma unroll
  for (int i = 0; i < TN_MAXT; ++i) {
    const int id = warp + 8 * i;
    if (id >= T) break;
    const int mt = id % MT, nt = id / MT;
    const int r = mt * 16 + g, c = nt * 8 + 2 * t;
    *reinterpret_cast<float2*>(out + (int64_t)r * no + c) = make_float2(acc[i][0], acc[i][1]);
    *reinterpret_cast<float2*>(out + (int64_t)(r + 8) * no + c) = make_float2(acc[i][2], acc[i][3]);
  }
}

__global__ void gemm3_reduce_kernel(const float* __restrict__ partial, int nparts, int mo, int no, float* __restrict__ c, int64_t ldc,
                                    int beta_one) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= mo * no) return;
  double s = 0.0;                       // fixed order, fp64: hundreds of partials of a cancelling sum
  for (int q = 0; q < nparts; ++q) s += (double)partial[(int64_t)q * mo * no + idx];
  float* dst = c + (int64_t)(idx / no) * ldc + (idx % no);
  *dst = beta_one ? *dst + (float)s : (float)s;
}

inline int tn_grid(int r_total) {
  const int nchunks = (r_total + TC_ROWS - 1) / TC_ROWS;
  return nchunks < HGB_NUM_SMS * 2 ? nchunks : HGB_NUM_SMS * 2;
}

}  // namespace

// which (m, n, k, trans_a, trans_b) the tensor-core fp32 path takes; everything else stays on the SIMT kernels
extern "C" int32_t hgb_gemm3_supported(int32_t m, int32_t n, int32_t k, int32_t trans_a, int32_t trans_b, int64_t lda, int64_t ldb,
                                       int64_t ldc) {
  if (!trans_a) {      // rows kernel: C[m, n] = A[m, k] op(B)
    return (m >= 512 && k >= 8 && k <= KMAX && k % 8 == 0 && n >= 8 && n % 8 == 0 && lda % 4 == 0 && ldc % 2 == 0) ? 1 : 0;
  }
  if (trans_b) return 0;
  // TN: C[m, n] = A[k, m]^T B[k, n], reduction over k rows
  return (k >= 2048 && m % 16 == 0 && n % 8 == 0 && m >= 16 && n >= 8 && (m / 16) * (n / 8) <= 8 * TN_MAXT && lda % 4 == 0 &&
          ldb % 4 == 0 && m <= 512 && n <= 512) ? 1 : 0;
}

extern "C" int64_t hgb_gemm3_workspace_bytes(int32_t m, int32_t n, int32_t k, int32_t trans_a) {
  if (!trans_a) return 0;
  return (int64_t)tn_grid(k) * m * n * 4 + 256;
}

// C = act(op(A) op(B) + bias) [+ C if beta_one]; z (optional, rows form only) receives the pre-activation.
extern "C" int hgb_gemm3(const float* a, const float* b, float* c, int32_t m, int32_t n, int32_t k, int32_t trans_a, int32_t trans_b,
                         int64_t lda, int64_t ldb, int64_t ldc, int32_t beta_one, const float* bias, int32_t act, float act_param,
                         float* z, void* workspace, hgb_stream_t stream) {
  HGB_REQUIRE(a && b && c && hgb_gemm3_supported(m, n, k, trans_a, trans_b, lda, ldb, ldc), "gemm3: unsupported shape %d x %d x %d (ta %d tb %d)",
              m, n, k, trans_a, trans_b);
  HGB_REQUIRE(((uintptr_t)a % 16 == 0) && ((uintptr_t)c % 8 == 0), "gemm3: operands must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (!trans_a) {
    RowsParams p{a, b, bias, nullptr, c, z, m, n, k, trans_b, lda, ldb, ldc, act, beta_one, act_param};
    const size_t bytes = (size_t)(2 * PN + RT) * (k + 4) * 4;
    cudaFuncSetAttribute(gemm3_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    const int ntiles = (m + RT - 1) / RT;
    const int per_sm = bytes > 110 * 1024 ? 1 : (bytes > 72 * 1024 ? 2 : 3);
    const int grid = ntiles < HGB_NUM_SMS * per_sm ? ntiles : HGB_NUM_SMS * per_sm;
    gemm3_rows_kernel<<<grid, 256, bytes, st>>>(p);
    HGB_LAUNCH_CHECK("gemm3_rows");
    return HGB_OK;
  }
  HGB_REQUIRE(workspace && !bias && !act && !z && ((uintptr_t)b % 16 == 0), "gemm3 (TN): needs a workspace, no epilogue");
  const int grid = tn_grid(k);
  const int nchunks = (k + TC_ROWS - 1) / TC_ROWS;
  const int cpc = (nchunks + grid - 1) / grid;
  const size_t bytes = (size_t)TC_ROWS * (m + 8 + n + 8) * 4;
  cudaFuncSetAttribute(gemm3_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  const int used = (nchunks + cpc - 1) / cpc;
  gemm3_tn_kernel<<<used, 256, bytes, st>>>(a, b, lda, ldb, k, m, n, cpc, (float*)workspace);
  HGB_LAUNCH_CHECK("gemm3_tn");
  gemm3_reduce_kernel<<<(m * n + 255) / 256, 256, 0, st>>>((const float*)workspace, used, m, n, c, ldc, beta_one);
  HGB_LAUNCH_CHECK("gemm3_reduce");
  return HGB_OK;
}
