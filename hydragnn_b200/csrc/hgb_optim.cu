// libhgb.so -- loss (value + gradient in one pass) and the fused flat AdamW step.
#include "hgb_common.cuh"

// single block: deterministic tree reduction; count is small (number of targets in the batch)
__global__ void loss_fwd_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ target, int64_t count, int mode,
                                    float gscale, float* __restrict__ loss, float* __restrict__ gpred,
                                    const int32_t* __restrict__ valid_rows, int row_width) {
  __shared__ float sm[1024];
  float acc = 0.f;
  const int64_t total = count;
  if (valid_rows) {          // capacity-padded batch: only the first *valid_rows rows are real (hydragnn_b200/padded.py)
    const int64_t v = (int64_t)valid_rows[0] * row_width;
    count = v < count ? (v > 0 ? v : 1) : count;
    for (int64_t i = count + threadIdx.x; gpred && i < total; i += blockDim.x) gpred[i] = 0.f;
  }
  const float inv = 1.f / (float)count;
  for (int64_t i = threadIdx.x; i < count; i += blockDim.x) {
    const float d = pred[i] - target[i];
    if (mode == 0) {
      acc += d * d;
      if (gpred) gpred[i] = 2.f * d * inv * gscale;
    } else {
      acc += fabsf(d);
      if (gpred) gpred[i] = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * inv * gscale;
    }
  }
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = sm[0] * inv;
}

extern "C" int hgb_loss_fwd_bwd(const float* pred, const float* target, int64_t count, int32_t mode, float gscale, float* loss,
                                float* gpred, const int32_t* valid_rows, int32_t row_width, hgb_stream_t stream) {
  HGB_REQUIRE(count > 0 && pred && target && loss && (mode == 0 || mode == 1), "loss_fwd_bwd: bad arguments");
  HGB_REQUIRE(!valid_rows || row_width > 0, "loss_fwd_bwd: row_width must be positive with valid_rows");
  loss_fwd_bwd_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(pred, target, count, mode, gscale, loss, gpred, valid_rows, row_width);
  HGB_LAUNCH_CHECK("loss_fwd_bwd");
  return HGB_OK;
}

// torch.optim.AdamW semantics (decoupled weight decay, bias correction, eps outside the sqrt):
//   p *= 1 - lr*wd;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
//   p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                             int64_t count, float lr, float b1, float b2, float eps, float wd, float gscale,
                             const float* __restrict__ step_dev, const float* __restrict__ hyper_dev) {
  if (hyper_dev) {           // learning rate / gradient scale live on the device: a captured step follows the scheduler
    lr = hyper_dev[0];
    gscale = hyper_dev[1];
  }
  const float t = step_dev[0] + 1.f;
  const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    float pi = p[i] * (1.f - lr * wd);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    pi -= step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
    p[i] = pi;
  }
}
__global__ void step_inc_kernel(float* step_dev) { step_dev[0] += 1.f; }

extern "C" int hgb_adamw_step(float* p, const float* g, float* m, float* v, int64_t count, float lr, float beta1, float beta2,
                              float eps, float weight_decay, float grad_scale, float* step_dev, const float* hyper_dev,
                              hgb_stream_t stream) {
  HGB_REQUIRE(count >= 0 && p && g && m && v && step_dev, "adamw_step: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (count > 0) {
    adamw_kernel<<<hgb_grid_for(count, 256), 256, 0, st>>>(p, g, m, v, count, lr, beta1, beta2, eps, weight_decay, grad_scale, step_dev,
                                                           hyper_dev);
    HGB_LAUNCH_CHECK("adamw");
  }
  step_inc_kernel<<<1, 1, 0, st>>>(step_dev);
  HGB_LAUNCH_CHECK("adamw_step_inc");
  return HGB_OK;
}
