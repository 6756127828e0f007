// One-launch Adam over the flat parameter / gradient buffers of the data-parallel step.
//
// The reference trains with torch.optim.Adam (exp/run_exp.py:343), which at this model size (~170 small tensors) is a
// dozen multi-tensor launches per step plus a memset of the gradients; with every gradient already living in one flat
// bucket (cwn_b200/dist.py) and the parameters flattened the same way, the update is a single streaming kernel that
// also clears the gradient for the next step. The step counter lives on the device (CUDA-graph friendly) and is
// advanced by the last CTA to finish.
#include "common.cuh"

namespace cwn {

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n, float lr,
            float b1, float b2, float eps, float wd, int32_t* step, int32_t* counter, int zero_grad) {
  pdl_trigger();  // programmatic dependent launch: see common.cuh
  pdl_wait();
  const int t = *reinterpret_cast<volatile int32_t*>(step) + 1;
  const float bc1 = (float)(1.0 - pow((double)b1, (double)t));
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, (double)t));
  const float step_size = lr / bc1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i];
    const float pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - step_size * (mi / denom);
    if (zero_grad) g[i] = 0.f;
  }
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(counter, 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (s_last && threadIdx.x == 0) {  // every CTA has read `step` before it bumped the counter
    *step = t;
    *counter = 0;
  }
}

}  // namespace cwn

using namespace cwn;

extern "C" int cwn_adam_step_f32(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                                 float beta1, float beta2, float eps, float weight_decay, int32_t* step,
                                 int32_t* counter, int32_t zero_grad, cwn_stream_t stream) {
  if (n < 0) return fail(CWN_E_SHAPE, "cwn_adam_step_f32: negative size");
  if (n == 0) return CWN_OK;
  if (!param || !grad || !exp_avg || !exp_avg_sq || !step || !counter) return fail(CWN_E_NULL, "cwn_adam_step_f32");
  int64_t blocks = (n + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  launch_pdl(adam_kernel, (int)blocks, 256, 0, (cudaStream_t)stream, param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                             weight_decay, step, counter, zero_grad);
  return launched("adam_kernel");
}
