// trex_policy.h -- fused policy / value forward for the rollout loop, included by trex_capi.cu.
//
// What the reference's trainer evaluates once per env step around TrexBulletEnv.step (trex_gym/trex_train.py:48,107:
// baselines ppo2 with MlpPolicy [RECALL]: two tanh MLP trunks 75 -> 64 -> 64, a linear Gaussian-mean head (25) with a
// state-independent log-std, a linear value head (1); actions a = mean + exp(logstd) * eps, neglogp of DiagGaussianPd),
// preceded by VecNormalize's observation filter (trex_train.py:45).  One kernel reads the observation rows in place
// (normalising them on the way in), keeps every activation on chip, and writes action / neglogp / value.
//
// Mapping: 128 rows (environments) per CTA, one row per thread.  The row's 64 accumulators live in registers; the
// layer's weights sit in shared memory and are read as 128-bit broadcasts (every thread reads the same weight), the
// row's inputs come from a row-major tile (stride 75 = 11 mod 32: conflict-free) resp. a transposed hidden tile.
// FP32 FFMA throughout: the op is GEMM-shaped but 2.6 GFLOP per 65,536 rows (~1 % of an env step) and must match an
// FP32 reference to 1e-5, so tensor-core (tf32/bf16) arithmetic is not used.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace trex_policy {

constexpr int OBS = 75, HID = 64, ACT = 25, ROWS = 128, ACT_PAD = 28;
// packed parameter vector (floats), row-major [in][out] matrices:
//   pi: W1[75][64] b1[64] W2[64][64] b2[64] Wo[64][25] bo[25] | vf: W1[75][64] b1[64] W2[64][64] b2[64] Wo[64][1] bo[1] | logstd[25]
constexpr int TRUNK_COMMON = OBS * HID + HID + HID * HID + HID;
constexpr int PI_SIZE = TRUNK_COMMON + HID * ACT + ACT;
constexpr int VF_SIZE = TRUNK_COMMON + HID + 1;
constexpr int PARAM_COUNT = PI_SIZE + VF_SIZE + ACT;
constexpr int SMEM_FLOATS = ROWS * OBS + HID * ROWS + OBS * HID + HID;  // x tile, hidden tile, weights, bias
constexpr size_t SMEM_BYTES = SMEM_FLOATS * sizeof(float);

__device__ __forceinline__ void philox_round10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

// acc[j] += x * w[j], j < 64, weights of input i at ws[i * 64 ..] (four floats per 128-bit broadcast load)
__device__ __forceinline__ void axpy64(float (&acc)[HID], float x, const float* __restrict__ wrow) {
  const float4* w4 = reinterpret_cast<const float4*>(wrow);
#pragma unroll
  for (int j = 0; j < HID / 4; j++) {
    const float4 w = w4[j];
    acc[4 * j] = fmaf(x, w.x, acc[4 * j]);
    acc[4 * j + 1] = fmaf(x, w.y, acc[4 * j + 1]);
    acc[4 * j + 2] = fmaf(x, w.z, acc[4 * j + 2]);
    acc[4 * j + 3] = fmaf(x, w.w, acc[4 * j + 3]);
  }
}

__device__ __forceinline__ void stage(float* dst, const float* __restrict__ src, int n) {
  for (int i = threadIdx.x; i < n; i += ROWS) dst[i] = __ldg(src + i);
}

// the two hidden layers of one trunk; leaves tanh(h2) of row r in hs[j * ROWS + r]
__device__ __forceinline__ void trunk(const float* __restrict__ p, const float* xs, float* hs, float* ws, float* bs, int r) {
  __syncthreads();  // previous users of ws are done
  stage(ws, p, OBS * HID);
  stage(bs, p + OBS * HID, HID);
  __syncthreads();
  float acc[HID];
#pragma unroll
  for (int j = 0; j < HID; j++) acc[j] = bs[j];
#pragma unroll 1
  for (int i = 0; i < OBS; i++) axpy64(acc, xs[r * OBS + i], ws + i * HID);
#pragma unroll
  for (int j = 0; j < HID; j++) hs[j * ROWS + r] = tanhf(acc[j]);
  __syncthreads();
  stage(ws, p + OBS * HID + HID, HID * HID);
  stage(bs, p + OBS * HID + HID + HID * HID, HID);
  __syncthreads();
#pragma unroll
  for (int j = 0; j < HID; j++) acc[j] = bs[j];
#pragma unroll 1
  for (int i = 0; i < HID; i++) axpy64(acc, hs[i * ROWS + r], ws + i * HID);
  // every thread reads and writes only its own column of hs: no barrier needed in between
#pragma unroll
  for (int j = 0; j < HID; j++) hs[j * ROWS + r] = tanhf(acc[j]);
}

__global__ void __launch_bounds__(ROWS, 2)
forward_kernel(const float* __restrict__ obs, const float* __restrict__ ob_mean, const float* __restrict__ ob_var, float eps, float clip,
               const float* __restrict__ params, uint32_t seed, uint64_t step, int64_t env_offset, int deterministic,
               float* __restrict__ action, float* __restrict__ neglogp, float* __restrict__ value, float* __restrict__ mean_out,
               int64_t n_rows) {
  extern __shared__ __align__(16) float sm[];
  float* xs = sm;                       // [ROWS][OBS] normalised observations (later: the action tile [ROWS][ACT])
  float* hs = xs + ROWS * OBS;          // [HID][ROWS]
  float* ws = hs + HID * ROWS;          // current layer's weights
  float* bs = ws + OBS * HID;           // current layer's bias
  const int r = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * ROWS;
  const int64_t row = row0 + r;
  const int rows_here = (int)((n_rows - row0) < ROWS ? (n_rows - row0) : ROWS);
  // observation tile: coalesced read, VecNormalize filter on the way in (ob_mean == NULL: identity)
  for (int i = threadIdx.x; i < ROWS * OBS; i += ROWS) {
    float x = 0.0f;
    if (i < rows_here * OBS) {
      x = obs[row0 * OBS + i];
      if (ob_mean != nullptr) {
        const int d = i % OBS;
        const float z = (x - __ldg(ob_mean + d)) / sqrtf(__ldg(ob_var + d) + eps);
        x = fminf(fmaxf(z, -clip), clip);
      }
    }
    xs[i] = x;
  }
  // ---- value trunk first (the policy trunk then reuses xs for the action tile) ----
  const float* pv = params + PI_SIZE;
  trunk(pv, xs, hs, ws, bs, r);
  {
    __syncthreads();
    stage(ws, pv + TRUNK_COMMON, HID + 1);
    __syncthreads();
    float v = ws[HID];
#pragma unroll 8
    for (int i = 0; i < HID; i++) v = fmaf(hs[i * ROWS + r], ws[i], v);
    if (row < n_rows && value != nullptr) value[row] = v;
  }
  // ---- policy trunk ----
  trunk(params, xs, hs, ws, bs, r);
  __syncthreads();
  // head weights padded to ACT_PAD columns: ws[i * ACT_PAD + d]; bias in bs; logstd behind it
  for (int i = threadIdx.x; i < HID * ACT_PAD; i += ROWS) {
    const int hi = i / ACT_PAD, d = i % ACT_PAD;
    ws[i] = d < ACT ? __ldg(params + TRUNK_COMMON + hi * ACT + d) : 0.0f;
  }
  if (threadIdx.x < ACT) {
    bs[threadIdx.x] = __ldg(params + TRUNK_COMMON + HID * ACT + threadIdx.x);
    bs[32 + threadIdx.x] = __ldg(params + PI_SIZE + VF_SIZE + threadIdx.x);
  }
  __syncthreads();
  float mu[ACT_PAD];
#pragma unroll
  for (int d = 0; d < ACT_PAD; d++) mu[d] = d < ACT ? bs[d] : 0.0f;
#pragma unroll 1
  for (int i = 0; i < HID; i++) {
    const float x = hs[i * ROWS + r];
    const float4* w4 = reinterpret_cast<const float4*>(ws + i * ACT_PAD);
#pragma unroll
    for (int j = 0; j < ACT_PAD / 4; j++) {
      const float4 w = w4[j];
      mu[4 * j] = fmaf(x, w.x, mu[4 * j]);
      mu[4 * j + 1] = fmaf(x, w.y, mu[4 * j + 1]);
      mu[4 * j + 2] = fmaf(x, w.z, mu[4 * j + 2]);
      mu[4 * j + 3] = fmaf(x, w.w, mu[4 * j + 3]);
    }
  }
  // ---- a = mean + exp(logstd) * eps, eps ~ N(0,1) from Philox keyed by (seed, global env, step); neglogp of the diagonal Gaussian
  float nrm[ACT_PAD];
#pragma unroll
  for (int d = 0; d < ACT_PAD; d++) nrm[d] = 0.0f;
  if (!deterministic) {
    const uint64_t genv = (uint64_t)(env_offset + row);
#pragma unroll
    for (int b = 0; b < ACT_PAD / 4; b++) {
      uint32_t c[4] = {(uint32_t)genv, (uint32_t)(genv >> 32), (uint32_t)step, (uint32_t)(step >> 32) ^ ((uint32_t)b << 24)};
      philox_round10(c, seed, 0x9a55u);
      // two Box-Muller pairs per block
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const float u1 = ((float)(c[2 * h] >> 8) + 1.0f) * (1.0f / 16777216.0f);  // (0, 1]
        const float u2 = (float)(c[2 * h + 1] >> 8) * (1.0f / 16777216.0f);       // [0, 1)
        const float rad = sqrtf(-2.0f * logf(u1));
        float sn, cs;
        sincospif(2.0f * u2, &sn, &cs);
        nrm[4 * b + 2 * h] = rad * cs;
        nrm[4 * b + 2 * h + 1] = rad * sn;
      }
    }
  }
  float nlp = 0.5f * 1.8378770664093453f * (float)ACT;  // 0.5 * log(2 pi) * D
  __syncthreads();                                       // xs is free now: it becomes the action tile [ROWS][ACT]
#pragma unroll
  for (int d = 0; d < ACT; d++) {
    const float ls = bs[32 + d];
    const float a = fmaf(expf(ls), nrm[d], mu[d]);
    nlp += 0.5f * nrm[d] * nrm[d] + ls;                  // ((a - mean) / std)^2 = eps^2
    xs[r * ACT + d] = a;
    ws[r * ACT + d] = mu[d];                             // the means, for the optional output (ws is free after the barrier)
  }
  if (row < n_rows && neglogp != nullptr) neglogp[row] = nlp;
  __syncthreads();
  for (int i = threadIdx.x; i < rows_here * ACT; i += ROWS) {
    action[row0 * ACT + i] = xs[i];
    if (mean_out != nullptr) mean_out[row0 * ACT + i] = ws[i];
  }
}

}  // namespace trex_policy
