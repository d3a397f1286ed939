// Training-side kernels of the data-parallel DQN step (sm_100a): packed observation frames (replay storage and
// host wire format), n-step returns over the device-resident replay ring, fused Adam over the flat parameter buffer.
//
// Reference:
//   replay contract     graph_env/env/utils/collectors/multi_agent_collector.py:240-308 (one sub-buffer per (env, agent);
//                       a transition is completed by the agent's next observation and carries the reward of the world
//                       step in between) + tianshou VectorReplayBuffer(ignore_obs_next=True) (l_dgn.py:170-182)
//   n-step return       tianshou BasePolicy.compute_nstep_return as called by DQNPolicy.process_fn
//                       (estimation_step = --n-step 4, discount_factor = --gamma 0.99; l_dgn.py:69-76)
//   optimiser           torch.optim.Adam(net.parameters(), lr=args.lr)  (l_dgn.py:66)
// In the batched form every active agent of an episode acts in every round until its TTL ends (graph.py:330-345), so
// the successor of transition (round r, episode b, agent a) is (r+1, b, a): the replay ring is a dense
// [ring round][episode][agent] array and chains are implicit.
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------- packed frames
// obs row (graph.py:254-271) = [x, y, degree, messages_transmitted, last action, interested, has_message, dm] fp32.
// Columns 2..7 are small non-negative integers / flags: packed as one word
//   bits 0..2 has_message | interested | action, bits 3..8 messages (< 64), bits 9..16 degree (< 256), bit 31 dm
// (bits 0..16 are exactly the feature key of dgn_forward_bf16.cu).  12 bytes per node instead of 32.
struct PackedNode { uint32_t x, y, w; };     // x, y: fp32 bit patterns; accessed as three 4-byte words (12-byte stride)

__global__ void obs_pack_kernel(const float* __restrict__ obs, long long rows, PackedNode* __restrict__ out, int* __restrict__ errors) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float4 lo = reinterpret_cast<const float4*>(obs)[r * 2], hi = reinterpret_cast<const float4*>(obs)[r * 2 + 1];
  const float deg = lo.z, msgs = lo.w, act = hi.x, intr = hi.y, hm = hi.z, dm = hi.w;
  const bool ok = deg >= 0.f && deg < 256.f && deg == floorf(deg) && msgs >= 0.f && msgs < 64.f && msgs == floorf(msgs) &&
                  (act == 0.f || act == 1.f) && (intr == 0.f || intr == 1.f) && (hm == 0.f || hm == 1.f) && (dm == 0.f || dm == 1.f);
  uint32_t w = 0;
  if (ok) w = ((((((uint32_t)deg << 6) | (uint32_t)msgs) << 1 | (uint32_t)act) << 1 | (uint32_t)intr) << 1 | (uint32_t)hm) | ((uint32_t)dm << 31);
  else if (errors) atomicAdd(errors, 1);
  uint32_t* o = reinterpret_cast<uint32_t*>(out) + r * 3;
  o[0] = __float_as_uint(lo.x); o[1] = __float_as_uint(lo.y); o[2] = w;
}

// out row m = frame rows of src_row[m] (NULL: m), N nodes each, expanded to 8 floats per node; with `agent` the row is
// the reference's agent observation [8N + 1] whose last column is the controlling index (graph.py:181-200).
__global__ void obs_unpack_kernel(const PackedNode* __restrict__ in, const long long* __restrict__ src_frame, const int* __restrict__ agent,
                                  int N, long long n_frames, long long out_stride, float* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_frames * N) return;
  const long long m = t / N;
  const int i = (int)(t - m * N);
  const long long f = src_frame ? src_frame[m] : m;
  const uint32_t* pw = reinterpret_cast<const uint32_t*>(in) + (f * N + i) * 3;
  const uint32_t w = pw[2];
  float* o = out + m * out_stride + (long long)i * 8;
  o[0] = __uint_as_float(pw[0]); o[1] = __uint_as_float(pw[1]);
  o[2] = (float)((w >> 9) & 255u); o[3] = (float)((w >> 3) & 63u); o[4] = (float)((w >> 2) & 1u);
  o[5] = (float)((w >> 1) & 1u); o[6] = (float)(w & 1u); o[7] = (float)(w >> 31);
  if (agent && i == 0) out[m * out_stride + (long long)N * 8] = (float)agent[m];
}

// ---------------------------------------------------------------------------------------------- n-step returns
// flags bit 0: the agent acted in that round (a stored transition), bit 1: terminated after it (graph.py:330-334;
// tianshou value mask: no bootstrap).  ret[m] = sum_{k < K} gamma^k rew[(rho+k) % R][b][a], K = steps until the chain
// terminates or n_step, accumulated in fp64 like tianshou's numpy code, then narrowed.  boot_round[m] = ring round of the
// bootstrap observation ((rho + n) % R) when the chain is still alive after n steps, else -1; boot_gamma[m] = gamma^n.
__global__ void nstep_return_kernel(const double* __restrict__ rew, const uint8_t* __restrict__ flags, int R, long long B, int N,
                                    const int* __restrict__ rho, const int* __restrict__ ep, const int* __restrict__ agent, int M,
                                    int n_step, double gamma, float* __restrict__ ret, int* __restrict__ boot_round,
                                    float* __restrict__ boot_gamma) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const long long cell = (long long)ep[m] * N + agent[m];
  double g = 0.0, gk = 1.0;
  int r = rho[m];
  bool alive = true;
  for (int k = 0; k < n_step && alive; ++k) {
    const long long at = (long long)r * B * N + cell;
    g += gk * rew[at];
    gk *= gamma;
    alive = !(flags[at] & 2u);
    r = r + 1 == R ? 0 : r + 1;
  }
  ret[m] = (float)g;
  boot_round[m] = alive ? r : -1;
  boot_gamma[m] = (float)gk;
}

// ---------------------------------------------------------------------------------------------- Adam
// torch.optim.Adam (amsgrad=False, maximize=False) on flat fp32 buffers:
//   g = grad * grad_scale (+ weight_decay * p);  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v,
                            long long n, float b1, float b2, float eps, float wd, float step_size, float inv_bc2_sqrt, float grad_scale) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float g = grad[i] * grad_scale;
    const float pi = p[i];
    if (wd != 0.f) g = fmaf(wd, pi, g);
    const float mi = fmaf(b1, m[i], (1.f - b1) * g);
    const float vi = fmaf(b2, v[i], (1.f - b2) * g * g);
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) * inv_bc2_sqrt + eps;
    p[i] = pi - step_size * (mi / denom);
  }
}

}  // namespace

extern "C" int mls_obs_pack(const float* obs, int64_t n_node_rows, void* packed, int32_t* errors, void* stream) {
  MLS_CHECK_ARG(obs && packed && n_node_rows >= 0, "NULL argument");
  if (n_node_rows == 0) return MLS_OK;
  obs_pack_kernel<<<(unsigned)((n_node_rows + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      obs, n_node_rows, reinterpret_cast<PackedNode*>(packed), errors);
  mls_count_launch();
  MLS_LAUNCH_CHECK();
  return MLS_OK;
}

extern "C" int mls_obs_unpack(const void* packed, const int64_t* src_frame, const int32_t* agent, int32_t n_nodes, int64_t n_frames,
                              int64_t out_stride, float* out, void* stream) {
  MLS_CHECK_ARG(packed && out && n_nodes > 0 && n_frames >= 0, "NULL argument");
  MLS_CHECK_ARG(out_stride >= (int64_t)n_nodes * 8 + (agent ? 1 : 0), "output rows too short");
  if (n_frames == 0) return MLS_OK;
  const long long total = (long long)n_frames * n_nodes;
  obs_unpack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const PackedNode*>(packed), reinterpret_cast<const long long*>(src_frame), agent, n_nodes, n_frames, out_stride, out);
  mls_count_launch();
  MLS_LAUNCH_CHECK();
  return MLS_OK;
}

extern "C" int mls_nstep_returns(const double* rew, const uint8_t* flags, int32_t ring_rounds, int64_t n_episodes, int32_t n_nodes,
                                 const int32_t* round_index, const int32_t* episode, const int32_t* agent, int32_t n_samples,
                                 int32_t n_step, double gamma, float* returns, int32_t* boot_round, float* boot_gamma, void* stream) {
  MLS_CHECK_ARG(rew && flags && round_index && episode && agent && returns && boot_round && boot_gamma, "NULL argument");
  MLS_CHECK_ARG(ring_rounds > 0 && n_step >= 1 && n_step <= ring_rounds, "n_step must be in [1, ring rounds]");
  if (n_samples <= 0) return MLS_OK;
  nstep_return_kernel<<<(n_samples + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      rew, flags, ring_rounds, n_episodes, n_nodes, round_index, episode, agent, n_samples, n_step, gamma, returns, boot_round, boot_gamma);
  mls_count_launch();
  MLS_LAUNCH_CHECK();
  return MLS_OK;
}

extern "C" int mls_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                             float beta2, float eps, float weight_decay, int64_t step, float grad_scale, void* stream) {
  MLS_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && n >= 0 && step >= 1, "bad argument");
  if (n == 0) return MLS_OK;
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1), inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  adam_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, beta1, beta2, eps,
                                                                                     weight_decay, step_size, inv_bc2_sqrt, grad_scale);
  mls_count_launch();
  MLS_LAUNCH_CHECK();
  return MLS_OK;
}
