// Persistent bidirectional LSTM recurrence for Force_APTAI's RNN tail (models/modules.py:190-214: nn.LSTM(256, 256,
// bidirectional) over packed sequences), sm_100a thread-block clusters + distributed shared memory.
//
// The input projection W_ih x_t + b_ih + b_hh of every frame and both directions is one GEMM done beforehand
// (gates_in [B][T][2][4*H]).  This kernel runs the T sequential steps: one cluster of 8 CTAs per (direction, chunk of
// 8 utterances); CTA c owns hidden units [32c, 32c+32): its 4 x 32 rows of W_hh (128 x 256 fp32 = 128 KB) stay in
// shared memory for the whole sequence, h_{t-1} of the chunk (256 x 8 fp32) is replicated in every CTA and refreshed
// each step by st.shared::cluster writes from the owners, one barrier.cluster per step.  fp32 FMA throughout (the
// recurrence is latency-bound: T steps of a 256-deep dot product; tensor cores would not shorten the chain).
//
// Packed-sequence semantics (pack_padded_sequence / pad_packed_sequence): utterance b runs len[b] steps, the reverse
// direction starts at its last valid frame, outputs beyond len[b] are zero.
#include "common.h"
#include "ptx.cuh"

#include <math.h>

namespace aptai {

constexpr int LS_H = 256;          // hidden size (and input size) of Force_APTAI's LSTM
constexpr int LS_CL = 8;           // CTAs per cluster
constexpr int LS_U = LS_H / LS_CL; // hidden units per CTA
constexpr int LS_R = 4 * LS_U;     // gate rows per CTA
constexpr int LS_BC = 8;           // utterances per cluster
constexpr int LS_THREADS = 256;
constexpr int LS_SMEM = (LS_H * LS_R + 2 * LS_H * LS_BC + LS_R * LS_BC + LS_U * LS_BC) * 4;   // W^T | h x2 | gates | stage

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

__global__ void __cluster_dims__(LS_CL, 1, 1) __launch_bounds__(LS_THREADS, 1)
bilstm_kernel(const float* __restrict__ gates_in, const float* __restrict__ w_hh_f, const float* __restrict__ w_hh_r,
              const int* __restrict__ lens, int B, int T, float* __restrict__ out, float* __restrict__ save_gates,
              float* __restrict__ save_c) {
  extern __shared__ __align__(16) float lsm[];
  float* Wt = lsm;                          // [k][r]   r = gate*32 + unit
  float* hbuf = Wt + LS_H * LS_R;           // [2][k][b]
  float* gsm = hbuf + 2 * LS_H * LS_BC;     // [r][b]
  float* stage = gsm + LS_R * LS_BC;        // [unit][b]: this CTA's slice of h_t before it is pushed to the cluster
  const int tid = threadIdx.x;
  const uint32_t rank = cluster_ctarank();
  const int cid = blockIdx.x / LS_CL;       // cluster index = dir * n_chunks + chunk
  const int n_chunks = (B + LS_BC - 1) / LS_BC;
  const int dir = cid / n_chunks, chunk = cid - dir * n_chunks;
  const int b0 = chunk * LS_BC;
  const float* whh = dir ? w_hh_r : w_hh_f;

  // W_hh slice, transposed so that the 128 gate rows are contiguous (conflict-free reads across a warp)
  for (int i = tid; i < LS_R * LS_H; i += LS_THREADS) {
    const int r = i / LS_H, k = i - r * LS_H;
    const int gate = r / LS_U, u = r - gate * LS_U;
    Wt[k * LS_R + r] = __ldg(whh + static_cast<long long>(gate * LS_H + rank * LS_U + u) * LS_H + k);
  }
  for (int i = tid; i < 2 * LS_H * LS_BC; i += LS_THREADS) hbuf[i] = 0.f;
  __syncthreads();
  cluster_sync_all();

  int steps = 0;
  for (int b = 0; b < LS_BC; ++b)
    if (b0 + b < B) steps = max(steps, min(max(__ldg(lens + b0 + b), 0), T));

  // matvec role: gate row r_mv for 4 utterances; pointwise role: (unit u_pw, utterance b_pw)
  const int r_mv = tid & (LS_R - 1), bh = tid >> 7;
  const int u_pw = tid & (LS_U - 1), b_pw = tid >> 5;
  const int gb = b0 + b_pw;
  const int my_len = gb < B ? min(max(__ldg(lens + gb), 0), T) : 0;
  float c_state = 0.f;
  const uint32_t remote_h = map_to_cta(hbuf, tid >> 5);     // publishing role: warp w pushes to CTA w of the cluster

  auto load_gin = [&](int s, float (&g)[4]) {
    if (s < my_len) {
      const int t = dir ? (my_len - 1 - s) : s;
      const float* p = gates_in + ((static_cast<long long>(gb) * T + t) * 2 + dir) * (4 * LS_H) + rank * LS_U + u_pw;
#pragma unroll
      for (int q = 0; q < 4; ++q) g[q] = __ldg(p + q * LS_H);
    }
  };
  float gin[4] = {0.f, 0.f, 0.f, 0.f};
  load_gin(0, gin);

  for (int s = 0; s < steps; ++s) {
    const float* hp = hbuf + (s & 1) * LS_H * LS_BC;
    // recurrent matvec: acc[j] = sum_k W[r][k] * h[k][bh*4 + j]
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
    for (int k = 0; k < LS_H; ++k) {
      const float w = Wt[k * LS_R + r_mv];
      const float4 h4 = *reinterpret_cast<const float4*>(hp + k * LS_BC + bh * 4);
      a0 = fmaf(w, h4.x, a0); a1 = fmaf(w, h4.y, a1); a2 = fmaf(w, h4.z, a2); a3 = fmaf(w, h4.w, a3);
    }
    *reinterpret_cast<float4*>(gsm + r_mv * LS_BC + bh * 4) = make_float4(a0, a1, a2, a3);
    float gnext[4] = {0.f, 0.f, 0.f, 0.f};
    load_gin(s + 1, gnext);          // prefetch the next step's input projection under the barrier
    __syncthreads();
    // pointwise: torch gate order i, f, g, o
    float h_new = 0.f;
    const bool active = s < my_len;
    if (active) {
      const float gi = sigmoidf_(gsm[(0 * LS_U + u_pw) * LS_BC + b_pw] + gin[0]);
      const float gf = sigmoidf_(gsm[(1 * LS_U + u_pw) * LS_BC + b_pw] + gin[1]);
      const float gg = tanhf(gsm[(2 * LS_U + u_pw) * LS_BC + b_pw] + gin[2]);
      const float go = sigmoidf_(gsm[(3 * LS_U + u_pw) * LS_BC + b_pw] + gin[3]);
      c_state = fmaf(gf, c_state, gi * gg);
      h_new = go * tanhf(c_state);
      const int t = dir ? (my_len - 1 - s) : s;
      out[(static_cast<long long>(gb) * T + t) * (2 * LS_H) + dir * LS_H + rank * LS_U + u_pw] = h_new;
      if (save_gates) {              // training: the backward needs the gate activations and the cell state
        const long long fr = (static_cast<long long>(gb) * T + t) * 2 + dir;
        float* sg = save_gates + fr * (4 * LS_H) + rank * LS_U + u_pw;
        sg[0] = gi; sg[LS_H] = gf; sg[2 * LS_H] = gg; sg[3 * LS_H] = go;
        save_c[fr * LS_H + rank * LS_U + u_pw] = c_state;
      }
    } else if (gb < B && s < T) {
      out[(static_cast<long long>(gb) * T + s) * (2 * LS_H) + dir * LS_H + rank * LS_U + u_pw] = 0.f;   // padding frame
    }
    // publish this CTA's slice of h_t (32 units x 8 utterances = 1 KB, contiguous in every replica) to all CTAs of the
    // cluster as 16-byte st.shared::cluster (two per thread) instead of eight scalar remote stores per thread:
    // the scalar version spent most of the step in the SM-to-SM network
    stage[u_pw * LS_BC + b_pw] = h_new;            // finished utterances publish 0, nobody reads it
    __syncthreads();
    {
      const int i = (tid & 31) * 2;
      const uint32_t dst = remote_h +
                           static_cast<uint32_t>((((s + 1) & 1) * LS_H * LS_BC + rank * LS_U * LS_BC) * 4) + i * 16;
      st_cluster_v4(dst, *reinterpret_cast<const float4*>(stage + i * 4));
      st_cluster_v4(dst + 16, *reinterpret_cast<const float4*>(stage + i * 4 + 4));
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) gin[q] = gnext[q];
    cluster_sync_all();              // h_t visible everywhere; gsm reusable
  }
  // frames beyond the chunk's longest utterance
  if (gb < B)
    for (int t = steps; t < T; ++t)
      out[(static_cast<long long>(gb) * T + t) * (2 * LS_H) + dir * LS_H + rank * LS_U + u_pw] = 0.f;
  cluster_sync_all();                // no CTA exits while a peer may still write into its shared memory
}

// ---------------------------------------------------------------------------------------------------------------
// Backward through time.  Same decomposition as the forward: one cluster of 8 CTAs per (direction, 8 utterances), CTA c
// owns hidden units [32c, 32c+32) and keeps its 128 rows of W_hh in shared memory.  Step s (descending):
//   pointwise (unit, utterance):  dh = d_out[t] + dh_rec;  gate gradients from the saved activations and cell states
//   matvec:   partial[k][b] = sum over the CTA's 128 gate rows of dgate[r][b] * W_hh[r][k]   for ALL 256 units k
//   exchange: the partial of unit k goes to the CTA that owns k (st.shared::cluster), which sums the 8 partials at the
//             start of the next step -> dh_rec.  One barrier.cluster per step, partial buffers double-buffered.
// Output: the pre-activation gate gradients dG [2][B][T][1024] (zero on padding frames; the caller pre-zeroes it),
// from which dW_ih, dW_hh, the biases and dx are tensor-core GEMMs.
constexpr int LB_SMEM = (LS_R * LS_H + LS_R * LS_BC + 2 * LS_CL * LS_U * LS_BC) * 4;   // W | dgates | partials

__global__ void __cluster_dims__(LS_CL, 1, 1) __launch_bounds__(LS_THREADS, 1)
bilstm_bwd_kernel(const float* __restrict__ d_out, const float* __restrict__ gates_act, const float* __restrict__ cells,
                  const float* __restrict__ w_hh_f, const float* __restrict__ w_hh_r, const int* __restrict__ lens,
                  int B, int T, float* __restrict__ dG) {
  extern __shared__ __align__(16) float lsm[];
  float* W = lsm;                           // [r][k]   r = gate*32 + unit (rows of W_hh owned by this CTA)
  float* gsm = W + LS_R * LS_H;             // [r][b]   gate gradients of this step
  float* part = gsm + LS_R * LS_BC;         // [2][src CTA][unit][b]
  const int tid = threadIdx.x;
  const uint32_t rank = cluster_ctarank();
  const int cid = blockIdx.x / LS_CL;
  const int n_chunks = (B + LS_BC - 1) / LS_BC;
  const int dir = cid / n_chunks, chunk = cid - dir * n_chunks;
  const int b0 = chunk * LS_BC;
  const float* whh = dir ? w_hh_r : w_hh_f;
  for (int i = tid; i < LS_R * LS_H; i += LS_THREADS) {
    const int r = i / LS_H, k = i - r * LS_H;
    const int gate = r / LS_U, u = r - gate * LS_U;
    W[i] = __ldg(whh + static_cast<long long>(gate * LS_H + rank * LS_U + u) * LS_H + k);
  }
  for (int i = tid; i < 2 * LS_CL * LS_U * LS_BC; i += LS_THREADS) part[i] = 0.f;
  __syncthreads();
  cluster_sync_all();

  int steps = 0;
  for (int b = 0; b < LS_BC; ++b)
    if (b0 + b < B) steps = max(steps, min(max(__ldg(lens + b0 + b), 0), T));
  const int u_pw = tid & (LS_U - 1), b_pw = tid >> 5;
  const int gb = b0 + b_pw;
  const int my_len = gb < B ? min(max(__ldg(lens + gb), 0), T) : 0;
  const int unit = rank * LS_U + u_pw;
  // exchange role: thread tid = unit k of the whole layer; its partial goes to CTA k / 32, slot [rank][k % 32][*]
  const uint32_t dst_part = map_to_cta(part, tid / LS_U) +
                            static_cast<uint32_t>(((rank * LS_U + (tid & (LS_U - 1))) * LS_BC) * 4);
  float dc_state = 0.f;

  for (int j = 0, s = steps - 1; s >= 0; --s, ++j) {
    const float* pin = part + (j & 1) * (LS_CL * LS_U * LS_BC);
    float di = 0.f, df = 0.f, dg = 0.f, dgo = 0.f;
    if (s < my_len) {
      const int t = dir ? (my_len - 1 - s) : s;
      const long long fr = (static_cast<long long>(gb) * T + t) * 2 + dir;
      float dh = __ldg(d_out + (static_cast<long long>(gb) * T + t) * (2 * LS_H) + dir * LS_H + unit);
#pragma unroll
      for (int c = 0; c < LS_CL; ++c) dh += pin[(c * LS_U + u_pw) * LS_BC + b_pw];
      const float* ga = gates_act + fr * (4 * LS_H) + unit;
      const float gi = __ldg(ga), gf = __ldg(ga + LS_H), gg = __ldg(ga + 2 * LS_H), go = __ldg(ga + 3 * LS_H);
      const float c_t = __ldg(cells + fr * LS_H + unit);
      float c_prev = 0.f;
      if (s > 0) {
        const int tp = dir ? t + 1 : t - 1;
        c_prev = __ldg(cells + ((static_cast<long long>(gb) * T + tp) * 2 + dir) * LS_H + unit);
      }
      const float tc = tanhf(c_t);
      const float dc = fmaf(dh * go, 1.0f - tc * tc, dc_state);
      dgo = dh * tc * go * (1.0f - go);
      di = dc * gg * gi * (1.0f - gi);
      df = dc * c_prev * gf * (1.0f - gf);
      dg = dc * gi * (1.0f - gg * gg);
      dc_state = dc * gf;
      float* o = dG + ((static_cast<long long>(dir) * B + gb) * T + t) * (4 * LS_H) + unit;
      o[0] = di; o[LS_H] = df; o[2 * LS_H] = dg; o[3 * LS_H] = dgo;
    }
    gsm[(0 * LS_U + u_pw) * LS_BC + b_pw] = di;
    gsm[(1 * LS_U + u_pw) * LS_BC + b_pw] = df;
    gsm[(2 * LS_U + u_pw) * LS_BC + b_pw] = dg;
    gsm[(3 * LS_U + u_pw) * LS_BC + b_pw] = dgo;
    __syncthreads();
    if (s > 0) {
      // partial dh_rec of unit k = tid over this CTA's 128 gate rows, for the 8 utterances
      float a[LS_BC];
#pragma unroll
      for (int b = 0; b < LS_BC; ++b) a[b] = 0.f;
#pragma unroll 4
      for (int r = 0; r < LS_R; ++r) {
        const float w = W[r * LS_H + tid];
        const float4 g0 = *reinterpret_cast<const float4*>(gsm + r * LS_BC);
        const float4 g1 = *reinterpret_cast<const float4*>(gsm + r * LS_BC + 4);
        a[0] = fmaf(w, g0.x, a[0]); a[1] = fmaf(w, g0.y, a[1]); a[2] = fmaf(w, g0.z, a[2]); a[3] = fmaf(w, g0.w, a[3]);
        a[4] = fmaf(w, g1.x, a[4]); a[5] = fmaf(w, g1.y, a[5]); a[6] = fmaf(w, g1.z, a[6]); a[7] = fmaf(w, g1.w, a[7]);
      }
      const uint32_t dst = dst_part + static_cast<uint32_t>((((j + 1) & 1) * (LS_CL * LS_U * LS_BC)) * 4);
      st_cluster_v4(dst, make_float4(a[0], a[1], a[2], a[3]));
      st_cluster_v4(dst + 16, make_float4(a[4], a[5], a[6], a[7]));
    }
    cluster_sync_all();              // partials visible at their owners; gsm reusable
  }
  cluster_sync_all();
}

}  // namespace aptai

using namespace aptai;

static int bilstm_launch(const float* gates_in, const float* w_hh_fwd, const float* w_hh_rev, const int32_t* lens, int B,
                         int T, float* out, float* save_gates, float* save_c, void* stream) {
  if (int rc = check_arch()) return rc;
  APTAI_REQUIRE(gates_in && w_hh_fwd && w_hh_rev && lens && out, "bilstm: null pointer");
  APTAI_REQUIRE(B >= 1 && T >= 1, "bilstm: bad shape");
  APTAI_REQUIRE((save_gates == nullptr) == (save_c == nullptr), "bilstm: gate and cell buffers come together");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(bilstm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LS_SMEM);
    if (e != cudaSuccess) {
      set_error("bilstm: cudaFuncSetAttribute(%d bytes): %s", LS_SMEM, cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    attr_set = true;
  }
  const int n_chunks = (B + LS_BC - 1) / LS_BC;
  bilstm_kernel<<<2 * n_chunks * LS_CL, LS_THREADS, LS_SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(
      gates_in, w_hh_fwd, w_hh_rev, lens, B, T, out, save_gates, save_c);
  return after_launch("bilstm_256");
}

extern "C" int aptai_bilstm_256(const float* gates_in, const float* w_hh_fwd, const float* w_hh_rev,
                                const int32_t* lens, int B, int T, float* out, void* stream) {
  return bilstm_launch(gates_in, w_hh_fwd, w_hh_rev, lens, B, T, out, nullptr, nullptr, stream);
}

extern "C" int aptai_bilstm_256_train(const float* gates_in, const float* w_hh_fwd, const float* w_hh_rev,
                                      const int32_t* lens, int B, int T, float* out, float* gates_act, float* cells,
                                      void* stream) {
  APTAI_REQUIRE(gates_act && cells, "bilstm_train: null pointer");
  return bilstm_launch(gates_in, w_hh_fwd, w_hh_rev, lens, B, T, out, gates_act, cells, stream);
}

extern "C" int aptai_bilstm_256_bwd(const float* d_out, const float* gates_act, const float* cells,
                                    const float* w_hh_fwd, const float* w_hh_rev, const int32_t* lens, int B, int T,
                                    float* d_gates, void* stream) {
  if (int rc = check_arch()) return rc;
  APTAI_REQUIRE(d_out && gates_act && cells && w_hh_fwd && w_hh_rev && lens && d_gates, "bilstm_bwd: null pointer");
  APTAI_REQUIRE(B >= 1 && T >= 1, "bilstm_bwd: bad shape");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(bilstm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LB_SMEM);
    if (e != cudaSuccess) {
      set_error("bilstm_bwd: cudaFuncSetAttribute(%d bytes): %s", LB_SMEM, cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    attr_set = true;
  }
  const int n_chunks = (B + LS_BC - 1) / LS_BC;
  bilstm_bwd_kernel<<<2 * n_chunks * LS_CL, LS_THREADS, LB_SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(
      d_out, gates_act, cells, w_hh_fwd, w_hh_rev, lens, B, T, d_gates);
  return after_launch("bilstm_256_bwd");
}
