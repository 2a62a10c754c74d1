// Persistent bidirectional LSTM recurrence for Force_APTAI's RNN tail (models/modules.py:190-214: nn.LSTM(256, 256,
// bidirectional) over packed sequences), sm_100a thread-block clusters + distributed shared memory.
//
// The input projection W_ih x_t + b_ih + b_hh of every frame and both directions is one GEMM done beforehand
// (gates_in [B][T][2][4*H]).  This kernel runs the T sequential steps: one cluster of 8 CTAs per (direction, chunk of
// 8 utterances); CTA c owns hidden units [32c, 32c+32): its 4 x 32 rows of W_hh (128 x 256 fp32 = 128 KB) stay in
// shared memory for the whole sequence, h_{t-1} of the chunk (256 x 8 fp32) is replicated in every CTA and refreshed
// each step by 16-byte st.async pushes from the owners that complete_tx on the receiver's mbarrier: the data carries
// the synchronisation, there is no barrier.cluster (and no MEMBAR.ALL.GPU) inside the loop.  fp32 FMA throughout (the
// recurrence is latency-bound: T steps of a 256-deep dot product; tensor cores would not shorten the chain); the
// matvec is blocked 4 rows x 8 utterances per thread because the shared-memory return path (128 B/clk), not the FMA
// pipe, bounds it (profiles/r01_bilstm_fwd.md).
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
constexpr int LS_PART = 8 * LS_R * LS_BC;      // per-warp partial gate sums of the k-split matvec (32 KB)
constexpr int LS_SMEM = (LS_H * LS_R + 2 * LS_H * LS_BC + LS_R * LS_BC + LS_U * LS_BC + LS_PART) * 4;   // W^T | h x2 | gates | stage | partials

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// 16-byte store into a peer CTA's shared memory that signals the peer's mbarrier (complete_tx of 16 bytes): the data
// itself carries the synchronisation, so the step needs neither barrier.cluster nor a release fence (a cluster-scope
// release compiles to MEMBAR.ALL.GPU, which also waits for the step's global stores: ~5 us of a 10 us step)
__device__ __forceinline__ void st_async_v4(uint32_t addr, float4 v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(mbar) : "memory");
}
constexpr uint32_t LS_STEP_BYTES = LS_CL * LS_U * LS_BC * 4;   // what one CTA receives per step: 1 KB from each of 8

__global__ void __cluster_dims__(LS_CL, 1, 1) __launch_bounds__(LS_THREADS, 1)
bilstm_kernel(const float* __restrict__ gates_in, const float* __restrict__ w_hh_f, const float* __restrict__ w_hh_r,
              const int* __restrict__ lens, int B, int T, float* __restrict__ out, float* __restrict__ save_gates,
              float* __restrict__ save_c) {
  extern __shared__ __align__(16) float lsm[];
  float* Wt = lsm;                          // [k][r]   r = gate*32 + unit
  float* hbuf = Wt + LS_H * LS_R;           // [2][k][b]
  float* gsm = hbuf + 2 * LS_H * LS_BC;     // [b][r]  recurrent part of the gates
  float* stage = gsm + LS_R * LS_BC;        // [unit][b]: this CTA's slice of h_t before it is pushed to the cluster
  float4* psm = reinterpret_cast<float4*>(stage + LS_U * LS_BC);   // [warp][b][lane]: k-split partial sums
  __shared__ __align__(8) uint64_t hbar[2]; // hbar[p]: "all 8 slices of the h vector in hbuf[p] have landed"
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
  if (tid == 0) {
    mbar_init(&hbar[0], 1);
    mbar_init(&hbar[1], 1);
    fence_mbar_init();
    mbar_expect_tx(&hbar[0], LS_STEP_BYTES);      // first phases: h_1 lands in buffer 1, h_2 in buffer 0
    mbar_expect_tx(&hbar[1], LS_STEP_BYTES);
  }
  __syncthreads();
  cluster_sync_all();

  int steps = 0;
  for (int b = 0; b < LS_BC; ++b)
    if (b0 + b < B) steps = max(steps, min(max(__ldg(lens + b0 + b), 0), T));

  // pointwise role: (unit u_pw, utterance b_pw)
  const int lane = tid & 31, warp = tid >> 5;
  const int u_pw = tid & (LS_U - 1), b_pw = tid >> 5;
  const int gb = b0 + b_pw;
  const int my_len = gb < B ? min(max(__ldg(lens + gb), 0), T) : 0;
  float c_state = 0.f;
  const uint32_t remote_h = map_to_cta(hbuf, tid >> 5);     // publishing role: warp w pushes to CTA w of the cluster
  const uint32_t remote_bar = map_to_cta(hbar, tid >> 5);

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
    float gnext[4] = {0.f, 0.f, 0.f, 0.f};
    load_gin(s + 1, gnext);          // the next step's input projection (DRAM latency hidden behind this whole step)
    if (s > 0) {
      // h_s complete in buffer s&1 (use number (s-1)/2 or s/2-1 of that barrier); re-arm it for its next use
      mbar_wait(&hbar[s & 1], (s & 1) ? ((s - 1) >> 1) & 1 : ((s >> 1) - 1) & 1);
      __syncthreads();               // everybody is past the wait before the barrier is re-armed into its next phase
      if (tid == 0) mbar_expect_tx(&hbar[s & 1], LS_STEP_BYTES);
    }
    // recurrent matvec, k split across the 8 warps: warp w owns k in [32w, 32w+32), lane l the four gate rows
    // 4l..4l+3 (one LDS.128 of W^T) for all 8 utterances (two broadcast LDS.128 of h): 48 bytes returned from shared
    // memory per 32 FMAs.  (One row x 4 utterances per thread over all k returned 20 bytes per 4 FMAs, and the
    // 128 B/clk shared-memory return path, not the FMA pipe, set the step time: profiles/r01_bilstm_fwd.md.)
    {
      // packed fp32x2 FMAs (FFMA2): utterance pairs share one instruction, the issue slots of the 4x8 tile halve
      uint64_t acc2[4][LS_BC / 2];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int b = 0; b < LS_BC / 2; ++b) acc2[i][b] = 0ull;
      const float* wp = Wt + (warp * 32) * LS_R + lane * 4;
      const float* hq = hp + (warp * 32) * LS_BC;
#pragma unroll 4
      for (int kk = 0; kk < 32; ++kk) {
        const float4 w4 = *reinterpret_cast<const float4*>(wp + kk * LS_R);
        const float4 h0 = *reinterpret_cast<const float4*>(hq + kk * LS_BC);
        const float4 h1 = *reinterpret_cast<const float4*>(hq + kk * LS_BC + 4);
        const uint64_t wv[4] = {f32x2_pack(w4.x, w4.x), f32x2_pack(w4.y, w4.y), f32x2_pack(w4.z, w4.z),
                                f32x2_pack(w4.w, w4.w)};
        const uint64_t hv[LS_BC / 2] = {f32x2_pack(h0.x, h0.y), f32x2_pack(h0.z, h0.w), f32x2_pack(h1.x, h1.y),
                                        f32x2_pack(h1.z, h1.w)};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int b = 0; b < LS_BC / 2; ++b) acc2[i][b] = f32x2_fma(wv[i], hv[b], acc2[i][b]);
      }
      float acc[4][LS_BC];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int b = 0; b < LS_BC / 2; ++b) f32x2_unpack(acc2[i][b], acc[i][2 * b], acc[i][2 * b + 1]);
      float4* pw = psm + warp * 256;             // slot (utterance b, lane) = the lane's four rows of utterance b
#pragma unroll
      for (int b = 0; b < LS_BC; ++b) pw[b * 32 + lane] = make_float4(acc[0][b], acc[1][b], acc[2][b], acc[3][b]);
    }
    __syncthreads();
    {   // reduce the 8 partials of slot tid = (utterance, lane) into gsm[b][row]: every access of the reduction and
        // of the pointwise role (row = gate * 32 + lane) is bank-conflict free
      float4 sum = psm[tid];
#pragma unroll
      for (int w = 1; w < 8; ++w) {
        const float4 v = psm[w * 256 + tid];
        sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
      }
      *reinterpret_cast<float4*>(gsm + (tid >> 5) * LS_R + 4 * (tid & 31)) = sum;
    }
    __syncthreads();
    // pointwise: torch gate order i, f, g, o
    float h_new = 0.f;
    const bool active = s < my_len;
    if (active) {
      const float gi = sigmoidf_(gsm[b_pw * LS_R + 0 * LS_U + u_pw] + gin[0]);
      const float gf = sigmoidf_(gsm[b_pw * LS_R + 1 * LS_U + u_pw] + gin[1]);
      const float gg = tanhf(gsm[b_pw * LS_R + 2 * LS_U + u_pw] + gin[2]);
      const float go = sigmoidf_(gsm[b_pw * LS_R + 3 * LS_U + u_pw] + gin[3]);
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
    // cluster: two 16-byte st.async per thread (warp w serves CTA w), each completing 16 bytes on the receiver's barrier
    stage[u_pw * LS_BC + b_pw] = h_new;            // finished utterances publish 0, nobody reads it
    __syncthreads();
    if (s + 1 < steps) {
      const int i = (tid & 31) * 2;
      const uint32_t dst = remote_h +
                           static_cast<uint32_t>((((s + 1) & 1) * LS_H * LS_BC + rank * LS_U * LS_BC) * 4) + i * 16;
      const uint32_t bar = remote_bar + ((s + 1) & 1) * 8;
      st_async_v4(dst, *reinterpret_cast<const float4*>(stage + i * 4), bar);
      st_async_v4(dst + 16, *reinterpret_cast<const float4*>(stage + i * 4 + 4), bar);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) gin[q] = gnext[q];
    // no cluster barrier: a peer can only overwrite hbuf[s&1] (with h_{s+2}) after it has received this CTA's
    // h_{s+1}, which is sent after this step's matvec has read hbuf[s&1]
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
//   exchange: the partial of unit k goes to the CTA that owns k (16-byte st.async completing on the owner's mbarrier),
//             which sums the 8 partials at the start of the next step -> dh_rec.  Partial buffers double-buffered; no
//             barrier.cluster inside the loop.
// Output: the pre-activation gate gradients dG [2][B][T][1024] (zero on padding frames; the caller pre-zeroes it),
// from which dW_ih, dW_hh, the biases and dx are tensor-core GEMMs.
constexpr int LB_SMEM = (LS_R * LS_H + LS_R * LS_BC + 2 * LS_CL * LS_U * LS_BC + 4 * LS_H * LS_BC) * 4;   // W | dgates | received partials | row-split partials

__global__ void __cluster_dims__(LS_CL, 1, 1) __launch_bounds__(LS_THREADS, 1)
bilstm_bwd_kernel(const float* __restrict__ d_out, const float* __restrict__ gates_act, const float* __restrict__ cells,
                  const float* __restrict__ w_hh_f, const float* __restrict__ w_hh_r, const int* __restrict__ lens,
                  int B, int T, float* __restrict__ dG) {
  extern __shared__ __align__(16) float lsm[];
  float* W = lsm;                           // [r][k]   r = gate*32 + unit (rows of W_hh owned by this CTA)
  float* gsm = W + LS_R * LS_H;             // [r][b]   gate gradients of this step
  float* part = gsm + LS_R * LS_BC;         // [2][src CTA][b][unit]
  float4* psm = reinterpret_cast<float4*>(part + 2 * LS_CL * LS_U * LS_BC);   // [row group][k half][b][lane]
  __shared__ __align__(8) uint64_t pbar[2]; // pbar[p]: "all 8 partials in part[p] have landed"
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
  if (tid == 0) {
    mbar_init(&pbar[0], 1);
    mbar_init(&pbar[1], 1);
    fence_mbar_init();
    mbar_expect_tx(&pbar[0], LS_STEP_BYTES);
    mbar_expect_tx(&pbar[1], LS_STEP_BYTES);
  }
  __syncthreads();
  cluster_sync_all();

  int steps = 0;
  for (int b = 0; b < LS_BC; ++b)
    if (b0 + b < B) steps = max(steps, min(max(__ldg(lens + b0 + b), 0), T));
  const int u_pw = tid & (LS_U - 1), b_pw = tid >> 5;
  const int gb = b0 + b_pw;
  const int my_len = gb < B ? min(max(__ldg(lens + gb), 0), T) : 0;
  const int unit = rank * LS_U + u_pw;
  const int lane = tid & 31, warp = tid >> 5;
  float dc_state = 0.f;
  // saved activations of one step for this (unit, utterance): d_out, i, f, g, o, c_t, c_{t-1}; fetched one step ahead
  // so that their DRAM latency hides behind the previous step instead of heading every step
  auto load_step = [&](int s, float (&v)[7]) {
#pragma unroll
    for (int q = 0; q < 7; ++q) v[q] = 0.f;
    if (s >= 0 && s < my_len) {
      const int t = dir ? (my_len - 1 - s) : s;
      const long long fr = (static_cast<long long>(gb) * T + t) * 2 + dir;
      v[0] = __ldg(d_out + (static_cast<long long>(gb) * T + t) * (2 * LS_H) + dir * LS_H + unit);
      const float* ga = gates_act + fr * (4 * LS_H) + unit;
      v[1] = __ldg(ga); v[2] = __ldg(ga + LS_H); v[3] = __ldg(ga + 2 * LS_H); v[4] = __ldg(ga + 3 * LS_H);
      v[5] = __ldg(cells + fr * LS_H + unit);
      if (s > 0) {
        const int tp = dir ? t + 1 : t - 1;
        v[6] = __ldg(cells + ((static_cast<long long>(gb) * T + tp) * 2 + dir) * LS_H + unit);
      }
    }
  };
  float cur[7], nxt[7];
  load_step(steps - 1, cur);

  for (int j = 0, s = steps - 1; s >= 0; --s, ++j) {
    const float* pin = part + (j & 1) * (LS_CL * LS_U * LS_BC);
    load_step(s - 1, nxt);
    if (j > 0) {
      mbar_wait(&pbar[j & 1], (j & 1) ? ((j - 1) >> 1) & 1 : ((j >> 1) - 1) & 1);
      __syncthreads();               // also: the previous matvec is done with gsm
      if (tid == 0) mbar_expect_tx(&pbar[j & 1], LS_STEP_BYTES);
    }
    float di = 0.f, df = 0.f, dg = 0.f, dgo = 0.f;
    if (s < my_len) {
      const int t = dir ? (my_len - 1 - s) : s;
      float dh = cur[0];
#pragma unroll
      for (int c = 0; c < LS_CL; ++c) dh += pin[(c * LS_BC + b_pw) * LS_U + u_pw];
      const float gi = cur[1], gf = cur[2], gg = cur[3], go = cur[4], c_t = cur[5], c_prev = cur[6];
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
      // partial dh_rec over this CTA's 128 gate rows, split like the forward's matvec: warp w owns the 32 rows of
      // group w/2 and the 128 units of half w%2, lane l the units 4l..4l+3 (one LDS.128 of W) for all 8 utterances
      uint64_t acc2[4][LS_BC / 2];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int b = 0; b < LS_BC / 2; ++b) acc2[i][b] = 0ull;
      const int kb = warp & 1, rg = warp >> 1;
      const float* wp = W + (rg * 32) * LS_H + kb * 128 + lane * 4;
      const float* gp = gsm + (rg * 32) * LS_BC;
#pragma unroll 4
      for (int rr = 0; rr < 32; ++rr) {
        const float4 w4 = *reinterpret_cast<const float4*>(wp + rr * LS_H);
        const float4 g0 = *reinterpret_cast<const float4*>(gp + rr * LS_BC);
        const float4 g1 = *reinterpret_cast<const float4*>(gp + rr * LS_BC + 4);
        const uint64_t wv[4] = {f32x2_pack(w4.x, w4.x), f32x2_pack(w4.y, w4.y), f32x2_pack(w4.z, w4.z),
                                f32x2_pack(w4.w, w4.w)};
        const uint64_t gv[LS_BC / 2] = {f32x2_pack(g0.x, g0.y), f32x2_pack(g0.z, g0.w), f32x2_pack(g1.x, g1.y),
                                        f32x2_pack(g1.z, g1.w)};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int b = 0; b < LS_BC / 2; ++b) acc2[i][b] = f32x2_fma(wv[i], gv[b], acc2[i][b]);
      }
      float acc[4][LS_BC];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int b = 0; b < LS_BC / 2; ++b) f32x2_unpack(acc2[i][b], acc[i][2 * b], acc[i][2 * b + 1]);
      // slot (utterance b, lane) = the lane's four consecutive units of utterance b: the owner stores received
      // partials as [src][b][unit], which its pointwise role (unit = lane) reads without bank conflicts
      float4* pw = psm + rg * 512 + kb * 256;
#pragma unroll
      for (int b = 0; b < LS_BC; ++b) pw[b * 32 + lane] = make_float4(acc[0][b], acc[1][b], acc[2][b], acc[3][b]);
      __syncthreads();
      // sum the 4 row groups of slots tid and tid + 256 and send each 16-byte result to the CTA that owns its unit
#pragma unroll
      for (int hslot = 0; hslot < 2; ++hslot) {
        const int slot = tid + 256 * hslot;
        float4 sum = psm[slot];
#pragma unroll
        for (int g = 1; g < 4; ++g) {
          const float4 v = psm[g * 512 + slot];
          sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
        }
        const int bq = (slot >> 5) & 7;                                  // utterance
        const int k = (slot >> 8) * 128 + 4 * (slot & 31);               // first of the four units
        const uint32_t owner = static_cast<uint32_t>(k / LS_U);
        const uint32_t dst = map_to_cta(part, owner) +
                             static_cast<uint32_t>((((j + 1) & 1) * (LS_CL * LS_U * LS_BC) +
                                                    (rank * LS_BC + bq) * LS_U + (k & (LS_U - 1))) * 4);
        st_async_v4(dst, sum, map_to_cta(pbar, owner) + ((j + 1) & 1) * 8);
      }
    }
#pragma unroll
    for (int q = 0; q < 7; ++q) cur[q] = nxt[q];
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
