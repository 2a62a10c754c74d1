// Fused cross-attention block of Force_APTAI (models/modules.py:129-153 + models/force_aptai.py:118-130):
//   phn = Embedding(ids) + PE            (models/force_aptai.py:118-119, modules.py:233)
//   q = W_q frame + b_q ; k = W_k phn + b_k
//   energy = q k^T - 1000 * (1 - mask)   (unscaled, modules.py:144-146)
//   att_out = LayerNorm_256(cat[softmax(energy) k, q])
//   att = log_softmax(energy - 1000 * (1 - mask))   (mask applied a second time, force_aptai.py:128-130)
// and its backward for training (train/train_force_aptai.py).
//
// One CTA = a run of frames of one utterance; W_q and the utterance's 60 projected phoneme keys live in shared memory.
// Inside the CTA every WARP owns a frame: lane l holds channels {l, l+32, l+64, l+96} of the 128-wide vectors and
// slots {l, l+32} of the 60 phoneme slots, reductions are warp shuffles, and the frame loop has no CTA barrier (the
// first version gave a frame to the whole CTA and spent its time in a dozen __syncthreads and serial 60-slot scans per
// frame: profiles/r01_force_tail_bwd.md).  fp32 throughout: 60 keys x 128 dims is far below a tensor-core tile, and
// the alignment argmax wants fp32 energies.
#include "common.h"
#include "ptx.cuh"

namespace aptai {

constexpr int XA_D = 128;      // frame / phoneme / attention hidden dim
constexpr int XA_N = 60;       // max phoneme sequence length
constexpr int XA_F = 16;       // frames per CTA when the grid would otherwise not fill the GPU; 64 for large batches
constexpr int XA_WARPS = 16;   // frames in flight per CTA
constexpr int XA_THREADS = XA_WARPS * 32;
constexpr int XA_NG = XA_THREADS / XA_D;   // thread groups of 128 (one channel per thread) in the CTA-wide phases
constexpr int XA_LD = XA_D + 1;            // padded row pitch of W and k: row-per-lane reads are conflict-free

struct XAttnArgs {
  const float* frame;     // [B][T][128]
  const int* phn_ids;     // [B][60] (0 = padding); with phn_hidden: any non-zero = valid slot
  const float* phn_hidden; // optional [B][60][128]: phoneme embeddings already computed (CrossAttention.forward)
  const float* emb;       // [V][128]
  const float* pe;        // [60][128]
  const float* wq; const float* bq; const float* wk; const float* bk;   // [128][128], [128]
  const float* ln_w; const float* ln_b;                                 // [256]
  float* att_out;         // [B][T][256]
  float* energy;          // [B][T][60]  (after the first mask)
  float* att;             // [B][T][60]  log_softmax(energy + mask)
  int B, T;
  float eps;
  int frames;             // frames per CTA: the 60 key projections are recomputed per CTA, so large batches amortise
                          // them over 64 frames
};

// shared memory: W [128][129] | k [60][129] | mask [64] | per-warp scratch [XA_WARPS][2][128] | (backward) stash
constexpr int XA_SMEM_FWD = (XA_D * XA_LD + XA_N * XA_LD + 64 + XA_WARPS * 2 * XA_D) * 4;
constexpr int XA_STASH = 2 * XA_D + 2 * 64;      // per frame of a group: d_o | q | p | d_e
constexpr int XA_SMEM_BWD = XA_SMEM_FWD + XA_WARPS * XA_STASH * 4;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// CTA prologue shared by the forward and the backward: padding mask, k = W_k (emb + pe) + b_k for the utterance's 60
// slots into s_k, then W_q into s_w.  `s_x` is scratch of at least XA_NG * 128 floats.  Ends with a CTA barrier.
__device__ __forceinline__ void xa_project_keys(const XAttnArgs& a, int b, int tid, float* s_w, float* s_k, float* s_x,
                                                float* s_mask) {
  const int c = tid & (XA_D - 1), grp = tid >> 7;
  for (int i = tid; i < XA_D * XA_D; i += XA_THREADS) s_w[(i / XA_D) * XA_LD + (i % XA_D)] = a.wk[i];
  if (tid < 64) s_mask[tid] = (tid < XA_N && a.phn_ids[b * XA_N + tid] != 0) ? 0.f : -1000.f;
  __syncthreads();
  for (int n0 = 0; n0 < XA_N; n0 += XA_NG) {
    const int n = n0 + grp;
    if (n < XA_N) {
      const int id = a.phn_ids[b * XA_N + n];
      s_x[grp * XA_D + c] = a.phn_hidden ? a.phn_hidden[(static_cast<long long>(b) * XA_N + n) * XA_D + c]
                                         : a.emb[id * XA_D + c] + a.pe[n * XA_D + c];
    }
    __syncthreads();
    if (n < XA_N) {
      float acc = a.bk[c];
      const float* xr = s_x + grp * XA_D;
#pragma unroll 8
      for (int k = 0; k < XA_D; ++k) acc = fmaf(xr[k], s_w[c * XA_LD + k], acc);
      s_k[n * XA_LD + c] = acc;
    }
    __syncthreads();
  }
  for (int i = tid; i < XA_D * XA_D; i += XA_THREADS) s_w[(i / XA_D) * XA_LD + (i % XA_D)] = a.wq[i];
  __syncthreads();
}

// Everything the forward computes for one frame, by one warp.  Lane l: channels l + 32 j (j < 4), slots l and l + 32.
struct XaFrame {
  float q[4], o[4];        // projected query, context (softmax(energy) k)
  float e[2];              // energies after the first mask (-inf for slots >= 60)
  float p[2], p2[2];       // softmax(energy), softmax(energy + mask)
  float lse2;              // max + log-sum of (energy + mask): att = e + mask - lse2
  float mean, rstd;        // LayerNorm statistics of cat[o, q]
};

__device__ __forceinline__ void xa_frame_forward(const XAttnArgs& a, long long row, int lane, const float* s_w,
                                                 const float* s_k, const float* s_mask, float* w_a, float* w_b,
                                                 XaFrame& f) {
  // q = W_q x + b_q: x broadcast from the warp's scratch, lane reads its four rows of W_q (pitch 129: conflict-free)
#pragma unroll
  for (int j = 0; j < 4; ++j) w_a[lane + 32 * j] = a.frame[row * XA_D + lane + 32 * j];
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 4; ++j) f.q[j] = a.bq[lane + 32 * j];
#pragma unroll 4
  for (int k = 0; k < XA_D; k += 4) {
    const float4 x4 = *reinterpret_cast<const float4*>(w_a + k);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float* wr = s_w + (lane + 32 * j) * XA_LD + k;
      f.q[j] = fmaf(x4.w, wr[3], fmaf(x4.z, wr[2], fmaf(x4.y, wr[1], fmaf(x4.x, wr[0], f.q[j]))));
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) w_b[lane + 32 * j] = f.q[j];
  __syncwarp();
  // energies of slots lane and lane + 32 (first mask), two softmaxes over the 60 slots
  float m1 = -INFINITY, m2 = -INFINITY;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int n = lane + 32 * h;
    f.e[h] = -INFINITY;
    if (n < XA_N) {
      float acc = 0.f;
      const float* kr = s_k + n * XA_LD;
#pragma unroll 8
      for (int c = 0; c < XA_D; ++c) acc = fmaf(w_b[c], kr[c], acc);
      f.e[h] = acc + s_mask[n];
      m1 = fmaxf(m1, f.e[h]);
      m2 = fmaxf(m2, f.e[h] + s_mask[n]);
    }
  }
  m1 = warp_max(m1);
  m2 = warp_max(m2);
  float z1 = 0.f, z2 = 0.f;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int n = lane + 32 * h;
    f.p[h] = f.p2[h] = 0.f;
    if (n < XA_N) {
      f.p[h] = expf(f.e[h] - m1);
      f.p2[h] = expf(f.e[h] + s_mask[n] - m2);
    }
    z1 += f.p[h];
    z2 += f.p2[h];
  }
  z1 = warp_sum(z1);
  z2 = warp_sum(z2);
  f.lse2 = m2 + logf(z2);
  const float r1 = 1.0f / z1, r2 = 1.0f / z2;
  __syncwarp();                      // everybody has read x from w_a: reuse it for the probabilities
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    f.p[h] *= r1;
    f.p2[h] *= r2;
    w_a[lane + 32 * h] = f.p[h];     // slots 60..63 hold 0
  }
  __syncwarp();
  // context = sum_n p[n] k[n][:]
#pragma unroll
  for (int j = 0; j < 4; ++j) f.o[j] = 0.f;
#pragma unroll 4
  for (int n = 0; n < XA_N; ++n) {
    const float pn = w_a[n];
    const float* kr = s_k + n * XA_LD + lane;
#pragma unroll
    for (int j = 0; j < 4; ++j) f.o[j] = fmaf(pn, kr[32 * j], f.o[j]);
  }
  // LayerNorm over cat[o, q] (256 values, eight per lane), exact two-pass statistics
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) s += f.o[j] + f.q[j];
  f.mean = warp_sum(s) * (1.0f / 256);
  float v = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) v += (f.o[j] - f.mean) * (f.o[j] - f.mean) + (f.q[j] - f.mean) * (f.q[j] - f.mean);
  f.rstd = rsqrtf(warp_sum(v) * (1.0f / 256) + a.eps);
}

__global__ void __launch_bounds__(XA_THREADS, 1)
xattn_kernel(const XAttnArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* s_w = sm;
  float* s_k = s_w + XA_D * XA_LD;
  float* s_mask = s_k + XA_N * XA_LD;
  float* s_warp = s_mask + 64;
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t0 = blockIdx.x * a.frames;
  xa_project_keys(a, b, tid, s_w, s_k, s_warp, s_mask);
  float* w_a = s_warp + warp * 2 * XA_D;
  float* w_b = w_a + XA_D;
  const int t_end = min(t0 + a.frames, a.T);
  for (int t = t0 + warp; t < t_end; t += XA_WARPS) {
    const long long row = static_cast<long long>(b) * a.T + t;
    XaFrame f;
    xa_frame_forward(a, row, lane, s_w, s_k, s_mask, w_a, w_b, f);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = lane + 32 * h;
      if (n < XA_N) {
        a.energy[row * XA_N + n] = f.e[h];
        a.att[row * XA_N + n] = f.e[h] + s_mask[n] - f.lse2;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = lane + 32 * j;
      a.att_out[row * 256 + c] = fmaf((f.o[j] - f.mean) * f.rstd, a.ln_w[c], a.ln_b[c]);
      a.att_out[row * 256 + XA_D + c] = fmaf((f.q[j] - f.mean) * f.rstd, a.ln_w[XA_D + c], a.ln_b[XA_D + c]);
    }
    __syncwarp();                    // the warp's scratch is rewritten by the next frame
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Backward.  Inputs: d_att_out [B][T][256] (from the BiLSTM) and d_att [B][T][60] (from the forward-sum loss; optional).
// Outputs: d_q [B][T][128] (gradient of the projected queries; dW_q, db_q and d_frame are GEMMs on it), d_k [B][60][128]
// (gradient of the projected keys, accumulated with atomics over the CTAs of an utterance: pre-zero it), d_ln_w /
// d_ln_b [256] (accumulated).  Frames are processed in groups of XA_WARPS: each warp recomputes the forward of its frame
// and runs the LayerNorm / softmax / log-softmax backward with warp shuffles; the key gradient
// d_k = P^T dO + dE^T Q of the group is then accumulated by the whole CTA from a shared-memory stash (thread = one
// channel x 60 / XA_NG slots in registers), two CTA barriers per group.
struct XAttnBwdArgs {
  XAttnArgs f;             // forward inputs (att_out / energy / att unused)
  const float* d_att_out;  // [B][T][256]
  const float* d_att;      // [B][T][60] or null
  float* d_q;              // [B][T][128]
  float* d_k;              // [B][60][128]
  float* d_ln_w; float* d_ln_b;   // [256]
};

__global__ void __launch_bounds__(XA_THREADS, 1)
xattn_bwd_kernel(const XAttnBwdArgs g) {
  const XAttnArgs& a = g.f;
  extern __shared__ __align__(16) float sm[];
  float* s_w = sm;
  float* s_k = s_w + XA_D * XA_LD;
  float* s_mask = s_k + XA_N * XA_LD;
  float* s_warp = s_mask + 64;
  float* s_stash = s_warp + XA_WARPS * 2 * XA_D;       // [XA_WARPS][d_o 128 | q 128 | p 64 | d_e 64]
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t0 = blockIdx.x * a.frames;
  xa_project_keys(a, b, tid, s_w, s_k, s_warp, s_mask);
  float* w_a = s_warp + warp * 2 * XA_D;
  float* w_b = w_a + XA_D;
  float* st = s_stash + warp * XA_STASH;
  const int t_end = min(t0 + a.frames, a.T);
  constexpr int NK = XA_N / XA_NG;                     // key slots per thread in the CTA-wide d_k phase
  static_assert(XA_N % XA_NG == 0, "slots must split evenly over the channel groups");
  float dk[NK];
#pragma unroll
  for (int i = 0; i < NK; ++i) dk[i] = 0.f;
  float dgw[8], dgb[8];                                // LayerNorm gradients of channels lane + 32 j (o | q)
#pragma unroll
  for (int i = 0; i < 8; ++i) dgw[i] = dgb[i] = 0.f;

  for (int tg = t0; tg < t_end; tg += XA_WARPS) {      // uniform across the CTA
    const int t = tg + warp;
    if (t < t_end) {
      const long long row = static_cast<long long>(b) * a.T + t;
      XaFrame f;
      xa_frame_forward(a, row, lane, s_w, s_k, s_mask, w_a, w_b, f);
      // LayerNorm backward
      float go[4], gq[4], xo[4], xq[4], s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = lane + 32 * j;
        const float dyo = g.d_att_out[row * 256 + c], dyq = g.d_att_out[row * 256 + XA_D + c];
        xo[j] = (f.o[j] - f.mean) * f.rstd;
        xq[j] = (f.q[j] - f.mean) * f.rstd;
        dgw[j] = fmaf(dyo, xo[j], dgw[j]); dgw[4 + j] = fmaf(dyq, xq[j], dgw[4 + j]);
        dgb[j] += dyo; dgb[4 + j] += dyq;
        go[j] = dyo * a.ln_w[c];
        gq[j] = dyq * a.ln_w[XA_D + c];
        s1 += go[j] + gq[j];
        s2 += go[j] * xo[j] + gq[j] * xq[j];
      }
      const float c1 = warp_sum(s1) * (1.0f / 256), c2 = warp_sum(s2) * (1.0f / 256);
      float d_o[4], d_q[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        d_o[j] = f.rstd * (go[j] - c1 - xo[j] * c2);
        d_q[j] = f.rstd * (gq[j] - c1 - xq[j] * c2);
        st[lane + 32 * j] = d_o[j];
        st[XA_D + lane + 32 * j] = f.q[j];
      }
      __syncwarp();
      // d_p[n] = d_o . k[n];  softmax and log-softmax backward
      float d_p[2], da[2], spd = 0.f, sda = 0.f;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int n = lane + 32 * h;
        d_p[h] = da[h] = 0.f;
        if (n < XA_N) {
          const float* kr = s_k + n * XA_LD;
          float acc = 0.f;
#pragma unroll 8
          for (int c = 0; c < XA_D; ++c) acc = fmaf(st[c], kr[c], acc);
          d_p[h] = acc;
          if (g.d_att) da[h] = g.d_att[row * XA_N + n];
        }
        spd += f.p[h] * d_p[h];
        sda += da[h];
      }
      spd = warp_sum(spd);
      sda = warp_sum(sda);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float de = f.p[h] * (d_p[h] - spd) + (da[h] - f.p2[h] * sda);      // 0 for slots >= 60
        st[2 * XA_D + lane + 32 * h] = f.p[h];
        st[2 * XA_D + 64 + lane + 32 * h] = de;
      }
      __syncwarp();
      // d_q += sum_n d_e[n] k[n][:]
#pragma unroll 4
      for (int n = 0; n < XA_N; ++n) {
        const float de = st[2 * XA_D + 64 + n];
        const float* kr = s_k + n * XA_LD + lane;
#pragma unroll
        for (int j = 0; j < 4; ++j) d_q[j] = fmaf(de, kr[32 * j], d_q[j]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) g.d_q[row * XA_D + lane + 32 * j] = d_q[j];
    }
    __syncthreads();
    // d_k[n][c] += sum over the group's frames of p[n] d_o[c] + d_e[n] q[c]; thread = channel c, slots n = grp + XA_NG i
    {
      const int c = tid & (XA_D - 1), grp = tid >> 7;
      const int nf = min(XA_WARPS, t_end - tg);
      for (int fr = 0; fr < nf; ++fr) {
        const float* sf = s_stash + fr * XA_STASH;
        const float dov = sf[c], qv = sf[XA_D + c];
#pragma unroll
        for (int i = 0; i < NK; ++i) {
          const int n = grp + XA_NG * i;
          dk[i] = fmaf(sf[2 * XA_D + n], dov, fmaf(sf[2 * XA_D + 64 + n], qv, dk[i]));
        }
      }
    }
    __syncthreads();
  }
  if (t0 < a.T) {
    const int c = tid & (XA_D - 1), grp = tid >> 7;
#pragma unroll
    for (int i = 0; i < NK; ++i)
      atomicAdd(g.d_k + (static_cast<long long>(b) * XA_N + grp + XA_NG * i) * XA_D + c, dk[i]);
    // LayerNorm gradients: the CTA's warps are summed through shared memory first (one atomic per channel per CTA).
    // W_q is no longer needed after the last group barrier: [XA_WARPS][512] floats fit in its 66 KB.
    static_assert(XA_WARPS * 512 <= XA_D * XA_LD, "LayerNorm-gradient reduction does not fit in the W_q buffer");
    static_assert(XA_THREADS == 512, "the reduction below maps one thread to one of the 2 x 256 LayerNorm gradients");
    float* rw = s_w;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      rw[warp * 512 + lane + 32 * j] = dgw[j];
      rw[warp * 512 + XA_D + lane + 32 * j] = dgw[4 + j];
      rw[warp * 512 + 256 + lane + 32 * j] = dgb[j];
      rw[warp * 512 + 256 + XA_D + lane + 32 * j] = dgb[4 + j];
    }
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < XA_WARPS; ++w) s += rw[w * 512 + tid];
    if (tid < 256) atomicAdd(g.d_ln_w + tid, s);
    else atomicAdd(g.d_ln_b + tid - 256, s);
  }
}

}  // namespace aptai

using namespace aptai;

static int xa_frames_per_cta(int B, int T) {
  // 64 frames per CTA (four groups of 16) once that still gives every SM two CTAs' worth of work, else 16
  const long long ctas64 = static_cast<long long>(B) * ((T + 63) / 64);
  return ctas64 >= 2LL * num_sms() ? 64 : XA_F;
}

template <typename K>
static int xa_set_smem(K kernel, int bytes, const char* what) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    set_error("%s: cudaFuncSetAttribute(%d bytes): %s", what, bytes, cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return 0;
}

extern "C" int aptai_cross_attention(const float* frame, const int32_t* phn_ids, const float* phn_hidden,
                                     const float* emb, int vocab, const float* pe, const float* wq, const float* bq, const float* wk,
                                     const float* bk, const float* ln_w, const float* ln_b, float eps, int B, int T,
                                     float* att_out, float* energy, float* att, void* stream) {
  if (int rc = check_arch()) return rc;
  APTAI_REQUIRE(frame && phn_ids && wq && bq && wk && bk && ln_w && ln_b && att_out && energy && att,
                "cross_attention: null pointer");
  APTAI_REQUIRE(phn_hidden || (emb && pe), "cross_attention: need phn_hidden or (emb, pe)");
  APTAI_REQUIRE(B >= 1 && T >= 1 && vocab >= 1 && B <= 65535, "cross_attention: bad shape");
  XAttnArgs a;
  a.frame = frame; a.phn_ids = phn_ids; a.phn_hidden = phn_hidden; a.emb = emb; a.pe = pe; a.wq = wq; a.bq = bq; a.wk = wk; a.bk = bk;
  a.ln_w = ln_w; a.ln_b = ln_b; a.att_out = att_out; a.energy = energy; a.att = att; a.B = B; a.T = T; a.eps = eps;
  static bool attr_set = false;
  if (!attr_set) {
    if (int rc = xa_set_smem(xattn_kernel, XA_SMEM_FWD, "cross_attention")) return rc;
    attr_set = true;
  }
  a.frames = xa_frames_per_cta(B, T);
  xattn_kernel<<<dim3((T + a.frames - 1) / a.frames, B), XA_THREADS, XA_SMEM_FWD, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  return after_launch("cross_attention");
}

extern "C" int aptai_cross_attention_bwd(const float* frame, const int32_t* phn_ids, const float* phn_hidden,
                                         const float* wq, const float* bq, const float* wk, const float* bk,
                                         const float* ln_w, float eps, int B, int T, const float* d_att_out,
                                         const float* d_att, float* d_q, float* d_k, float* d_ln_w, float* d_ln_b,
                                         void* stream) {
  if (int rc = check_arch()) return rc;
  APTAI_REQUIRE(frame && phn_ids && phn_hidden && wq && bq && wk && bk && ln_w && d_att_out && d_q && d_k && d_ln_w &&
                d_ln_b, "cross_attention_bwd: null pointer");
  APTAI_REQUIRE(B >= 1 && T >= 1 && B <= 65535, "cross_attention_bwd: bad shape");
  XAttnBwdArgs g;
  XAttnArgs& a = g.f;
  a.frame = frame; a.phn_ids = phn_ids; a.phn_hidden = phn_hidden; a.emb = nullptr; a.pe = nullptr;
  a.wq = wq; a.bq = bq; a.wk = wk; a.bk = bk; a.ln_w = ln_w; a.ln_b = nullptr;
  a.att_out = nullptr; a.energy = nullptr; a.att = nullptr; a.B = B; a.T = T; a.eps = eps;
  g.d_att_out = d_att_out; g.d_att = d_att; g.d_q = d_q; g.d_k = d_k; g.d_ln_w = d_ln_w; g.d_ln_b = d_ln_b;
  static bool attr_set = false;
  if (!attr_set) {
    if (int rc = xa_set_smem(xattn_bwd_kernel, XA_SMEM_BWD, "cross_attention_bwd")) return rc;
    attr_set = true;
  }
  a.frames = xa_frames_per_cta(B, T);
  xattn_bwd_kernel<<<dim3((T + a.frames - 1) / a.frames, B), XA_THREADS, XA_SMEM_BWD, reinterpret_cast<cudaStream_t>(stream)>>>(g);
  return after_launch("cross_attention_bwd");
}
