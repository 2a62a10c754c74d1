// Fused cross-attention block of Force_APTAI (models/modules.py:129-153 + models/force_aptai.py:118-130):
//   phn = Embedding(ids) + PE            (models/force_aptai.py:118-119, modules.py:233)
//   q = W_q frame + b_q ; k = W_k phn + b_k
//   energy = q k^T - 1000 * (1 - mask)   (unscaled, modules.py:144-146)
//   att_out = LayerNorm_256(cat[softmax(energy) k, q])
//   att = log_softmax(energy - 1000 * (1 - mask))   (mask applied a second time, force_aptai.py:128-130)
// One CTA = 16 frames of one utterance; W_q and the utterance's 60 projected phoneme keys live in shared memory.
// fp32 throughout: 60 keys x 128 dims is far below a tensor-core tile, and the alignment argmax wants fp32 energies.
#include "common.h"
#include "ptx.cuh"

namespace aptai {

constexpr int XA_D = 128;      // frame / phoneme / attention hidden dim
constexpr int XA_N = 60;       // max phoneme sequence length
constexpr int XA_F = 16;       // frames per CTA when the grid would otherwise not fill the GPU; 64 for large batches
constexpr int XA_THREADS = 128;

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
  int frames;             // frames per CTA: the 60 key projections are recomputed per CTA (2/3 of a 16-frame CTA's FMAs),
                          // so large batches amortise them over 64 frames
};

// CTA prologue shared by the forward and the backward: padding mask, k = W_k (emb + pe) + b_k for the utterance's 60
// slots into s_k, then W_q into s_w (both as [128][129]).  Ends with a CTA barrier.
__device__ __forceinline__ void xa_project_keys(const XAttnArgs& a, int b, int tid, float* s_w, float* s_k, float* s_x,
                                                float* s_mask) {
  for (int i = tid; i < XA_D * XA_D; i += XA_THREADS) s_w[(i / XA_D) * (XA_D + 1) + (i % XA_D)] = a.wk[i];
  if (tid < XA_N) s_mask[tid] = a.phn_ids[b * XA_N + tid] != 0 ? 0.f : -1000.f;
  __syncthreads();
  for (int n = 0; n < XA_N; ++n) {
    const int id = a.phn_ids[b * XA_N + n];
    s_x[tid] = a.phn_hidden ? a.phn_hidden[(static_cast<long long>(b) * XA_N + n) * XA_D + tid]
                            : a.emb[id * XA_D + tid] + a.pe[n * XA_D + tid];
    __syncthreads();
    float acc = a.bk[tid];
#pragma unroll 8
    for (int k = 0; k < XA_D; ++k) acc = fmaf(s_x[k], s_w[tid * (XA_D + 1) + k], acc);
    s_k[n * (XA_D + 1) + tid] = acc;
    __syncthreads();
  }
  for (int i = tid; i < XA_D * XA_D; i += XA_THREADS) s_w[(i / XA_D) * (XA_D + 1) + (i % XA_D)] = a.wq[i];
  __syncthreads();
}

__global__ void __launch_bounds__(XA_THREADS)
xattn_kernel(const XAttnArgs a) {
  extern __shared__ float sm[];
  float* s_w = sm;                              // W_q, later reused row-wise: [128][129]
  float* s_k = s_w + XA_D * (XA_D + 1);         // k_phn [60][129]
  float* s_x = s_k + XA_N * (XA_D + 1);         // staging: phoneme embedding row / frame row [128]
  float* s_q = s_x + XA_D;                      // q [128]
  float* s_p = s_q + XA_D;                      // energies / probabilities [64]
  float* s_o = s_p + 64;                        // att_out [128]
  float* s_red = s_o + XA_D;                    // reductions [8]
  __shared__ float s_mask[XA_N];
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t0 = blockIdx.x * a.frames;
  // ---- k = W_k (emb + pe) + b_k for the utterance's 60 slots
  xa_project_keys(a, b, tid, s_w, s_k, s_x, s_mask);
  const float bq = a.bq[tid];
  for (int f = 0; f < a.frames; ++f) {
    const int t = t0 + f;
    if (t >= a.T) break;                        // uniform across the CTA
    const long long row = static_cast<long long>(b) * a.T + t;
    s_x[tid] = a.frame[row * XA_D + tid];
    __syncthreads();
    float q = bq;
#pragma unroll 8
    for (int k = 0; k < XA_D; ++k) q = fmaf(s_x[k], s_w[tid * (XA_D + 1) + k], q);
    s_q[tid] = q;
    __syncthreads();
    // energies (threads 0..59), first mask
    float e = -INFINITY;
    if (tid < XA_N) {
      float acc = 0.f;
#pragma unroll 8
      for (int c = 0; c < XA_D; ++c) acc = fmaf(s_q[c], s_k[tid * (XA_D + 1) + c], acc);
      e = acc + s_mask[tid];
      a.energy[row * XA_N + tid] = e;
      s_p[tid] = e;
    }
    __syncthreads();
    // softmax(energy) and log_softmax(energy + mask) over the 60 slots: every thread scans the 60 values
    float m1 = -INFINITY, m2 = -INFINITY;
    for (int n = 0; n < XA_N; ++n) {
      m1 = fmaxf(m1, s_p[n]);
      m2 = fmaxf(m2, s_p[n] + s_mask[n]);
    }
    float z1 = 0.f, z2 = 0.f;
    for (int n = 0; n < XA_N; ++n) {
      z1 += expf(s_p[n] - m1);
      z2 += expf(s_p[n] + s_mask[n] - m2);
    }
    if (tid < XA_N) a.att[row * XA_N + tid] = (e + s_mask[tid]) - m2 - logf(z2);
    __syncthreads();
    if (tid < XA_N) s_p[tid] = expf(e - m1) / z1;
    __syncthreads();
    float o = 0.f;
    for (int n = 0; n < XA_N; ++n) o = fmaf(s_p[n], s_k[n * (XA_D + 1) + tid], o);
    // LayerNorm over cat[att_out, q] (256 values: two per thread), exact two-pass statistics
    float s = o + q;
    for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) s_red[warp] = s;
    __syncthreads();
    const float mean = (s_red[0] + s_red[1] + s_red[2] + s_red[3]) * (1.0f / 256);
    float v = (o - mean) * (o - mean) + (q - mean) * (q - mean);
    for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) s_red[4 + warp] = v;
    __syncthreads();
    const float rstd = rsqrtf((s_red[4] + s_red[5] + s_red[6] + s_red[7]) * (1.0f / 256) + a.eps);
    a.att_out[row * 256 + tid] = fmaf((o - mean) * rstd, a.ln_w[tid], a.ln_b[tid]);
    a.att_out[row * 256 + XA_D + tid] = fmaf((q - mean) * rstd, a.ln_w[XA_D + tid], a.ln_b[XA_D + tid]);
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Backward of the block (training of Force_APTAI, train/train_force_aptai.py): same decomposition as the forward (one
// CTA = 16 frames of one utterance, W_q and the 60 projected keys in shared memory, everything the forward computed
// for a frame is recomputed from `frame`).  Inputs: d_att_out [B][T][256] (from the BiLSTM) and d_att [B][T][60] (from
// the forward-sum loss; optional).  Outputs: d_q [B][T][128] (gradient of the projected queries; dW_q, db_q and d_frame
// are GEMMs on it), d_k [B][60][128] (gradient of the projected keys, accumulated with atomics over the frame
// chunks: pre-zero it), d_ln_w / d_ln_b [256] (accumulated).
struct XAttnBwdArgs {
  XAttnArgs f;             // forward inputs (att_out / energy / att unused)
  const float* d_att_out;  // [B][T][256]
  const float* d_att;      // [B][T][60] or null
  float* d_q;              // [B][T][128]
  float* d_k;              // [B][60][128]
  float* d_ln_w; float* d_ln_b;   // [256]
};

__device__ __forceinline__ float xa_block_sum(float v, float* s_red, int lane, int warp) {
  for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  __syncthreads();                    // s_red free again
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  return s_red[0] + s_red[1] + s_red[2] + s_red[3];
}

__global__ void __launch_bounds__(XA_THREADS)
xattn_bwd_kernel(const XAttnBwdArgs g) {
  const XAttnArgs& a = g.f;
  extern __shared__ float sm[];
  float* s_w = sm;                              // W_k, then W_q: [128][129]
  float* s_k = s_w + XA_D * (XA_D + 1);         // k_phn [60][129]
  float* s_x = s_k + XA_N * (XA_D + 1);         // staging row [128]
  float* s_q = s_x + XA_D;                      // q [128]
  float* s_p = s_q + XA_D;                      // energies, then softmax(energy) [64]
  float* s_o = s_p + 64;                        // d(context) [128]
  float* s_red = s_o + XA_D;                    // reductions [8]
  float* s_de = s_red + 8;                      // d(energy) [64]
  float* s_da = s_de + 64;                      // d(att) [64]
  float* s_pd = s_da + 64;                      // p * d_p [64]
  __shared__ float s_mask[XA_N];
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t0 = blockIdx.x * a.frames;
  if (tid < 64) { s_de[tid] = 0.f; s_da[tid] = 0.f; s_pd[tid] = 0.f; }
  xa_project_keys(a, b, tid, s_w, s_k, s_x, s_mask);
  const float bq = a.bq[tid];
  const float lw_o = a.ln_w[tid], lw_q = a.ln_w[XA_D + tid];
  float dk[XA_N];
#pragma unroll
  for (int n = 0; n < XA_N; ++n) dk[n] = 0.f;
  float dgw_o = 0.f, dgw_q = 0.f, dgb_o = 0.f, dgb_q = 0.f;
  for (int f = 0; f < a.frames; ++f) {
    const int t = t0 + f;
    if (t >= a.T) break;                        // uniform across the CTA
    const long long row = static_cast<long long>(b) * a.T + t;
    __syncthreads();
    s_x[tid] = a.frame[row * XA_D + tid];
    __syncthreads();
    float q = bq;
#pragma unroll 8
    for (int k = 0; k < XA_D; ++k) q = fmaf(s_x[k], s_w[tid * (XA_D + 1) + k], q);
    s_q[tid] = q;
    __syncthreads();
    float e = -INFINITY;
    if (tid < XA_N) {
      float acc = 0.f;
#pragma unroll 8
      for (int c = 0; c < XA_D; ++c) acc = fmaf(s_q[c], s_k[tid * (XA_D + 1) + c], acc);
      e = acc + s_mask[tid];
      s_p[tid] = e;
      s_da[tid] = g.d_att ? g.d_att[row * XA_N + tid] : 0.f;
    }
    __syncthreads();
    float m1 = -INFINITY, m2 = -INFINITY;
    for (int n = 0; n < XA_N; ++n) {
      m1 = fmaxf(m1, s_p[n]);
      m2 = fmaxf(m2, s_p[n] + s_mask[n]);
    }
    float z1 = 0.f, z2 = 0.f, sda = 0.f;
    for (int n = 0; n < XA_N; ++n) {
      z1 += expf(s_p[n] - m1);
      z2 += expf(s_p[n] + s_mask[n] - m2);
      sda += s_da[n];
    }
    __syncthreads();
    float p1 = 0.f, p2 = 0.f;
    if (tid < XA_N) {
      p1 = expf(e - m1) / z1;
      p2 = expf(e + s_mask[tid] - m2) / z2;
      s_p[tid] = p1;
    }
    __syncthreads();
    float o = 0.f;
    for (int n = 0; n < XA_N; ++n) o = fmaf(s_p[n], s_k[n * (XA_D + 1) + tid], o);
    const float mean = xa_block_sum(o + q, s_red, lane, warp) * (1.0f / 256);
    const float var = xa_block_sum((o - mean) * (o - mean) + (q - mean) * (q - mean), s_red, lane, warp) * (1.0f / 256);
    const float rstd = rsqrtf(var + a.eps);
    const float xo = (o - mean) * rstd, xq = (q - mean) * rstd;
    const float dyo = g.d_att_out[row * 256 + tid], dyq = g.d_att_out[row * 256 + XA_D + tid];
    dgw_o = fmaf(dyo, xo, dgw_o); dgw_q = fmaf(dyq, xq, dgw_q);
    dgb_o += dyo; dgb_q += dyq;
    const float go = dyo * lw_o, gq = dyq * lw_q;
    const float c1 = xa_block_sum(go + gq, s_red, lane, warp) * (1.0f / 256);
    const float c2 = xa_block_sum(go * xo + gq * xq, s_red, lane, warp) * (1.0f / 256);
    const float d_o = rstd * (go - c1 - xo * c2);
    float d_q = rstd * (gq - c1 - xq * c2);
    s_o[tid] = d_o;
    __syncthreads();
    float d_p = 0.f;
    if (tid < XA_N) {
#pragma unroll 8
      for (int c = 0; c < XA_D; ++c) d_p = fmaf(s_o[c], s_k[tid * (XA_D + 1) + c], d_p);
      s_pd[tid] = p1 * d_p;
    }
    __syncthreads();
    float spd = 0.f;
    for (int n = 0; n < XA_N; ++n) spd += s_pd[n];
    if (tid < XA_N) s_de[tid] = p1 * (d_p - spd) + (s_da[tid] - p2 * sda);
    __syncthreads();
#pragma unroll
    for (int n = 0; n < XA_N; ++n) {
      const float de = s_de[n];
      d_q = fmaf(de, s_k[n * (XA_D + 1) + tid], d_q);
      dk[n] = fmaf(s_p[n], d_o, fmaf(de, q, dk[n]));
    }
    g.d_q[row * XA_D + tid] = d_q;
  }
  if (t0 < a.T) {
#pragma unroll
    for (int n = 0; n < XA_N; ++n) atomicAdd(g.d_k + (static_cast<long long>(b) * XA_N + n) * XA_D + tid, dk[n]);
    atomicAdd(g.d_ln_w + tid, dgw_o); atomicAdd(g.d_ln_w + XA_D + tid, dgw_q);
    atomicAdd(g.d_ln_b + tid, dgb_o); atomicAdd(g.d_ln_b + XA_D + tid, dgb_q);
  }
}

}  // namespace aptai

using namespace aptai;

static int xa_frames_per_cta(int B, int T) {
  // 64 frames per CTA once that still gives every SM three CTAs, else 16 (single utterances, small batches)
  const long long ctas64 = static_cast<long long>(B) * ((T + 63) / 64);
  return ctas64 >= 3LL * num_sms() ? 64 : XA_F;
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
  const size_t smem = sizeof(float) * (XA_D * (XA_D + 1) + XA_N * (XA_D + 1) + XA_D + XA_D + 64 + XA_D + 8);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(xattn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) {
      set_error("cross_attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    attr_set = true;
  }
  a.frames = xa_frames_per_cta(B, T);
  xattn_kernel<<<dim3((T + a.frames - 1) / a.frames, B), XA_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(a);
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
  const size_t smem = sizeof(float) * (XA_D * (XA_D + 1) + XA_N * (XA_D + 1) + XA_D + XA_D + 64 + XA_D + 8 + 3 * 64);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(xattn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) {
      set_error("cross_attention_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    attr_set = true;
  }
  a.frames = xa_frames_per_cta(B, T);
  xattn_bwd_kernel<<<dim3((T + a.frames - 1) / a.frames, B), XA_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(g);
  return after_launch("cross_attention_bwd");
}
