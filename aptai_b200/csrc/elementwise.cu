// HBM-bound row kernels of the encoder and the APTAI heads: LayerNorm, bf16 cast+halo pad, weight-norm fold,
// the two small output heads + argmax, the 51-tap low-pass FIR, and the masked MSE / CE losses.
#include "common.h"
#include "ptx.cuh"

#include <math.h>

namespace aptai {

// ---------------------------------------------------------------------------------------------- LayerNorm
// One warp per row; the row lives in registers (NV float4 per lane), exact two-pass statistics in fp32.
template <int NV, int IN_FMT>   // IN_FMT: 0 fp32, 1 bf16, 2 fp16
__global__ void __launch_bounds__(256)
layernorm_kernel(const void* __restrict__ xin, long long rows, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, float* __restrict__ out_f32,
                 __nv_bfloat16* __restrict__ out_bf16, int out16_fp16, int reverse) {
  constexpr int COLS = NV * 128;
  const int lane = threadIdx.x & 31;
  long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  if (reverse) row = rows - 1 - row;          // aptai_set_traversal: start on the rows the producer wrote last
  float4 v[NV];
  if (IN_FMT != 0) {
    const uint2* p = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(xin) + row * COLS);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const uint2 u = __ldg(p + i * 32 + lane);
      float2 a, b;
      if (IN_FMT == 1) {
        a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
        b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
      } else {
        a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
        b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
      }
      v[i] = make_float4(a.x, a.y, b.x, b.y);
    }
  } else {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(xin) + row * COLS);
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = __ldg(p + i * 32 + lane);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.0f / COLS);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q * (1.0f / COLS) + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = (i * 32 + lane) * 4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + col));
    const float4 e = __ldg(reinterpret_cast<const float4*>(beta + col));
    float4 y;
    y.x = fmaf((v[i].x - mean) * rstd, g.x, e.x);
    y.y = fmaf((v[i].y - mean) * rstd, g.y, e.y);
    y.z = fmaf((v[i].z - mean) * rstd, g.z, e.z);
    y.w = fmaf((v[i].w - mean) * rstd, g.w, e.w);
    if (out_f32) reinterpret_cast<float4*>(out_f32 + row * COLS)[i * 32 + lane] = y;
    if (out_bf16)
      reinterpret_cast<uint2*>(out_bf16 + row * COLS)[i * 32 + lane] =
          make_uint2(pack_h16(y.x, y.y, out16_fp16), pack_h16(y.z, y.w, out16_fp16));
  }
}

template <int NV>
static void launch_ln(const void* x, int x_fmt, long long rows, const float* gamma, const float* beta, float eps,
                      float* of, void* ob, int o16, cudaStream_t st) {
  const int wpb = 8;
  const unsigned grid = static_cast<unsigned>((rows + wpb - 1) / wpb);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(ob);
  const int rev = traversal_reverse();
  if (x_fmt == 1) layernorm_kernel<NV, 1><<<grid, wpb * 32, 0, st>>>(x, rows, gamma, beta, eps, of, o, o16, rev);
  else if (x_fmt == 2) layernorm_kernel<NV, 2><<<grid, wpb * 32, 0, st>>>(x, rows, gamma, beta, eps, of, o, o16, rev);
  else layernorm_kernel<NV, 0><<<grid, wpb * 32, 0, st>>>(x, rows, gamma, beta, eps, of, o, o16, rev);
}

// ---------------------------------------------------------------------------------------------- cast + halo pad
__global__ void cast_pad_kernel(const float* __restrict__ x, int rows, int cols4, int halo,
                                __nv_bfloat16* __restrict__ out, int fp16) {
  // grid.y = segment; each thread moves 4 elements; halo rows are written as zeros
  const int seg = blockIdx.y;
  const long long prow = rows + 2 * halo;
  const long long total = prow * cols4;
  uint2* o = reinterpret_cast<uint2*>(out) + seg * total;
  const float4* in = reinterpret_cast<const float4*>(x) + static_cast<long long>(seg) * rows * cols4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cols4 - halo;
    uint2 u = make_uint2(0u, 0u);
    if (r >= 0 && r < rows) {
      const float4 v = __ldg(in + r * cols4 + (i % cols4));
      u = make_uint2(pack_h16(v.x, v.y, fp16), pack_h16(v.z, v.w, fp16));
    }
    o[i] = u;
  }
}

// ---------------------------------------------------------------------------------------------- weight-norm fold
__global__ void posconv_norm_kernel(const float* __restrict__ v, int n_oc, int taps, float* __restrict__ norm) {
  // one block per tap: ||v[:, :, j]||_2 over all (o, c), fp64 accumulation, deterministic tree
  const int j = blockIdx.x;
  double s = 0;
  for (int i = threadIdx.x; i < n_oc; i += blockDim.x) {
    const double a = v[static_cast<long long>(i) * taps + j];
    s += a * a;
  }
  __shared__ double red[256];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) norm[j] = static_cast<float>(sqrt(red[0]));
}

__global__ void posconv_fold_kernel(const float* __restrict__ g, const float* __restrict__ v,
                                    const float* __restrict__ norm, int H, int cin, int taps, int cpad,
                                    __nv_bfloat16* __restrict__ w, int fp16) {
  // w[o][j][c] = g[j] * v[o][c][j] / norm[j]   (c >= cin -> 0)
  const long long total = static_cast<long long>(H) * taps * cpad;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cpad);
    const int j = static_cast<int>((i / cpad) % taps);
    const long long o = i / (static_cast<long long>(cpad) * taps);
    float val = 0.f;
    if (c < cin) val = g[j] * v[(o * cin + c) * taps + j] / norm[j];
    if (fp16) reinterpret_cast<__half*>(w)[i] = __float2half_rn(val);
    else w[i] = __float2bfloat16(val);
  }
}

// ---------------------------------------------------------------------------------------------- heads
// Both heads in one pass over h: 32 rows per block, outputs in 64 slots (head A: slots 0..15, head B: 16..63).
constexpr int HD_ROWS = 32;
constexpr int HD_KC = 32;
constexpr int HD_SLOTS = 64;
constexpr int HD_A = 16;

__device__ __forceinline__ float head_act(float x, int act) {
  if (act == 1) return tanhf(x);
  if (act == 2) return x > 0.f ? x : 0.01f * x;
  return x;
}

__global__ void __launch_bounds__(256)
heads_kernel(const float* __restrict__ h, long long rows, int H, const float* __restrict__ wa,
             const float* __restrict__ ba, int na, int act_a, float* __restrict__ out_a,
             const float* __restrict__ wb, const float* __restrict__ bb, int nb, int act_b,
             float* __restrict__ out_b, long long* __restrict__ argmax_b) {
  __shared__ float hA[HD_ROWS][HD_KC + 1];
  __shared__ float hB[HD_ROWS][HD_KC + 1];
  __shared__ float ws[HD_SLOTS][HD_KC + 1];
  __shared__ float res[HD_ROWS][HD_SLOTS + 1];
  const int tid = threadIdx.x;
  const int tx = tid & 15;        // 4 output slots each
  const int ty = tid >> 4;        // 2 rows each
  const long long row0 = static_cast<long long>(blockIdx.x) * HD_ROWS;
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  const bool is_a = tx < HD_A / 4;
  for (int k0 = 0; k0 < H; k0 += HD_KC) {
    for (int i = tid; i < HD_ROWS * HD_KC; i += 256) {
      const int r = i / HD_KC, k = i % HD_KC;
      const long long row = row0 + r;
      const float x = row < rows ? __ldg(h + row * H + k0 + k) : 0.f;
      hA[r][k] = head_act(x, act_a);
      hB[r][k] = head_act(x, act_b);
    }
    for (int i = tid; i < HD_SLOTS * HD_KC; i += 256) {
      const int s = i / HD_KC, k = i % HD_KC;
      float wv = 0.f;
      if (s < HD_A) {
        if (s < na) wv = __ldg(wa + static_cast<long long>(s) * H + k0 + k);
      } else if (s - HD_A < nb) {
        wv = __ldg(wb + static_cast<long long>(s - HD_A) * H + k0 + k);
      }
      ws[s][k] = wv;
    }
    __syncthreads();
    const float (*hs)[HD_KC + 1] = is_a ? hA : hB;
#pragma unroll 8
    for (int k = 0; k < HD_KC; ++k) {
      const float x0 = hs[ty * 2][k], x1 = hs[ty * 2 + 1][k];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float wv = ws[tx * 4 + j][k];
        acc[0][j] = fmaf(x0, wv, acc[0][j]);
        acc[1][j] = fmaf(x1, wv, acc[1][j]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int s = tx * 4 + j;
    float bv = 0.f;
    if (s < HD_A) {
      if (s < na) bv = ba[s];
    } else if (s - HD_A < nb) {
      bv = bb[s - HD_A];
    }
    res[ty * 2][s] = acc[0][j] + bv;
    res[ty * 2 + 1][s] = acc[1][j] + bv;
  }
  __syncthreads();
  for (int i = tid; i < HD_ROWS * HD_SLOTS; i += 256) {
    const int r = i / HD_SLOTS, s = i % HD_SLOTS;
    const long long row = row0 + r;
    if (row >= rows) continue;
    if (s < HD_A) {
      if (s < na) out_a[row * na + s] = res[r][s];
    } else if (s - HD_A < nb) {
      out_b[row * nb + (s - HD_A)] = res[r][s];
    }
  }
  if (argmax_b && nb > 0 && tid < HD_ROWS) {
    const long long row = row0 + tid;
    if (row < rows) {
      int best = 0;
      float bv = res[tid][HD_A];
      for (int s = 1; s < nb; ++s) {
        const float v = res[tid][HD_A + s];
        if (v > bv) {   // strict: first maximum wins, as torch.argmax
          bv = v;
          best = s;
        }
      }
      argmax_b[row] = best;
    }
  }
}

// ---------------------------------------------------------------------------------------------- low-pass FIR
__global__ void lowpass_kernel(const float* __restrict__ x, int T, int C, const double* __restrict__ taps, int ntaps,
                               float* __restrict__ y, long long total) {
  extern __shared__ double tp[];
  for (int i = threadIdx.x; i < ntaps; i += blockDim.x) tp[i] = taps[i];
  __syncthreads();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % C);
  const int t = static_cast<int>((i / C) % T);
  const long long b = i / (static_cast<long long>(C) * T);
  const float* xb = x + b * T * C + c;
  // torch Conv1d(padding='same') is a cross-correlation: y[t] = sum_j w[j] * x[t + j - (N-1)/2], zeros outside
  const int half = (ntaps - 1) / 2;
  double acc = 0.0;
  for (int j = 0; j < ntaps; ++j) {
    const int tt = t + j - half;
    if (tt >= 0 && tt < T) acc += tp[j] * static_cast<double>(xb[static_cast<long long>(tt) * C]);
  }
  y[i] = static_cast<float>(acc);
}

// ---------------------------------------------------------------------------------------------- row softmax
__global__ void softmax_rows_kernel(const float* __restrict__ x, long long rows, int V, int log_out,
                                    float* __restrict__ y) {
  // one warp per row (V is small: 46 phoneme classes or 60 phoneme slots)
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * V;
  float m = -INFINITY;
  for (int c = lane; c < V; c += 32) m = fmaxf(m, xr[c]);
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float s = 0.f;
  for (int c = lane; c < V; c += 32) s += expf(xr[c] - m);
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float ls = logf(s);
  for (int c = lane; c < V; c += 32) {
    const float d = xr[c] - m;
    y[row * V + c] = log_out ? d - ls : expf(d) / s;
  }
}

// ---------------------------------------------------------------------------------------------- masked MSE + CE
// accum: [0] sum sq err, [1] count tv, [2] sum ce, [3] count ce   (doubles)
__global__ void mse_ce_kernel(const float* __restrict__ tv_pred, const float* __restrict__ tv_tgt,
                              const float* __restrict__ logits, const long long* __restrict__ phn_tgt,
                              long long rows, int ntv, int V, double* __restrict__ accum) {
  double se = 0, ce = 0;
  long long ntvv = 0, nce = 0;
  for (long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; r < rows;
       r += static_cast<long long>(gridDim.x) * blockDim.x) {
    for (int j = 0; j < ntv; ++j) {
      const float tg = tv_tgt[r * ntv + j];
      if (tg != -100.0f) {
        const float d = tv_pred[r * ntv + j] - tg;
        se += static_cast<double>(d * d);
        ++ntvv;
      }
    }
    const long long tg = phn_tgt[r];
    if (tg != 0) {   // padding mask and ignore_index=0 coincide (models/aptai.py:73,95-100)
      const float* lg = logits + r * V;
      float m = lg[0];
      for (int c = 1; c < V; ++c) m = fmaxf(m, lg[c]);
      float s = 0.f;
      for (int c = 0; c < V; ++c) s += expf(lg[c] - m);
      ce += static_cast<double>(logf(s) + m - lg[tg]);
      ++nce;
    }
  }
  double v4[4] = {se, static_cast<double>(ntvv), ce, static_cast<double>(nce)};
  __shared__ double red[4];
  if (threadIdx.x < 4) red[threadIdx.x] = 0;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    double v = v4[k];
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&red[k], v);
  }
  __syncthreads();
  if (threadIdx.x < 4) atomicAdd(&accum[threadIdx.x], red[threadIdx.x]);
}

__global__ void mse_ce_finalize_kernel(const double* __restrict__ accum, float* __restrict__ out3) {
  const double mse = accum[0] / accum[1];   // 0/0 -> NaN exactly like F.mse_loss on an empty selection
  const double ce = accum[2] / accum[3];
  out3[1] = static_cast<float>(mse);
  out3[2] = static_cast<float>(ce);
  out3[0] = 0.5f * static_cast<float>(mse) + 0.5f * static_cast<float>(ce);
}

}  // namespace aptai

using namespace aptai;

extern "C" int aptai_layernorm(const void* x, int x_fmt, int64_t rows, int cols, const float* gamma,
                               const float* beta, float eps, float* out_f32, void* out_bf16, int out16_fp16,
                               void* stream) {
  if (int rc = check_arch()) return rc;
  APTAI_REQUIRE(x && gamma && beta && (out_f32 || out_bf16), "layernorm: null pointer");
  APTAI_REQUIRE(rows >= 1, "layernorm: rows=%lld", (long long)rows);
  APTAI_REQUIRE(x_fmt >= 0 && x_fmt <= 2, "layernorm: x_fmt must be 0 (f32), 1 (bf16) or 2 (fp16)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (cols) {
    case 128: launch_ln<1>(x, x_fmt, rows, gamma, beta, eps, out_f32, out_bf16, out16_fp16, st); break;
    case 256: launch_ln<2>(x, x_fmt, rows, gamma, beta, eps, out_f32, out_bf16, out16_fp16, st); break;
    case 384: launch_ln<3>(x, x_fmt, rows, gamma, beta, eps, out_f32, out_bf16, out16_fp16, st); break;
    case 512: launch_ln<4>(x, x_fmt, rows, gamma, beta, eps, out_f32, out_bf16, out16_fp16, st); break;
    case 768: launch_ln<6>(x, x_fmt, rows, gamma, beta, eps, out_f32, out_bf16, out16_fp16, st); break;
    case 1024: launch_ln<8>(x, x_fmt, rows, gamma, beta, eps, out_f32, out_bf16, out16_fp16, st); break;
    case 1280: launch_ln<10>(x, x_fmt, rows, gamma, beta, eps, out_f32, out_bf16, out16_fp16, st); break;
    default:
      set_error("layernorm: unsupported width %d (128, 256, 384, 512, 768, 1024, 1280)", cols);
      return APTAI_ERR_ARG;
  }
  return after_launch("layernorm");
}

extern "C" int aptai_cast_pad_h16(const float* x, int segs, int rows, int cols, int halo, void* out_bf16,
                                  int half_fmt, void* stream) {
  if (int rc = check_arch()) return rc;
  APTAI_REQUIRE(x && out_bf16, "cast_pad: null pointer");
  APTAI_REQUIRE(segs >= 1 && rows >= 1 && cols % 4 == 0 && halo >= 0, "cast_pad: bad shape");
  const long long total = static_cast<long long>(rows + 2 * halo) * (cols / 4);
  int gx = static_cast<int>((total + 255) / 256);
  if (gx > 4096) gx = 4096;
  cast_pad_kernel<<<dim3(gx, segs), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, rows, cols / 4, halo, reinterpret_cast<__nv_bfloat16*>(out_bf16), half_fmt ? 1 : 0);
  return after_launch("cast_pad_bf16");
}

extern "C" int aptai_cast_pad_bf16(const float* x, int segs, int rows, int cols, int halo, void* out_bf16,
                                   void* stream) {
  return aptai_cast_pad_h16(x, segs, rows, cols, halo, out_bf16, 0, stream);
}

extern "C" int aptai_posconv_fold_fmt(const float* g, const float* v, int H, int cin, int taps, int cpad, void* w_bf16,
                                      float* norm_ws, int half_fmt, void* stream) {
  if (int rc = check_arch()) return rc;
  APTAI_REQUIRE(g && v && w_bf16 && norm_ws, "posconv_fold: null pointer");
  APTAI_REQUIRE(cpad >= cin && H >= 1 && taps >= 1, "posconv_fold: bad shape");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  posconv_norm_kernel<<<taps, 256, 0, st>>>(v, H * cin, taps, norm_ws);
  if (int rc = after_launch("posconv_norm")) return rc;
  const long long total = static_cast<long long>(H) * taps * cpad;
  int gx = static_cast<int>((total + 255) / 256);
  if (gx > 8192) gx = 8192;
  posconv_fold_kernel<<<gx, 256, 0, st>>>(g, v, norm_ws, H, cin, taps, cpad, reinterpret_cast<__nv_bfloat16*>(w_bf16),
                                          half_fmt ? 1 : 0);
  return after_launch("posconv_fold");
}

extern "C" int aptai_posconv_fold(const float* g, const float* v, int H, int cin, int taps, int cpad, void* w_bf16,
                                  float* norm_ws, void* stream) {
  return aptai_posconv_fold_fmt(g, v, H, cin, taps, cpad, w_bf16, norm_ws, 0, stream);
}

extern "C" int aptai_heads(const float* h, int64_t rows, int H, const float* wa, const float* ba, int na, int act_a,
                           float* out_a, const float* wb, const float* bb, int nb, int act_b, float* out_b,
                           int64_t* argmax_b, void* stream) {
  if (int rc = check_arch()) return rc;
  APTAI_REQUIRE(h && rows >= 1 && H % HD_KC == 0, "heads: bad input (H must be a multiple of %d)", HD_KC);
  APTAI_REQUIRE(na >= 0 && na <= HD_A && nb >= 0 && nb <= HD_SLOTS - HD_A, "heads: na <= %d and nb <= %d", HD_A,
                HD_SLOTS - HD_A);
  APTAI_REQUIRE(na == 0 || (wa && ba && out_a), "heads: head A pointers");
  APTAI_REQUIRE(nb == 0 || (wb && bb && out_b), "heads: head B pointers");
  const unsigned grid = static_cast<unsigned>((rows + HD_ROWS - 1) / HD_ROWS);
  heads_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      h, rows, H, wa, ba, na, act_a, out_a, wb, bb, nb, act_b, out_b, reinterpret_cast<long long*>(argmax_b));
  return after_launch("heads");
}

namespace aptai {
// ---------------------------------------------------------------------------------------------------------------
// Tail of the APTAI / Wav2Vec2_PR forward in ONE kernel: (final LayerNorm of the pre-LN encoder: HF:792) -> both heads
// (models/aptai.py:43-55: tv = W_tv tanh(h) + b, logits = W_phn leaky_relu(h) + b; w2v2_pr.py:58) -> argmax (first
// maximum, aptai.py:105-106) -> log_softmax of the phoneme logits (the alignment stage's input).
// CTA = 128 rows x 64 output slots (16 for head A, 48 for head B), 256 threads, thread tile 4 rows x 8 slots: per k
// three 16-byte shared loads feed 32 FMAs (the first version's 2 x 8 tile fed 16 FMAs from three loads and spent a third
// of its instructions in tanhf); the loop is bound by shared-memory wavefronts, see the lane mapping below.
// The next k-chunk's rows and weights are fetched into registers BEFORE the current chunk is multiplied, so their
// latency hides behind 1024 FMAs per thread; tanh = 1 - 2 / (exp(2|x|) + 1) on the MUFU (ex2 + rcp: ~1e-7 absolute,
// far inside the 2e-5 the heads are held to).  Pass 0: warp per row, exact two-pass LayerNorm statistics in registers;
// the main loop re-reads the rows (L2 hits: 512 KB per CTA) and normalises + activates them while staging the k-chunk.
// fp32 FMA throughout: N = 9 + 46 is too small for a tensor-core tile to pay, and the argmax parity needs fp32 logits.
constexpr int TL_ROWS = 128, TL_KC = 32, TL_SLOTS = 64, TL_A = 16;
constexpr int TL_RP = TL_ROWS + 4;     // row pitch of the staged activations (floats): 16-byte aligned k rows
constexpr int TL_WP = TL_SLOTS + 4;

__device__ __forceinline__ float tail_tanh(float x) {
  const float ax = fminf(fabsf(x), 15.f);                 // tanh(15) == 1 in fp32; keeps exp finite
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(ax * 2.885390081777927f));      // exp(2|x|)
  const float t = 1.f - __fdividef(2.f, e + 1.f);
  return copysignf(t, x);
}
__device__ __forceinline__ float tail_act(float x, int act) {
  if (act == 1) return tail_tanh(x);
  if (act == 2) return x > 0.f ? x : 0.01f * x;
  return x;
}

__global__ void __launch_bounds__(256)
tail_kernel(const float* __restrict__ h, long long rows, int H, const float* __restrict__ gamma,
            const float* __restrict__ beta, float eps, const float* __restrict__ wa, const float* __restrict__ ba,
            int na, int act_a, float* __restrict__ out_a, const float* __restrict__ wb, const float* __restrict__ bb,
            int nb, int act_b, float* __restrict__ out_b, long long* __restrict__ argmax_b, float* __restrict__ logp_b,
            float* __restrict__ h_norm) {
  // [k][row] activations of both heads (a thread's four rows are one 16-byte load), [k][slot] weights (a thread's
  // eight slots are two 16-byte loads); the result tile re-uses the activation buffers after the main loop
  __shared__ __align__(16) float hbuf[2 * TL_KC * TL_RP];
  __shared__ __align__(16) float ws[TL_KC][TL_WP];
  __shared__ float2 stat[TL_ROWS];                         // {mean, rstd}
  float (*hA)[TL_RP] = reinterpret_cast<float (*)[TL_RP]>(hbuf);
  float (*hB)[TL_RP] = reinterpret_cast<float (*)[TL_RP]>(hbuf + TL_KC * TL_RP);
  float (*res)[TL_SLOTS + 1] = reinterpret_cast<float (*)[TL_SLOTS + 1]>(hbuf);
  static_assert(TL_ROWS * (TL_SLOTS + 1) <= 2 * TL_KC * TL_RP, "result tile must fit the activation buffers");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long row0 = static_cast<long long>(blockIdx.x) * TL_ROWS;
  const bool ln = gamma != nullptr;
  // ---- pass 0: LayerNorm statistics, warp per row (two-pass from registers: H <= 1024 -> 32 values per lane)
  if (ln) {
    for (int r = warp; r < TL_ROWS; r += 8) {
      const long long row = row0 + r;
      float v[32];
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = (i * 32 + lane) * 4;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < rows && k < H) x = __ldg(reinterpret_cast<const float4*>(h + row * H + k));
        v[4 * i] = x.x; v[4 * i + 1] = x.y; v[4 * i + 2] = x.z; v[4 * i + 3] = x.w;
        s += (x.x + x.y) + (x.z + x.w);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s / H;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = (i * 32 + lane) * 4;
        if (k < H) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float d = v[4 * i + j] - mean;
            q = fmaf(d, d, q);
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      if (lane == 0) stat[r] = make_float2(mean, rsqrtf(q / H + eps));
    }
  } else if (tid < TL_ROWS) {
    stat[tid] = make_float2(0.f, 1.f);
  }
  __syncthreads();
  // A 16-byte shared load is served one quarter-warp (8 lanes) at a time: the 8 lanes of a quarter own 8 DIFFERENT
  // row groups (their x loads cover 128 contiguous bytes: one wavefront) and the SAME slot group (their weight loads
  // are one broadcast wavefront each) — 12 wavefronts per 32 FMAs; with tid = ty * 8 + tx the weight loads of a
  // quarter hit every bank twice and a k-step cost 20 (measured: 6.5 wavefronts per load, LSU pipe 73 % busy)
  const int tx = (lane >> 3) + 4 * (warp & 1);        // slots tx*8 .. tx*8+7 (tx < 2: head A)
  const int ty = (warp >> 1) * 8 + (lane & 7);        // rows ty*4 .. ty*4+3
  const bool is_a = tx < TL_A / 8;
  // staging role: a warp covers 32 consecutive rows at one k4 (conflict-free transposed stores); four h quads and two
  // weight quads per thread and chunk
  const int sr = tid & (TL_ROWS - 1), sk = (tid >> 7) * 4;             // h: rows sr, k4 = sk + 8 * m
  const int wsl = tid & (TL_SLOTS - 1), wk = (tid >> 6) * 4;           // w: slot wsl, k4 = wk + 16 * m
  const long long srow = row0 + sr;
  const float* wrow = nullptr;
  if (wsl < TL_A) {
    if (wsl < na) wrow = wa + static_cast<long long>(wsl) * H;
  } else if (wsl - TL_A < nb) {
    wrow = wb + static_cast<long long>(wsl - TL_A) * H;
  }
  float4 ph[4], pw[2];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      ph[m] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (srow < rows) ph[m] = __ldg(reinterpret_cast<const float4*>(h + srow * H + k0 + sk + 8 * m));
    }
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      pw[m] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (wrow != nullptr) pw[m] = __ldg(reinterpret_cast<const float4*>(wrow + k0 + wk + 16 * m));
    }
  };
  auto stage = [&](int k0) {
    const float2 st = stat[sr];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int k4 = sk + 8 * m;
      float xv[4] = {ph[m].x, ph[m].y, ph[m].z, ph[m].w};
      if (ln) {
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + k0 + k4));
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + k0 + k4));
        const float gv[4] = {g4.x, g4.y, g4.z, g4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) xv[j] = fmaf((xv[j] - st.x) * st.y, gv[j], bv[j]);
        if (h_norm != nullptr && srow < rows)
          *reinterpret_cast<float4*>(h_norm + srow * H + k0 + k4) = make_float4(xv[0], xv[1], xv[2], xv[3]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        hA[k4 + j][sr] = tail_act(xv[j], act_a);
        hB[k4 + j][sr] = tail_act(xv[j], act_b);
      }
    }
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      const int k4 = wk + 16 * m;
      ws[k4][wsl] = pw[m].x; ws[k4 + 1][wsl] = pw[m].y; ws[k4 + 2][wsl] = pw[m].z; ws[k4 + 3][wsl] = pw[m].w;
    }
  };
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  fetch(0);
  for (int k0 = 0; k0 < H; k0 += TL_KC) {
    stage(k0);
    __syncthreads();
    if (k0 + TL_KC < H) fetch(k0 + TL_KC);         // in flight under the 1024 FMAs below
    const float (*hs)[TL_RP] = is_a ? hA : hB;
#pragma unroll 8
    for (int k = 0; k < TL_KC; ++k) {
      const float4 x = *reinterpret_cast<const float4*>(&hs[k][ty * 4]);
      const float4 w0 = *reinterpret_cast<const float4*>(&ws[k][tx * 8]);
      const float4 w1 = *reinterpret_cast<const float4*>(&ws[k][tx * 8 + 4]);
      const float xr[4] = {x.x, x.y, x.z, x.w};
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(xr[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int sl = tx * 8 + j;
    float bv = 0.f;
    if (sl < TL_A) {
      if (sl < na) bv = ba[sl];
    } else if (sl - TL_A < nb) {
      bv = bb[sl - TL_A];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) res[ty * 4 + i][sl] = acc[i][j] + bv;
  }
  __syncthreads();
  for (int i = tid; i < TL_ROWS * TL_SLOTS; i += 256) {
    const int r = i / TL_SLOTS, sl = i % TL_SLOTS;
    const long long row = row0 + r;
    if (row >= rows) continue;
    if (sl < TL_A) {
      if (sl < na) out_a[row * na + sl] = res[r][sl];
    } else if (sl - TL_A < nb) {
      out_b[row * nb + (sl - TL_A)] = res[r][sl];
    }
  }
  // ---- argmax (first maximum) and log-softmax of head B: warp per row, lane = slots lane and lane + 32
  if (nb > 0 && (argmax_b != nullptr || logp_b != nullptr)) {
    for (int r = warp; r < TL_ROWS; r += 8) {
      const long long row = row0 + r;
      if (row >= rows) break;
      const float v0 = lane < nb ? res[r][TL_A + lane] : -INFINITY;
      const float v1 = lane + 32 < nb ? res[r][TL_A + lane + 32] : -INFINITY;
      float bv = v0;
      int bi = lane;
      if (v1 > bv) {
        bv = v1;
        bi = lane + 32;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) {
          bv = ov;
          bi = oi;
        }
      }
      if (argmax_b != nullptr && lane == 0) argmax_b[row] = bi;
      if (logp_b != nullptr) {
        float e = (lane < nb ? expf(v0 - bv) : 0.f) + (lane + 32 < nb ? expf(v1 - bv) : 0.f);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
        const float lse = bv + logf(e);
        if (lane < nb) logp_b[row * nb + lane] = v0 - lse;
        if (lane + 32 < nb) logp_b[row * nb + lane + 32] = v1 - lse;
      }
    }
  }
}
}  // namespace aptai

extern "C" int aptai_tail(const float* h, int64_t rows, int H, const float* ln_gamma, const float* ln_beta, float eps,
                          const float* wa, const float* ba, int na, int act_a, float* out_a, const float* wb,
                          const float* bb, int nb, int act_b, float* out_b, int64_t* argmax_b, float* logp_b,
                          float* h_norm, void* stream) {
  if (int rc = check_arch()) return rc;
  APTAI_REQUIRE(h && rows >= 1, "tail: null input");
  APTAI_REQUIRE(H >= TL_KC && H % TL_KC == 0 && H <= 1024, "tail: H must be a multiple of %d, at most 1024", TL_KC);
  APTAI_REQUIRE((ln_gamma == nullptr) == (ln_beta == nullptr), "tail: LayerNorm needs gamma and beta");
  APTAI_REQUIRE(na >= 0 && na <= TL_A && nb >= 0 && nb <= TL_SLOTS - TL_A, "tail: na <= %d and nb <= %d", TL_A,
                TL_SLOTS - TL_A);
  APTAI_REQUIRE(na == 0 || (wa && ba && out_a), "tail: head A pointers");
  APTAI_REQUIRE(nb == 0 || (wb && bb && out_b), "tail: head B pointers");
  APTAI_REQUIRE(h_norm == nullptr || ln_gamma != nullptr, "tail: h_norm is the LayerNorm output");
  const unsigned grid = static_cast<unsigned>((rows + TL_ROWS - 1) / TL_ROWS);
  tail_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      h, rows, H, ln_gamma, ln_beta, eps, wa, ba, na, act_a, out_a, wb, bb, nb, act_b, out_b,
      reinterpret_cast<long long*>(argmax_b), logp_b, h_norm);
  return after_launch("tail");
}

extern "C" int aptai_lowpass_fir(const float* x, int B, int T, int C, const double* taps, int ntaps, float* y,
                                 void* stream) {
  if (int rc = check_arch()) return rc;
  APTAI_REQUIRE(x && y && taps, "lowpass: null pointer");
  APTAI_REQUIRE(B >= 1 && T >= 1 && C >= 1 && ntaps >= 1 && (ntaps & 1), "lowpass: bad shape (ntaps must be odd)");
  const long long total = static_cast<long long>(B) * T * C;
  lowpass_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, ntaps * sizeof(double),
                   reinterpret_cast<cudaStream_t>(stream)>>>(x, T, C, taps, ntaps, y, total);
  return after_launch("lowpass_fir");
}

namespace aptai {
// frames per utterance after the conv feature encoder: n <- floor((n - k) / s) + 1 per layer (HF:1005-1024
// `_get_feat_extract_output_lengths`; floor division as torch.div(..., rounding_mode="floor"), so a too-short input goes
// negative exactly like the reference).  One launch instead of 3 ATen kernels per conv layer.
struct ConvGeom {
  int n, k[8], s[8];
};
__global__ void frame_lengths_kernel(const long long* __restrict__ samples, int B, ConvGeom g,
                                     long long* __restrict__ out64, int* __restrict__ out32) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  long long n = samples[b];
  for (int i = 0; i < g.n; ++i) {
    const long long d = n - g.k[i];
    long long q = d / g.s[i];
    if ((d % g.s[i] != 0) && ((d < 0) != (g.s[i] < 0))) --q;       // floor
    n = q + 1;
  }
  if (out64) out64[b] = n;
  if (out32) out32[b] = static_cast<int>(n);
}
}  // namespace aptai

extern "C" int aptai_frame_lengths(const int64_t* samples, int B, const int32_t* kernels, const int32_t* strides,
                                   int n_layers, int64_t* out_i64, int32_t* out_i32, void* stream) {
  if (int rc = check_arch()) return rc;
  APTAI_REQUIRE(samples && kernels && strides && (out_i64 || out_i32), "frame_lengths: null pointer");
  APTAI_REQUIRE(B >= 1 && n_layers >= 1 && n_layers <= 8, "frame_lengths: B >= 1 and 1..8 conv layers");
  ConvGeom g;
  g.n = n_layers;
  for (int i = 0; i < n_layers; ++i) {
    APTAI_REQUIRE(strides[i] >= 1 && kernels[i] >= 1, "frame_lengths: bad conv geometry");
    g.k[i] = kernels[i];
    g.s[i] = strides[i];
  }
  frame_lengths_kernel<<<(B + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(samples), B, g, reinterpret_cast<long long*>(out_i64), out_i32);
  return after_launch("frame_lengths");
}

extern "C" int aptai_softmax_rows(const float* x, int64_t rows, int V, int log_out, float* y, void* stream) {
  if (int rc = check_arch()) return rc;
  APTAI_REQUIRE(x && y && rows >= 1 && V >= 1, "softmax_rows: bad arguments");
  softmax_rows_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, rows, V, log_out, y);
  return after_launch("softmax_rows");
}

extern "C" int aptai_masked_mse_ce(const float* tv_pred, const float* tv_tgt, const float* logits,
                                   const int64_t* phn_tgt, int64_t rows, int ntv, int V, float* accum_ws,
                                   float* out3, void* stream) {
  if (int rc = check_arch()) return rc;
  APTAI_REQUIRE(tv_pred && tv_tgt && logits && phn_tgt && accum_ws && out3, "mse_ce: null pointer");
  APTAI_REQUIRE((reinterpret_cast<uintptr_t>(accum_ws) & 7) == 0, "mse_ce: accum_ws must be 8-byte aligned (4 doubles)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  double* acc = reinterpret_cast<double*>(accum_ws);
  cudaError_t e = cudaMemsetAsync(acc, 0, 4 * sizeof(double), st);
  if (e != cudaSuccess) {
    set_error("mse_ce: memset: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  int grid = static_cast<int>((rows + 127) / 128);
  if (grid > 1024) grid = 1024;
  mse_ce_kernel<<<grid, 128, 0, st>>>(tv_pred, tv_tgt, logits, reinterpret_cast<const long long*>(phn_tgt), rows,
                                       ntv, V, acc);
  if (int rc = after_launch("mse_ce")) return rc;
  mse_ce_finalize_kernel<<<1, 1, 0, st>>>(acc, out3);
  return after_launch("mse_ce_finalize");
}
