// HBM-bound kernels of the training step: bias gradients (column sums), LayerNorm backward, the backward of the two
// small output heads, masked MSE/CE backward, weight-norm backward of the positional conv, the attention
// backward's row-dot preprocessing, and the fused Adam update.  The tensor-core parts of the backward pass live
// in gemm_tc.cu (dgrad = the forward GEMM on transposed weights), gemm_tn.cu (wgrad) and attention_bwd.cu.
//
// What each kernel differentiates is cited at its entry point ("HF:n" = transformers 5.5.0 modeling_wav2vec2.py).
#include "common.h"
#include "ptx.cuh"

#include <math.h>

namespace aptai {

// ---------------------------------------------------------------------------------------------- column sums
// out[n] += scale * sum_m x[m][n].  Block = 8 warps x a slab of rows; a lane owns 8 (bf16) / 4 (fp32) adjacent columns
// and reads them with one 16-byte load per row, four rows in flight; the warps' partial sums meet in shared memory
// and leave as one atomicAdd per column and block.
template <bool BF16IN>
__global__ void __launch_bounds__(256)
colsum_kernel(const void* __restrict__ xin, long long M, int N, long long ld, int rows_per_block, float scale,
              float* __restrict__ out) {
  constexpr int CPL = BF16IN ? 8 : 4;            // columns per lane
  constexpr int CPB = 32 * CPL;                  // columns per block
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * CPB + lane * CPL;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_block;
  const long long r1 = min(M, r0 + rows_per_block);
  float acc[CPL];
#pragma unroll
  for (int i = 0; i < CPL; ++i) acc[i] = 0.f;
  auto add = [&](const uint4& u) {
    if (BF16IN) {
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
        acc[2 * j] += f.x;
        acc[2 * j + 1] += f.y;
      }
    } else {
      acc[0] += __uint_as_float(u.x); acc[1] += __uint_as_float(u.y);
      acc[2] += __uint_as_float(u.z); acc[3] += __uint_as_float(u.w);
    }
  };
  if (col + CPL <= N) {
    const char* base = reinterpret_cast<const char*>(xin) + static_cast<long long>(col) * (BF16IN ? 2 : 4);
    const long long pitch = ld * (BF16IN ? 2 : 4);
    long long r = r0 + warp;
    for (; r + 56 < r1; r += 64) {               // eight rows of this warp in flight together (4 KB per warp)
      uint4 u[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) u[j] = __ldg(reinterpret_cast<const uint4*>(base + (r + 8 * j) * pitch));
#pragma unroll
      for (int j = 0; j < 8; ++j) add(u[j]);
    }
    for (; r + 24 < r1; r += 32) {               // rows r, r+8, r+16, r+24 of this warp in flight together
      const uint4 u0 = __ldg(reinterpret_cast<const uint4*>(base + r * pitch));
      const uint4 u1 = __ldg(reinterpret_cast<const uint4*>(base + (r + 8) * pitch));
      const uint4 u2 = __ldg(reinterpret_cast<const uint4*>(base + (r + 16) * pitch));
      const uint4 u3 = __ldg(reinterpret_cast<const uint4*>(base + (r + 24) * pitch));
      add(u0); add(u1); add(u2); add(u3);
    }
    for (; r < r1; r += 8) add(__ldg(reinterpret_cast<const uint4*>(base + r * pitch)));
  } else if (col < N) {                          // ragged right edge: scalar
    for (long long r = r0 + warp; r < r1; r += 8)
      for (int i = 0; i < CPL && col + i < N; ++i)
        acc[i] += BF16IN ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(xin)[r * ld + col + i])
                         : reinterpret_cast<const float*>(xin)[r * ld + col + i];
  }
  __shared__ float red[8][CPB];
#pragma unroll
  for (int i = 0; i < CPL; ++i) red[warp][lane * CPL + i] = acc[i];
  __syncthreads();
  for (int c = threadIdx.x; c < CPB; c += 256) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][c];
    const int gc = blockIdx.x * CPB + c;
    if (gc < N) atomicAdd(out + gc, s * scale);
  }
}

// ---------------------------------------------------------------------------------------------- LayerNorm backward
// y = (x - mean) * rstd * gamma + beta.  Given dy:  dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy*gamma;
// dgamma += sum_rows dy * xhat, dbeta += sum_rows dy.  One warp per row (row in registers, statistics recomputed from
// the saved input exactly as the forward kernel computes them); dgamma/dbeta accumulate in registers over the rows
// a warp visits, are combined across the block's warps in shared memory and leave as atomics.
// `dres` (optional) is the gradient arriving over the residual connection: dx_out = dres + dx.
// CS: also accumulate dcol[c] += sum over rows of dx_out[row][c] — the bias gradient of the Linear whose output the
// normalised tensor's INPUT is (out-proj / FFN2 write the residual stream this LayerNorm reads), which would
// otherwise be a colsum launch re-reading dx_out.
template <int NV, bool CS>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, long long rows,
                     const float* __restrict__ gamma, float eps, const float* __restrict__ dres,
                     float* __restrict__ dx_f32, __nv_bfloat16* __restrict__ dx_bf16, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, float* __restrict__ dcol) {
  constexpr int COLS = NV * 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 ag[NV], ab[NV], ac[CS ? NV : 1];
#pragma unroll
  for (int i = 0; i < NV; ++i) ag[i] = ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < (CS ? NV : 1); ++i) ac[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* gmp = reinterpret_cast<const float4*>(gamma);
  for (long long row = static_cast<long long>(blockIdx.x) * 8 + warp; row < rows;
       row += static_cast<long long>(gridDim.x) * 8) {
    float4 v[NV], d[NV], rs[NV];
    const float4* xp = reinterpret_cast<const float4*>(x + row * COLS);
    const float4* dp = reinterpret_cast<const float4*>(dy + row * COLS);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i] = __ldg(xp + i * 32 + lane);
      d[i] = __ldg(dp + i * 32 + lane);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i)
      rs[i] = dres ? __ldg(reinterpret_cast<const float4*>(dres + row * COLS) + i * 32 + lane)
                   : make_float4(0.f, 0.f, 0.f, 0.f);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / COLS);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.0f / COLS) + eps);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;      // xhat
      ag[i].x = fmaf(d[i].x, v[i].x, ag[i].x); ag[i].y = fmaf(d[i].y, v[i].y, ag[i].y);
      ag[i].z = fmaf(d[i].z, v[i].z, ag[i].z); ag[i].w = fmaf(d[i].w, v[i].w, ag[i].w);
      ab[i].x += d[i].x; ab[i].y += d[i].y; ab[i].z += d[i].z; ab[i].w += d[i].w;
      const float4 gmi = __ldg(gmp + i * 32 + lane);
      d[i].x *= gmi.x; d[i].y *= gmi.y; d[i].z *= gmi.z; d[i].w *= gmi.w;   // g = dy * gamma
      sg += (d[i].x + d[i].y) + (d[i].z + d[i].w);
      sgx += (d[i].x * v[i].x + d[i].y * v[i].y) + (d[i].z * v[i].z + d[i].w * v[i].w);
    }
    for (int o = 16; o; o >>= 1) {
      sg += __shfl_xor_sync(0xffffffffu, sg, o);
      sgx += __shfl_xor_sync(0xffffffffu, sgx, o);
    }
    const float mg = sg * (1.0f / COLS), mgx = sgx * (1.0f / COLS);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 o4;
      o4.x = rstd * (d[i].x - mg - v[i].x * mgx);
      o4.y = rstd * (d[i].y - mg - v[i].y * mgx);
      o4.z = rstd * (d[i].z - mg - v[i].z * mgx);
      o4.w = rstd * (d[i].w - mg - v[i].w * mgx);
      o4.x += rs[i].x; o4.y += rs[i].y; o4.z += rs[i].z; o4.w += rs[i].w;
      if (dx_f32) reinterpret_cast<float4*>(dx_f32 + row * COLS)[i * 32 + lane] = o4;
      if (dx_bf16)
        reinterpret_cast<uint2*>(dx_bf16 + row * COLS)[i * 32 + lane] =
            make_uint2(pack_bf16(o4.x, o4.y), pack_bf16(o4.z, o4.w));
      if constexpr (CS) {
        ac[i].x += o4.x; ac[i].y += o4.y; ac[i].z += o4.z; ac[i].w += o4.w;
      }
    }
  }
  if constexpr (CS) {
    // combine the 8 warps' column sums in shared memory, then one atomic per column and CTA
    __shared__ float scs[COLS];
    for (int w = 0; w < 8; ++w) {
      if (warp == w) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          float4* pc = reinterpret_cast<float4*>(scs) + i * 32 + lane;
          if (w == 0) {
            *pc = ac[i];
          } else {
            float4 a = *pc;
            a.x += ac[i].x; a.y += ac[i].y; a.z += ac[i].z; a.w += ac[i].w;
            *pc = a;
          }
        }
      }
      __syncthreads();
    }
    for (int c = threadIdx.x; c < COLS; c += 256) atomicAdd(dcol + c, scs[c]);
  }
  if (dgamma == nullptr) return;
  // combine the 8 warps' partial dgamma/dbeta: warp w adds its registers into shared memory in turn
  __shared__ float sgm[COLS], sbt[COLS];
  for (int w = 0; w < 8; ++w) {
    if (warp == w) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float4* pg = reinterpret_cast<float4*>(sgm) + i * 32 + lane;
        float4* pb = reinterpret_cast<float4*>(sbt) + i * 32 + lane;
        if (w == 0) {
          *pg = ag[i];
          *pb = ab[i];
        } else {
          float4 a = *pg, b = *pb;
          a.x += ag[i].x; a.y += ag[i].y; a.z += ag[i].z; a.w += ag[i].w;
          b.x += ab[i].x; b.y += ab[i].y; b.z += ab[i].z; b.w += ab[i].w;
          *pg = a;
          *pb = b;
        }
      }
    }
    __syncthreads();
  }
  for (int c = threadIdx.x; c < COLS; c += 256) {
    atomicAdd(dgamma + c, sgm[c]);
    atomicAdd(dbeta + c, sbt[c]);
  }
}

// ---------------------------------------------------------------------------------------------- heads backward
__device__ __forceinline__ float head_act_f(float x, int act) {
  if (act == 1) return tanhf(x);
  if (act == 2) return x > 0.f ? x : 0.01f * x;
  return x;
}
__device__ __forceinline__ float head_act_grad(float x, int act) {
  if (act == 1) {
    const float t = tanhf(x);
    return 1.0f - t * t;
  }
  if (act == 2) return x > 0.f ? 1.0f : 0.01f;
  return 1.0f;
}

constexpr int HB_ROWS = 32;
constexpr int HB_MAXN = 64;

// dh[r][k] = act_a'(h) * sum_j dA[r][j] Wa[j][k] + act_b'(h) * sum_j dB[r][j] Wb[j][k]
__global__ void __launch_bounds__(256)
heads_bwd_dh_kernel(const float* __restrict__ h, long long rows, int H, const float* __restrict__ da, int na,
                    const float* __restrict__ wa, int act_a, const float* __restrict__ db, int nb,
                    const float* __restrict__ wb, int act_b, float* __restrict__ dh) {
  __shared__ float sa[HB_ROWS][HB_MAXN];
  __shared__ float sb[HB_ROWS][HB_MAXN];
  const long long row0 = static_cast<long long>(blockIdx.x) * HB_ROWS;
  for (int i = threadIdx.x; i < HB_ROWS * HB_MAXN; i += 256) {
    const int r = i / HB_MAXN, j = i % HB_MAXN;
    const long long row = row0 + r;
    sa[r][j] = (row < rows && j < na) ? da[row * na + j] : 0.f;
    sb[r][j] = (row < rows && j < nb) ? db[row * nb + j] : 0.f;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < H; k += 256) {
    float acc_a[HB_ROWS], acc_b[HB_ROWS];
#pragma unroll
    for (int r = 0; r < HB_ROWS; ++r) acc_a[r] = acc_b[r] = 0.f;
    for (int j = 0; j < na; ++j) {
      const float w = __ldg(wa + static_cast<long long>(j) * H + k);
#pragma unroll
      for (int r = 0; r < HB_ROWS; ++r) acc_a[r] = fmaf(sa[r][j], w, acc_a[r]);
    }
    for (int j = 0; j < nb; ++j) {
      const float w = __ldg(wb + static_cast<long long>(j) * H + k);
#pragma unroll
      for (int r = 0; r < HB_ROWS; ++r) acc_b[r] = fmaf(sb[r][j], w, acc_b[r]);
    }
#pragma unroll
    for (int r = 0; r < HB_ROWS; ++r) {
      const long long row = row0 + r;
      if (row < rows) {
        const float x = __ldg(h + row * H + k);
        float g = 0.f;
        if (na) g = head_act_grad(x, act_a) * acc_a[r];
        if (nb) g = fmaf(head_act_grad(x, act_b), acc_b[r], g);
        dh[row * H + k] = g;
      }
    }
  }
}

// dW[j][k] += sum_r dOut[r][j] * act(h[r][k]);  db[j] += sum_r dOut[r][j].   grid = (H/256, row slabs of 128)
constexpr int HW_SLAB = 128;
__global__ void __launch_bounds__(256)
heads_bwd_dw_kernel(const float* __restrict__ h, long long rows, int H, const float* __restrict__ dout, int n, int act,
                    float* __restrict__ dw, float* __restrict__ dbias) {
  __shared__ float sd[HW_SLAB][HB_MAXN];
  const long long row0 = static_cast<long long>(blockIdx.y) * HW_SLAB;
  for (int i = threadIdx.x; i < HW_SLAB * HB_MAXN; i += 256) {
    const int r = i / HB_MAXN, j = i % HB_MAXN;
    const long long row = row0 + r;
    sd[r][j] = (row < rows && j < n) ? dout[row * n + j] : 0.f;
  }
  __syncthreads();
  const int k = blockIdx.x * 256 + threadIdx.x;
  float acc[HB_MAXN];
#pragma unroll
  for (int j = 0; j < HB_MAXN; ++j) acc[j] = 0.f;
  if (k < H) {
    const int nr = static_cast<int>(min(static_cast<long long>(HW_SLAB), rows - row0));
    for (int r = 0; r < nr; ++r) {
      const float a = head_act_f(__ldg(h + (row0 + r) * H + k), act);
#pragma unroll
      for (int j = 0; j < HB_MAXN; ++j) acc[j] = fmaf(sd[r][j], a, acc[j]);
    }
#pragma unroll
    for (int j = 0; j < HB_MAXN; ++j)
      if (j < n) atomicAdd(dw + static_cast<long long>(j) * H + k, acc[j]);
  }
  if (blockIdx.x == 0 && threadIdx.x < n && dbias) {
    float s = 0.f;
    for (int r = 0; r < HW_SLAB; ++r) s += sd[r][threadIdx.x];
    atomicAdd(dbias + threadIdx.x, s);
  }
}

// ---------------------------------------------------------------------------------------------- masked MSE + CE backward
// loss = 0.5 * mean_{tgt != -100} (pred - tgt)^2 + 0.5 * mean_{phn != 0} CE(logits, phn)   (models/aptai.py:89-102)
// accum = the forward kernel's sums: [1] = number of TV targets, [3] = number of CE frames.  One warp per row.
__global__ void __launch_bounds__(256)
mse_ce_bwd_kernel(const float* __restrict__ tv_pred, const float* __restrict__ tv_tgt,
                  const float* __restrict__ logits, const long long* __restrict__ phn_tgt, long long rows, int ntv,
                  int V, const double* __restrict__ accum, const float* __restrict__ gscale,
                  float* __restrict__ d_tv, float* __restrict__ d_logits) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float up = gscale ? *gscale : 1.0f;
  const float ctv = up * static_cast<float>(1.0 / accum[1]);           // 0.5 * 2 / n_tv
  const float cce = up * static_cast<float>(0.5 / accum[3]);
  for (int j = lane; j < ntv; j += 32) {
    const float tg = tv_tgt[row * ntv + j];
    d_tv[row * ntv + j] = tg != -100.0f ? ctv * (tv_pred[row * ntv + j] - tg) : 0.f;
  }
  const long long tg = phn_tgt[row];
  const float* lg = logits + row * V;
  if (tg == 0) {
    for (int c = lane; c < V; c += 32) d_logits[row * V + c] = 0.f;
    return;
  }
  float m = -INFINITY;
  for (int c = lane; c < V; c += 32) m = fmaxf(m, lg[c]);
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float s = 0.f;
  for (int c = lane; c < V; c += 32) s += expf(lg[c] - m);
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float inv = 1.0f / s;
  for (int c = lane; c < V; c += 32) {
    const float pr = expf(lg[c] - m) * inv;
    d_logits[row * V + c] = cce * (pr - (c == tg ? 1.0f : 0.f));
  }
}

// ---------------------------------------------------------------------------------------------- weight-norm backward
// w[o][c][j] = g[j] v[o][c][j] / n[j], n[j] = ||v[:, :, j]||.  Given dW in the folded layout [H][taps][cpad]:
//   s[j] = sum_{o,c} dW v;   dg[j] += s[j] / n[j];   dv += g/n * (dW - v * s / n^2)
__global__ void posconv_wn_dot_kernel(const float* __restrict__ dwf, const float* __restrict__ v, int H, int cin,
                                      int taps, int cpad, double* __restrict__ dot, double* __restrict__ nrm2) {
  const int j = blockIdx.x;
  double s = 0, q = 0;
  const int n_oc = H * cin;
  for (int i = threadIdx.x; i < n_oc; i += blockDim.x) {
    const int o = i / cin, c = i - o * cin;
    const double vv = v[static_cast<long long>(i) * taps + j];
    s += static_cast<double>(dwf[(static_cast<long long>(o) * taps + j) * cpad + c]) * vv;
    q += vv * vv;
  }
  __shared__ double red[2][256];
  red[0][threadIdx.x] = s;
  red[1][threadIdx.x] = q;
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if (threadIdx.x < o) {
      red[0][threadIdx.x] += red[0][threadIdx.x + o];
      red[1][threadIdx.x] += red[1][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    dot[j] = red[0][0];
    nrm2[j] = red[1][0];
  }
}

__global__ void posconv_wn_apply_kernel(const float* __restrict__ dwf, const float* __restrict__ v,
                                        const float* __restrict__ g, const double* __restrict__ dot,
                                        const double* __restrict__ nrm2, int H, int cin, int taps, int cpad,
                                        float* __restrict__ dg, float* __restrict__ dv) {
  const long long total = static_cast<long long>(H) * cin * taps;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(i % taps);
    const long long oc = i / taps;
    const int c = static_cast<int>(oc % cin);
    const long long o = oc / cin;
    const double n2 = nrm2[j], n = sqrt(n2);
    const double dw = dwf[(o * taps + j) * cpad + c];
    dv[i] += static_cast<float>(static_cast<double>(g[j]) / n * (dw - static_cast<double>(v[i]) * dot[j] / n2));
    if (i < taps) dg[i] += static_cast<float>(dot[i] / sqrt(nrm2[i]));
  }
}

// ---------------------------------------------------------------------------------------------- attention backward helpers
// D[b][h][t] = sum_d dO[t][h*64+d] * O[t][h*64+d]   (bf16 inputs [M][heads*64]).  One warp per row: a lane reads 16 bytes
// (8 channels) of both tensors per step, 8 lanes cover one head and combine with three shuffles.
__global__ void __launch_bounds__(256)
attn_bwd_dot_kernel(const __nv_bfloat16* __restrict__ dO, const __nv_bfloat16* __restrict__ O, int B, int T,
                    int heads, float* __restrict__ D, float* __restrict__ zero_f32) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= static_cast<long long>(B) * T) return;
  const int H = heads * 64;
  const long long b_ = row / T, t = row - b_ * T;
  const uint4* pa = reinterpret_cast<const uint4*>(dO + row * H);
  const uint4* pb = reinterpret_cast<const uint4*>(O + row * H);
  for (int c0 = 0; c0 < H; c0 += 256) {
    const int c = c0 + lane * 8;
    float s = 0.f;
    if (c < H) {
      if (zero_f32 != nullptr) {           // the dQ accumulator of the backward kernel that follows: cleared here
        float4* z = reinterpret_cast<float4*>(zero_f32 + row * H + c);
        z[0] = make_float4(0.f, 0.f, 0.f, 0.f);
        z[1] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      const uint4 a = __ldg(pa + c / 8), b = __ldg(pb + c / 8);
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&aw[j]));
        const float2 y = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&bw[j]));
        s = fmaf(x.x, y.x, fmaf(x.y, y.y, s));
      }
    }
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    if ((lane & 7) == 0 && c < H) D[(b_ * heads + c / 64) * T + t] = s;
  }
}

// out_bf16[r][c] (row pitch ldo) = scale * x_f32[r][c]   (dq fp32 accumulator -> the q block of dqkv)
__global__ void scale_cast_kernel(const float* __restrict__ x, long long rows, int cols4, float scale,
                                  __nv_bfloat16* __restrict__ out, long long ldo) {
  const long long total = rows * cols4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cols4;
    const int c = static_cast<int>(i - r * cols4) * 4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    *reinterpret_cast<uint2*>(out + r * ldo + c) =
        make_uint2(pack_bf16(v.x * scale, v.y * scale), pack_bf16(v.z * scale, v.w * scale));
  }
}

// out[i] = dy[i] * gelu'(pre[i])   (backward of the positional conv's GELU, HF:366; pre = bf16 pre-activation)
__global__ void gelu_bwd_kernel(const float* __restrict__ dy, const __nv_bfloat16* __restrict__ pre, long long n4,
                                float* __restrict__ out) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 d = __ldg(reinterpret_cast<const float4*>(dy) + i);
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(pre) + i);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    reinterpret_cast<float4*>(out)[i] = make_float4(d.x * gelu_erf_grad(a.x), d.y * gelu_erf_grad(a.y),
                                                    d.z * gelu_erf_grad(b.x), d.w * gelu_erf_grad(b.y));
  }
}

// ---------------------------------------------------------------------------------------------- weight preparation
// After every optimizer step the fp32 master weights must be turned into the kernels' operand copies: bf16 [N][K] for
// the forward GEMMs (q rows pre-scaled by head_dim^-0.5, q/k/v fused), bf16 [K][N] (transposed, unscaled) for the
// dgrad GEMMs, fp32 fused biases.  One launch over a table of (source, destinations, scales) instead of ~800 small
// cast / cat / transpose launches; each 64x64 tile is read once and written in both orientations.
struct PrepEntry {
  const float* src;          // [rows][cols] fp32, contiguous
  __nv_bfloat16* dst;        // optional: dst[r * dst_ld + c] = scale * src
  __nv_bfloat16* dst_t;      // optional: dst_t[c * dst_t_ld + r] = scale_t * src
  float* dst_f32;            // optional: dst_f32[r * cols + c] = scale * src
  int rows, cols, dst_ld, dst_t_ld;
  float scale, scale_t;
  int tile0;                 // first tile of this entry in the launch
  int tiles_x;               // tiles per row of tiles
};

// one fp32 -> the 16 bits of a bf16 (fp16 = 0) or IEEE fp16 (fp16 = 1) value, carried in the bf16 storage type
__device__ __forceinline__ __nv_bfloat16 cvt_h16(float v, int fp16) {
  if (!fp16) return __float2bfloat16(v);
  const __half h = __float2half_rn(v);
  return *reinterpret_cast<const __nv_bfloat16*>(&h);
}

// One 64 x 64 tile per CTA: 16-byte loads (a row segment of the tile is 256 contiguous bytes), 8-byte stores of four
// 16-bit values in both orientations (the transposed one through a padded shared-memory tile), so every global access
// of a warp covers whole 128-byte lines; the first version moved one scalar per thread and reached a third of the HBM
// rate (1.05 ms per optimizer step for 302 M parameters: 8 B per parameter = 0.37 ms at the roofline).  Entries whose
// shapes or pitches are not multiples of 4 (none of the model's) take the scalar path.
constexpr int PW_T = 64;

__global__ void __launch_bounds__(256)
prepare_weights_kernel(const PrepEntry* __restrict__ entries, int n_entries, int fp16) {
  __shared__ float tile[PW_T][PW_T + 1];
  __shared__ PrepEntry e;
  if (threadIdx.x == 0) {
    int lo = 0, hi = n_entries - 1;
    const int t = blockIdx.x;
    while (lo < hi) {                       // last entry whose tile0 <= t
      const int mid = (lo + hi + 1) >> 1;
      if (entries[mid].tile0 <= t) lo = mid;
      else hi = mid - 1;
    }
    e = entries[lo];
  }
  __syncthreads();
  const int t = blockIdx.x - e.tile0;
  const int r0 = (t / e.tiles_x) * PW_T, c0 = (t % e.tiles_x) * PW_T;
  const bool vec = (e.cols & 3) == 0 && (e.dst_ld & 3) == 0 && (e.dst_t_ld & 3) == 0 &&
                   ((reinterpret_cast<uintptr_t>(e.src) | reinterpret_cast<uintptr_t>(e.dst) |
                     reinterpret_cast<uintptr_t>(e.dst_t) | reinterpret_cast<uintptr_t>(e.dst_f32)) & 15) == 0;
  if (vec) {
    const int q = threadIdx.x & 15, rr = threadIdx.x >> 4;        // 16 quads per row segment, 16 rows per pass
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + rr + 16 * i, c = c0 + 4 * q;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < e.rows && c < e.cols) {
        v = __ldg(reinterpret_cast<const float4*>(e.src + static_cast<long long>(r) * e.cols + c));
        const float4 sv = make_float4(v.x * e.scale, v.y * e.scale, v.z * e.scale, v.w * e.scale);
        if (e.dst)
          *reinterpret_cast<uint2*>(e.dst + static_cast<long long>(r) * e.dst_ld + c) =
              make_uint2(pack_h16(sv.x, sv.y, fp16), pack_h16(sv.z, sv.w, fp16));
        if (e.dst_f32) *reinterpret_cast<float4*>(e.dst_f32 + static_cast<long long>(r) * e.cols + c) = sv;
      }
      tile[rr + 16 * i][4 * q] = v.x; tile[rr + 16 * i][4 * q + 1] = v.y;
      tile[rr + 16 * i][4 * q + 2] = v.z; tile[rr + 16 * i][4 * q + 3] = v.w;
    }
    if (e.dst_t == nullptr) return;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = c0 + rr + 16 * i, r = r0 + 4 * q;              // four consecutive source rows of one column
      if (c < e.cols && r < e.rows) {                              // rows % 4 == 0 (dst_t_ld is, and it is >= rows)
        const float a0 = tile[4 * q][rr + 16 * i] * e.scale_t, a1 = tile[4 * q + 1][rr + 16 * i] * e.scale_t;
        const float a2 = tile[4 * q + 2][rr + 16 * i] * e.scale_t, a3 = tile[4 * q + 3][rr + 16 * i] * e.scale_t;
        if (r + 3 < e.rows) {
          *reinterpret_cast<uint2*>(e.dst_t + static_cast<long long>(c) * e.dst_t_ld + r) =
              make_uint2(pack_h16(a0, a1, fp16), pack_h16(a2, a3, fp16));
        } else {
          const float av[4] = {a0, a1, a2, a3};
          for (int j = 0; j < 4 && r + j < e.rows; ++j)
            e.dst_t[static_cast<long long>(c) * e.dst_t_ld + r + j] = cvt_h16(av[j], fp16);
        }
      }
    }
    return;
  }
  // scalar path
  for (int i = threadIdx.x; i < PW_T * PW_T; i += 256) {
    const int lr = i / PW_T, lc = i % PW_T;
    const int r = r0 + lr, c = c0 + lc;
    float v = 0.f;
    if (r < e.rows && c < e.cols) {
      v = __ldg(e.src + static_cast<long long>(r) * e.cols + c);
      if (e.dst) e.dst[static_cast<long long>(r) * e.dst_ld + c] = cvt_h16(v * e.scale, fp16);
      if (e.dst_f32) e.dst_f32[static_cast<long long>(r) * e.cols + c] = v * e.scale;
    }
    tile[lr][lc] = v;
  }
  if (e.dst_t == nullptr) return;
  __syncthreads();
  for (int i = threadIdx.x; i < PW_T * PW_T; i += 256) {
    const int lc = i / PW_T, lr = i % PW_T;
    const int r = r0 + lr, c = c0 + lc;
    if (r < e.rows && c < e.cols)
      e.dst_t[static_cast<long long>(c) * e.dst_t_ld + r] = cvt_h16(tile[lr][lc] * e.scale_t, fp16);
  }
}

// ---------------------------------------------------------------------------------------------- fused Adam
// torch.optim.Adam (train/train_aptai.py:350-356) over a table of tensors in one launch.  Gradients live in one flat
// fp32 buffer (element offset goff[i]), exp_avg / exp_avg_sq in two more (offset soff[i]); parameters stay the
// module's own tensors.
struct AdamChunk {
  int tensor;
  int _pad;
  long long start;
};

__global__ void __launch_bounds__(256)
adam_kernel(float* const* __restrict__ params, const long long* __restrict__ goff, const long long* __restrict__ soff,
            const long long* __restrict__ numel, const AdamChunk* __restrict__ chunks, int chunk_elems,
            const float* __restrict__ grad,
            float* __restrict__ m, float* __restrict__ v, float lr, float beta1, float beta2, float eps,
            float weight_decay, float bc1, float bc2_sqrt, float grad_scale) {
  const AdamChunk ck = chunks[blockIdx.x];
  float* p = params[ck.tensor];
  const long long n = numel[ck.tensor];
  const long long gbase = goff[ck.tensor];
  const long long base = soff[ck.tensor];
  const long long end = min(n, ck.start + chunk_elems);
  const float step = lr / bc1;
  auto update = [&](float g, float& w, float& mi, float& vi) {
    g *= grad_scale;
    if (weight_decay != 0.f) g = fmaf(weight_decay, w, g);
    mi = fmaf(beta1, mi - g, g);                               // beta1*m + (1-beta1)*g  (lerp form, as torch)
    vi = beta2 * vi + (1.0f - beta2) * g * g;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    w = w - step * (mi / denom);
  };
  long long i0 = ck.start;
  // 16-byte accesses (28 B per parameter move through this kernel: four times the bytes in flight per thread of the
  // scalar loop); the flat gradient / moment buffers start every tensor on a 256-byte boundary and chunks start on
  // multiples of the chunk size, so only the parameter tensor's own alignment has to be checked
  if ((reinterpret_cast<uintptr_t>(p) & 15) == 0 && (chunk_elems & 3) == 0) {
    const long long end4 = ck.start + ((end - ck.start) & ~3LL);
    for (long long i = ck.start + 4 * threadIdx.x; i < end4; i += 4 * 256) {
      const float4 g4 = *reinterpret_cast<const float4*>(grad + gbase + i);
      float4 w4 = *reinterpret_cast<const float4*>(p + i);
      float4 m4 = *reinterpret_cast<const float4*>(m + base + i);
      float4 v4 = *reinterpret_cast<const float4*>(v + base + i);
      update(g4.x, w4.x, m4.x, v4.x);
      update(g4.y, w4.y, m4.y, v4.y);
      update(g4.z, w4.z, m4.z, v4.z);
      update(g4.w, w4.w, m4.w, v4.w);
      *reinterpret_cast<float4*>(m + base + i) = m4;
      *reinterpret_cast<float4*>(v + base + i) = v4;
      *reinterpret_cast<float4*>(p + i) = w4;
    }
    i0 = end4;
  }
  for (long long i = i0 + threadIdx.x; i < end; i += 256) {
    float w = p[i], mi = m[base + i], vi = v[base + i];
    update(grad[gbase + i], w, mi, vi);
    m[base + i] = mi;
    v[base + i] = vi;
    p[i] = w;
  }
}

}  // namespace aptai

using namespace aptai;

extern "C" int aptai_colsum(const void* x, int x_bf16, int64_t M, int N, int64_t ld, float scale, float* out,
                            void* stream) {
  if (int rc = check_arch()) return rc;
  APTAI_REQUIRE(x && out && M >= 1 && N >= 1, "colsum: bad arguments");
  APTAI_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && ld % (x_bf16 ? 8 : 4) == 0,
                "colsum: x must be 16-byte aligned with a row pitch that is a multiple of 16 bytes");
  const int cpb = x_bf16 ? 256 : 128;
  // enough row slabs to fill the machine a few times over, at least 64 rows each
  const int col_blocks = (N + cpb - 1) / cpb;
  long long slabs = (4LL * num_sms() + col_blocks - 1) / col_blocks;
  if (slabs > (M + 63) / 64) slabs = (M + 63) / 64;
  if (slabs < 1) slabs = 1;
  const int rpb = static_cast<int>((M + slabs - 1) / slabs);
  dim3 grid(col_blocks, static_cast<unsigned>((M + rpb - 1) / rpb));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (x_bf16) colsum_kernel<true><<<grid, 256, 0, st>>>(x, M, N, ld, rpb, scale, out);
  else colsum_kernel<false><<<grid, 256, 0, st>>>(x, M, N, ld, rpb, scale, out);
  return after_launch("colsum");
}

template <int NV>
static void launch_ln_bwd(const float* dy, const float* x, long long rows, const float* gamma, float eps,
                          const float* dres, float* dx, void* dxb, float* dg, float* db, float* dcol, cudaStream_t st) {
  long long blocks = (rows + 7) / 8;
  const long long cap = 4LL * num_sms();
  if (blocks > cap) blocks = cap;
  if (dcol != nullptr)
    layernorm_bwd_kernel<NV, true><<<static_cast<unsigned>(blocks), 256, 0, st>>>(
        dy, x, rows, gamma, eps, dres, dx, reinterpret_cast<__nv_bfloat16*>(dxb), dg, db, dcol);
  else
    layernorm_bwd_kernel<NV, false><<<static_cast<unsigned>(blocks), 256, 0, st>>>(
        dy, x, rows, gamma, eps, dres, dx, reinterpret_cast<__nv_bfloat16*>(dxb), dg, db, nullptr);
}

extern "C" int aptai_layernorm_bwd_colsum(const float* dy, const float* x, int64_t rows, int cols, const float* gamma,
                                          float eps, const float* dres, float* dx_f32, void* dx_bf16, float* dgamma,
                                          float* dbeta, float* dcolsum, void* stream);

extern "C" int aptai_layernorm_bwd(const float* dy, const float* x, int64_t rows, int cols, const float* gamma,
                                   float eps, const float* dres, float* dx_f32, void* dx_bf16, float* dgamma,
                                   float* dbeta, void* stream) {
  return aptai_layernorm_bwd_colsum(dy, x, rows, cols, gamma, eps, dres, dx_f32, dx_bf16, dgamma, dbeta, nullptr, stream);
}

extern "C" int aptai_layernorm_bwd_colsum(const float* dy, const float* x, int64_t rows, int cols, const float* gamma,
                                          float eps, const float* dres, float* dx_f32, void* dx_bf16, float* dgamma,
                                          float* dbeta, float* dcolsum, void* stream) {
  if (int rc = check_arch()) return rc;
  APTAI_REQUIRE(dy && x && gamma && (dx_f32 || dx_bf16), "layernorm_bwd: null pointer");
  APTAI_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "layernorm_bwd: dgamma and dbeta come together");
  APTAI_REQUIRE(rows >= 1, "layernorm_bwd: rows");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (cols) {
    case 256: launch_ln_bwd<2>(dy, x, rows, gamma, eps, dres, dx_f32, dx_bf16, dgamma, dbeta, dcolsum, st); break;
    case 512: launch_ln_bwd<4>(dy, x, rows, gamma, eps, dres, dx_f32, dx_bf16, dgamma, dbeta, dcolsum, st); break;
    case 768: launch_ln_bwd<6>(dy, x, rows, gamma, eps, dres, dx_f32, dx_bf16, dgamma, dbeta, dcolsum, st); break;
    case 1024: launch_ln_bwd<8>(dy, x, rows, gamma, eps, dres, dx_f32, dx_bf16, dgamma, dbeta, dcolsum, st); break;
    default:
      set_error("layernorm_bwd: unsupported width %d (256, 512, 768, 1024)", cols);
      return APTAI_ERR_ARG;
  }
  return after_launch("layernorm_bwd");
}

extern "C" int aptai_heads_bwd(const float* h, int64_t rows, int H, const float* da, int na, const float* wa,
                               int act_a, float* dwa, float* dba, const float* db, int nb, const float* wb,
                               int act_b, float* dwb, float* dbb, float* dh, void* stream) {
  if (int rc = check_arch()) return rc;
  APTAI_REQUIRE(h && rows >= 1 && H >= 1, "heads_bwd: bad input");
  APTAI_REQUIRE(na >= 0 && na <= HB_MAXN && nb >= 0 && nb <= HB_MAXN && na + nb > 0, "heads_bwd: head widths");
  APTAI_REQUIRE(na == 0 || (da && wa), "heads_bwd: head A pointers");
  APTAI_REQUIRE(nb == 0 || (db && wb), "heads_bwd: head B pointers");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dh) {
    heads_bwd_dh_kernel<<<static_cast<unsigned>((rows + HB_ROWS - 1) / HB_ROWS), 256, 0, st>>>(
        h, rows, H, da, na, wa, act_a, db, nb, wb, act_b, dh);
    if (int rc = after_launch("heads_bwd_dh")) return rc;
  }
  dim3 grid((H + 255) / 256, static_cast<unsigned>((rows + HW_SLAB - 1) / HW_SLAB));
  if (na && dwa) {
    heads_bwd_dw_kernel<<<grid, 256, 0, st>>>(h, rows, H, da, na, act_a, dwa, dba);
    if (int rc = after_launch("heads_bwd_dw")) return rc;
  }
  if (nb && dwb) {
    heads_bwd_dw_kernel<<<grid, 256, 0, st>>>(h, rows, H, db, nb, act_b, dwb, dbb);
    if (int rc = after_launch("heads_bwd_dw")) return rc;
  }
  return APTAI_OK;
}

extern "C" int aptai_masked_mse_ce_bwd(const float* tv_pred, const float* tv_tgt, const float* logits,
                                       const int64_t* phn_tgt, int64_t rows, int ntv, int V, const float* accum_ws,
                                       const float* grad_scale, float* d_tv, float* d_logits, void* stream) {
  if (int rc = check_arch()) return rc;
  APTAI_REQUIRE(tv_pred && tv_tgt && logits && phn_tgt && accum_ws && d_tv && d_logits, "mse_ce_bwd: null pointer");
  mse_ce_bwd_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      tv_pred, tv_tgt, logits, reinterpret_cast<const long long*>(phn_tgt), rows, ntv, V,
      reinterpret_cast<const double*>(accum_ws), grad_scale, d_tv, d_logits);
  return after_launch("mse_ce_bwd");
}

extern "C" int aptai_posconv_weightnorm_bwd(const float* dw_folded, const float* g, const float* v, int H, int cin,
                                            int taps, int cpad, float* dg, float* dv, void* ws /* 2*taps doubles */,
                                            void* stream) {
  if (int rc = check_arch()) return rc;
  APTAI_REQUIRE(dw_folded && g && v && dg && dv && ws, "posconv_weightnorm_bwd: null pointer");
  APTAI_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 7) == 0, "posconv_weightnorm_bwd: ws must be 8-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  double* dot = reinterpret_cast<double*>(ws);
  double* nrm2 = dot + taps;
  posconv_wn_dot_kernel<<<taps, 256, 0, st>>>(dw_folded, v, H, cin, taps, cpad, dot, nrm2);
  if (int rc = after_launch("posconv_wn_dot")) return rc;
  const long long total = static_cast<long long>(H) * cin * taps;
  int gx = static_cast<int>((total + 255) / 256);
  if (gx > 8192) gx = 8192;
  posconv_wn_apply_kernel<<<gx, 256, 0, st>>>(dw_folded, v, g, dot, nrm2, H, cin, taps, cpad, dg, dv);
  return after_launch("posconv_wn_apply");
}

extern "C" int aptai_attention_bwd_dot_zero(const void* d_ctx, const void* ctx, int B, int T, int heads, float* D,
                                            float* zero_f32, void* stream) {
  if (int rc = check_arch()) return rc;
  APTAI_REQUIRE(d_ctx && ctx && D && B >= 1 && T >= 1 && heads >= 1, "attention_bwd_dot: bad arguments");
  APTAI_REQUIRE((reinterpret_cast<uintptr_t>(zero_f32) & 15) == 0, "attention_bwd_dot: zero_f32 must be 16-byte aligned");
  const long long total = static_cast<long long>(B) * T;
  attn_bwd_dot_kernel<<<static_cast<unsigned>((total + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(d_ctx), reinterpret_cast<const __nv_bfloat16*>(ctx), B, T, heads, D,
      zero_f32);
  return after_launch("attention_bwd_dot");
}

extern "C" int aptai_attention_bwd_dot(const void* d_ctx, const void* ctx, int B, int T, int heads, float* D,
                                       void* stream) {
  return aptai_attention_bwd_dot_zero(d_ctx, ctx, B, T, heads, D, nullptr, stream);
}

extern "C" int aptai_scale_cast_bf16(const float* x, int64_t rows, int cols, float scale, void* out_bf16,
                                     int64_t ldo, void* stream) {
  if (int rc = check_arch()) return rc;
  APTAI_REQUIRE(x && out_bf16 && rows >= 1 && cols % 4 == 0 && ldo % 4 == 0, "scale_cast: bad arguments");
  const long long total = rows * (cols / 4);
  int gx = static_cast<int>((total + 255) / 256);
  if (gx > 16384) gx = 16384;
  scale_cast_kernel<<<gx, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, rows, cols / 4, scale, reinterpret_cast<__nv_bfloat16*>(out_bf16), ldo);
  return after_launch("scale_cast_bf16");
}

extern "C" int aptai_gelu_bwd(const float* dy, const void* pre_bf16, int64_t n, float* out, void* stream) {
  if (int rc = check_arch()) return rc;
  APTAI_REQUIRE(dy && pre_bf16 && out && n >= 4 && n % 4 == 0, "gelu_bwd: bad arguments");
  const long long n4 = n / 4;
  int gx = static_cast<int>((n4 + 255) / 256);
  if (gx > 16384) gx = 16384;
  gelu_bwd_kernel<<<gx, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      dy, reinterpret_cast<const __nv_bfloat16*>(pre_bf16), n4, out);
  return after_launch("gelu_bwd");
}

extern "C" int aptai_prepare_weights_fmt(const void* entries_dev, int n_entries, int total_tiles, int half_fmt,
                                         void* stream) {
  if (int rc = check_arch()) return rc;
  APTAI_REQUIRE(entries_dev && n_entries >= 1 && total_tiles >= 1, "prepare_weights: bad arguments");
  static_assert(sizeof(PrepEntry) == 64, "PrepEntry layout is part of the C ABI (aptai_b200/lib.py PrepEntry)");
  prepare_weights_kernel<<<total_tiles, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const PrepEntry*>(entries_dev), n_entries, half_fmt ? 1 : 0);
  return after_launch("prepare_weights");
}

extern "C" int aptai_prepare_weights(const void* entries_dev, int n_entries, int total_tiles, void* stream) {
  return aptai_prepare_weights_fmt(entries_dev, n_entries, total_tiles, 0, stream);
}

extern "C" int aptai_adam_step(void* const* params_dev, const int64_t* grad_offsets_dev,
                               const int64_t* state_offsets_dev, const int64_t* numel_dev, const void* chunks_dev, int n_chunks, int chunk_elems, const float* grad, float* exp_avg,
                               float* exp_avg_sq, float lr, float beta1, float beta2, float eps, float weight_decay,
                               int step, float grad_scale, void* stream) {
  if (int rc = check_arch()) return rc;
  APTAI_REQUIRE(params_dev && grad_offsets_dev && state_offsets_dev && numel_dev && chunks_dev && grad && exp_avg &&
                    exp_avg_sq,
                "adam_step: null pointer");
  APTAI_REQUIRE(n_chunks >= 1 && chunk_elems >= 256 && step >= 1, "adam_step: bad arguments");
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), step);
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), step);
  adam_kernel<<<n_chunks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<float* const*>(params_dev), reinterpret_cast<const long long*>(grad_offsets_dev),
      reinterpret_cast<const long long*>(state_offsets_dev), reinterpret_cast<const long long*>(numel_dev), reinterpret_cast<const AdamChunk*>(chunks_dev), chunk_elems, grad,
      exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, static_cast<float>(bc1),
      static_cast<float>(sqrt(bc2)), grad_scale);
  return after_launch("adam_step");
}
