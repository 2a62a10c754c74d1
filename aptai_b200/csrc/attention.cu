// Legacy-tensor-path (mma.sync) attention, kept only as the A/B baseline of the tcgen05 kernel in attention_tc.cu
// (entry point aptai_attention_fwd_mma; the product path calls aptai_attention_fwd).
// Padding-masked fused attention for the wav2vec2 encoder (HF:500-549; SDPA, non-causal, key-padding mask).
//
// Flash-style: one CTA = 64 query rows of one (utterance, head); K/V streamed in 64-key tiles through a
// cp.async double buffer; S = Q K^T and O = P V on bf16 tensor-core MMAs with fp32 accumulation; softmax
// statistics in fp32 (base-2 domain).  Keys at or beyond key_len[b] are masked with -inf; every query row is
// computed, because the reference lets padded queries attend to the valid keys (HF:438-463) and later stages
// (the FIR low-pass over padded frames, models/modules.py:46-61) read those rows.
// q must already be scaled by head_dim^-0.5 (folded into the q projection weights at plan time).
#include "common.h"
#include "ptx.cuh"

namespace aptai {

constexpr int AT_BQ = 64;
constexpr int AT_BK = 64;
constexpr int AT_D = 64;
constexpr int AT_THREADS = 128;
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ void cp_async16(void* dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// tile [64 rows][64 bf16] = 8 chunks of 16 B per row, chunk index XOR-swizzled with (row & 7)
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

__device__ __forceinline__ void load_tile(uint8_t* dst, const __nv_bfloat16* src, long long ld, int row0, int nrows_valid,
                                          int tid) {
  // 64 rows x 8 chunks = 512 chunks / 128 threads
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = tid + i * AT_THREADS;
    const int r = idx >> 3, ch = idx & 7;
    const bool ok = row0 + r < nrows_valid;
    const __nv_bfloat16* s = src + static_cast<long long>(ok ? row0 + r : 0) * ld + ch * 8;
    cp_async16(dst + tile_off(r, ch), s, ok ? 16 : 0);
  }
}

__global__ void __launch_bounds__(AT_THREADS)
attention_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ ctx,
                     const int* __restrict__ key_len, int T, int H) {
  __shared__ __align__(128) uint8_t sQ[AT_BQ * 128];
  __shared__ __align__(128) uint8_t sK[2][AT_BK * 128];
  __shared__ __align__(128) uint8_t sV[2][AT_BK * 128];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q0 = blockIdx.x * AT_BQ;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const long long ld = 3LL * H;
  const __nv_bfloat16* base = qkv + static_cast<long long>(b) * T * ld;
  const __nv_bfloat16* gq = base + head * AT_D;
  const __nv_bfloat16* gk = base + H + head * AT_D;
  const __nv_bfloat16* gv = base + 2 * H + head * AT_D;
  int klen = key_len ? key_len[b] : T;
  klen = max(1, min(klen, T));
  const int ntiles = (klen + AT_BK - 1) / AT_BK;

  load_tile(sQ, gq, ld, q0, T, tid);
  load_tile(sK[0], gk, ld, 0, T, tid);
  load_tile(sV[0], gv, ld, 0, T, tid);
  cp_async_commit();

  uint32_t qf[4][4];
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;

  for (int j = 0; j < ntiles; ++j) {
    const int buf = j & 1;
    if (j + 1 < ntiles) {
      load_tile(sK[buf ^ 1], gk, ld, (j + 1) * AT_BK, T, tid);
      load_tile(sV[buf ^ 1], gv, ld, (j + 1) * AT_BK, T, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (j == 0) {
      const uint32_t qb = smem_u32(sQ);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        ldsm_x4(qb + tile_off(warp * 16 + (lane & 15), kk * 2 + (lane >> 4)), qf[kk][0], qf[kk][1], qf[kk][2],
                qf[kk][3]);
    }
    // ---- S = Q K^T  (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) s[i][c] = 0.f;
    const uint32_t kb = smem_u32(sK[buf]);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(kb + tile_off(np * 16 + (lane & 7) + ((lane >> 4) << 3), kk * 2 + ((lane >> 3) & 1)), b0, b1, b2, b3);
        mma_bf16(s[np * 2], qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], b0, b1);
        mma_bf16(s[np * 2 + 1], qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], b2, b3);
      }
    }
    // ---- mask + online softmax (base-2)
    const int kv0 = j * AT_BK;
    float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k0 = kv0 + i * 8 + (lane & 3) * 2;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float v = (k0 + (c & 1) < klen) ? s[i][c] * LOG2E : -INFINITY;
        s[i][c] = v;
      }
      mx_lo = fmaxf(mx_lo, fmaxf(s[i][0], s[i][1]));
      mx_hi = fmaxf(mx_hi, fmaxf(s[i][2], s[i][3]));
    }
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
    const float mn_lo = fmaxf(m_lo, mx_lo), mn_hi = fmaxf(m_hi, mx_hi);
    const float c_lo = exp2f(m_lo - mn_lo), c_hi = exp2f(m_hi - mn_hi);   // first tile: exp2(-inf) = 0
    m_lo = mn_lo;
    m_hi = mn_hi;
    float rs_lo = 0.f, rs_hi = 0.f;
    uint32_t pf[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float p0 = exp2f(s[i][0] - mn_lo), p1 = exp2f(s[i][1] - mn_lo);
      const float p2 = exp2f(s[i][2] - mn_hi), p3 = exp2f(s[i][3] - mn_hi);
      rs_lo += p0 + p1;
      rs_hi += p2 + p3;
      pf[i][0] = pack_bf16(p0, p1);
      pf[i][1] = pack_bf16(p2, p3);
    }
    l_lo = l_lo * c_lo + rs_lo;
    l_hi = l_hi * c_hi + rs_hi;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o[i][0] *= c_lo; o[i][1] *= c_lo; o[i][2] *= c_hi; o[i][3] *= c_hi;
    }
    // ---- O += P V
    const uint32_t vb = smem_u32(sV[buf]);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(vb + tile_off(kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), dp * 2 + (lane >> 4)), b0, b1, b2, b3);
        mma_bf16(o[dp * 2], pf[kk * 2][0], pf[kk * 2][1], pf[kk * 2 + 1][0], pf[kk * 2 + 1][1], b0, b1);
        mma_bf16(o[dp * 2 + 1], pf[kk * 2][0], pf[kk * 2][1], pf[kk * 2 + 1][0], pf[kk * 2 + 1][1], b2, b3);
      }
    }
    __syncthreads();   // everyone done with buf before the next iteration overwrites it
  }
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
  const float i_lo = 1.f / l_lo, i_hi = 1.f / l_hi;
  const int r_lo = q0 + warp * 16 + (lane >> 2), r_hi = r_lo + 8;
  __nv_bfloat16* out = ctx + static_cast<long long>(b) * T * H + head * AT_D + (lane & 3) * 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (r_lo < T)
      *reinterpret_cast<uint32_t*>(out + static_cast<long long>(r_lo) * H + i * 8) =
          pack_bf16(o[i][0] * i_lo, o[i][1] * i_lo);
    if (r_hi < T)
      *reinterpret_cast<uint32_t*>(out + static_cast<long long>(r_hi) * H + i * 8) =
          pack_bf16(o[i][2] * i_hi, o[i][3] * i_hi);
  }
}

}  // namespace aptai

using namespace aptai;

extern "C" int aptai_attention_fwd_mma(const void* qkv, void* ctx, const int32_t* key_len, int B, int T, int heads,
                                   void* stream) {
  if (int rc = check_arch()) return rc;
  APTAI_REQUIRE(qkv && ctx, "attention: null pointer");
  APTAI_REQUIRE(B >= 1 && T >= 1 && heads >= 1, "attention: bad shape");
  APTAI_REQUIRE(B <= 65535 && heads <= 65535, "attention: grid limit");
  dim3 grid((T + AT_BQ - 1) / AT_BQ, heads, B);
  attention_fwd_kernel<<<grid, AT_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<__nv_bfloat16*>(ctx), key_len, T, heads * AT_D);
  return after_launch("attention_fwd_mma");
}
