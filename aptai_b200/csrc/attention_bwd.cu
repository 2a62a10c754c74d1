// Backward of the padding-masked self-attention (HF:500-549 via SDPA; autograd of softmax(q k^T + mask) v) on
// tcgen05 / TMEM / TMA for sm_100a.
//
// Work item = (utterance b, head h, block of 128 keys).  K_j and V_j stay in shared memory, dK_j and dV_j accumulate
// in TMEM over all 128-query tiles i; per tile:
//
//   MMA:      S^T  = K_j Q_i^T           dP^T = V_j dO_i^T                       (TMEM, 2 x 128 columns)
//   compute:  P^T  = exp2(S^T log2e - lse_i)      dS^T = P^T * (dP^T - D_i)      (bf16 -> shared memory, SW128 K-major)
//   MMA:      dV_j += P^T dO_i     dK_j += dS^T Q_i     dQ_ij = dS K_j           (TMEM 64 columns each)
//   compute:  dQ_ij -> fp32 red.global.add into dq32[B*T][H]   (partial sums over the key blocks)
//
// Every transposition is a descriptor bit: Q_i / dO_i are K-major B operands for S^T / dP^T and MN-major B operands
// for dK / dV; the dS^T tile written once to shared memory is the K-major A operand of dK and the MN-major A operand
// of dQ (two 64-query blocks, LBO = 16 KB).  lse (log2 domain) comes from the forward kernel, D = rowsum(dO * O)
// from aptai_attention_bwd_dot.  q was pre-scaled by head_dim^-0.5 in the forward, so S needs no scale here; the
// caller folds the scale into dq when it converts dq32 to bf16.
#include "common.h"
#include "ptx.cuh"

#include <math.h>

namespace aptai {

constexpr int AB_T = 128;                 // queries per tile = keys per block
constexpr int AB_D = 64;
constexpr int AB_TILE = AB_T * AB_D * 2;  // 16 KB
constexpr int AB_PT = AB_T * AB_T * 2;    // 32 KB (two [128][64] sub-tiles)
constexpr int AB_DATA = 2 * AB_TILE /*K,V*/ + 4 * AB_TILE /*Q,dO x2*/ + 4 * AB_PT /*P^T, dS^T x2*/;   // 224 KB
constexpr int AB_STATS = 2 * 2 * AB_T * 4;    // [buf][lse | D][128]
constexpr int AB_SMEM = AB_DATA + AB_STATS + 256;
// warps 0..7 compute, 8 TMA, 9 MMA (S^T, dP^T), 10 TMEM alloc, 11 / 12 / 13 MMA issuers of dV / dK / dQ: a tcgen05.mma
// issue costs its thread ~100 cycles while an N=64 MMA is 32 cycles of tensor work, so the 24 phase-2 MMAs of a tile
// are issued by three threads in parallel instead of one after the other
constexpr int AB_THREADS = 448;
constexpr int AB_W_TMA = 8, AB_W_MMA = 9, AB_W_ALLOC = 10, AB_W_DV = 11, AB_W_DK = 12, AB_W_DQ = 13;
constexpr uint32_t TB_ST = 0, TB_DPT = 128, TB_DK = 256, TB_DV = 320, TB_DQ = 384;
constexpr float AB_LOG2E = 1.4426950408889634f;

struct AttnBwdParams {
  const int* key_len;
  const float* lse;     // [B][heads][T] log2-domain log-sum-exp of the forward
  const float* dvec;    // [B][heads][T] rowsum(dO * O)
  float* dq32;          // [B*T][H] fp32, zeroed by the caller
  __nv_bfloat16* dqkv;  // [B*T][3H]: the k and v blocks are written here
  int B, T, heads, H, n_t, items;
  uint32_t drop_thresh24;      // attention-probability dropout of the forward (0 = off), same counter-based mask
  float drop_inv_keep;
  unsigned long long drop_seed;
};

__device__ __forceinline__ uint64_t ab_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__device__ __forceinline__ float ab_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <bool DROP>
__global__ void __launch_bounds__(AB_THREADS, 1)
attention_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                     const AttnBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem;
  uint8_t* sV = smem + AB_TILE;
  uint8_t* sQ = smem + 2 * AB_TILE;          // [2]
  uint8_t* sDO = smem + 4 * AB_TILE;         // [2]
  uint8_t* sPT = smem + 6 * AB_TILE;         // [2]  (double buffered like Q / dO: tile g uses buffer g & 1)
  uint8_t* sDST = sPT + 2 * AB_PT;           // [2]
  float* stats = reinterpret_cast<float*>(smem + AB_DATA);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AB_DATA + AB_STATS);
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("aptai attention_bwd: dynamic shared memory base is not 1024-byte aligned\n");
    __trap();
  }
  uint64_t* kv_full = bars + 0;
  uint64_t* kv_empty = bars + 1;
  uint64_t* qdo_full = bars + 2;    // [2]
  uint64_t* qdo_empty = bars + 4;   // [2]
  uint64_t* s_full = bars + 6;
  uint64_t* pds_full = bars + 7;    // [2]
  uint64_t* dq_full = bars + 9;
  uint64_t* dq_empty = bars + 10;
  uint64_t* dkv_full = bars + 11;
  uint64_t* dkv_empty = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  // warp index through a shuffle: provably warp-uniform, so the single-issuer roles (whole warp converged,
  // elect.sync around the issue) keep loop state and descriptors in uniform registers (see gemm_tc.cu)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  if (warp == AB_W_TMA && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
  }
  if (warp == AB_W_MMA && lane == 0) {
    mbar_init(kv_full, 1);
    mbar_init(kv_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qdo_full[i], 1);
      mbar_init(&qdo_empty[i], 3);      // dV, dK and dQ issuers have all consumed Q_i / dO_i / P^T / dS^T
    }
    mbar_init(s_full, 1);
    mbar_init(&pds_full[0], 8);
    mbar_init(&pds_full[1], 8);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 8);
    mbar_init(dkv_full, 2);             // dV and dK issuers
    mbar_init(dkv_empty, 8);
    fence_mbar_init();
  }
  if (warp == AB_W_ALLOC) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == AB_W_TMA) {
    {
      uint32_t g = 0, it = 0;
      for (int w = blockIdx.x; w < p.items; w += gridDim.x) {
        const int j = w % p.n_t;
        const int bh = w / p.n_t;
        const int h = bh % p.heads, b = bh / p.heads;
        const int klen = __shfl_sync(0xffffffffu, max(1, min(__ldg(p.key_len + b), p.T)), 0);
        if (j * AB_T >= klen) continue;
        const int row0 = b * p.T;
        mbar_wait_backoff(kv_empty, (it & 1) ^ 1, 64);
        if (elect_one()) {
          mbar_expect_tx(kv_full, 2 * AB_TILE);
          tma_load_2d(&tmQKV, kv_full, sK, p.H + h * AB_D, row0 + j * AB_T);
          tma_load_2d(&tmQKV, kv_full, sV, 2 * p.H + h * AB_D, row0 + j * AB_T);
        }
        __syncwarp();
        for (int i = 0; i < p.n_t; ++i, ++g) {
          const uint32_t buf = g & 1;
          mbar_wait_backoff(&qdo_empty[buf], ((g >> 1) & 1) ^ 1, 64);
          if (elect_one()) {
            mbar_expect_tx(&qdo_full[buf], 2 * AB_TILE);
            tma_load_2d(&tmQKV, &qdo_full[buf], sQ + buf * AB_TILE, h * AB_D, row0 + i * AB_T);
            tma_load_2d(&tmDO, &qdo_full[buf], sDO + buf * AB_TILE, h * AB_D, row0 + i * AB_T);
          }
          __syncwarp();
        }
        ++it;
      }
    }
  } else if (warp == AB_W_MMA) {
    // ---------------------------------------------------------------- phase 1: S^T = K Q^T, dP^T = V dO^T
    {
      constexpr uint32_t ID_S0 = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(AB_T >> 4) << 24);   // K-major
      uint32_t g = 0, it = 0;
      for (int w = blockIdx.x; w < p.items; w += gridDim.x) {
        const int j = w % p.n_t;
        const int b = (w / p.n_t) / p.heads;
        const int klen = __shfl_sync(0xffffffffu, max(1, min(__ldg(p.key_len + b), p.T)), 0);
        if (j * AB_T >= klen) continue;
        mbar_wait(kv_full, it & 1);
        const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV);
        for (int i = 0; i < p.n_t; ++i, ++g) {
          const uint32_t buf = g & 1;
          const uint32_t q_addr = smem_u32(sQ + buf * AB_TILE), do_addr = smem_u32(sDO + buf * AB_TILE);
          mbar_wait(&qdo_full[buf], (g >> 1) & 1);
          // the compute warps have finished reading tile g-1's S^T / dP^T once its P^T / dS^T are published
          if (g > 0) mbar_wait(&pds_full[(g - 1) & 1], ((g - 1) >> 1) & 1);
          tc_fence_after();
          // only the queries that exist (rounded up to 16): the last tile of an utterance is mostly padding
          const int nq16 = (min(AB_T, p.T - i * AB_T) + 15) & ~15;
          const uint32_t ID_S = ID_S0 | (static_cast<uint32_t>(nq16 >> 3) << 17);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < AB_D / 16; ++k)
              umma_bf16(tmem_base + TB_ST, umma_desc_sw128(k_addr + k * 32), umma_desc_sw128(q_addr + k * 32), ID_S,
                        k != 0 ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < AB_D / 16; ++k)
              umma_bf16(tmem_base + TB_DPT, umma_desc_sw128(v_addr + k * 32), umma_desc_sw128(do_addr + k * 32), ID_S,
                        k != 0 ? 1u : 0u);
            umma_commit(s_full);
          }
          __syncwarp();
        }
        ++it;
      }
    }
  } else if (warp == AB_W_DV || warp == AB_W_DK || warp == AB_W_DQ) {
    // ---------------------------------------------------------------- phase 2: dV += P^T dO | dK += dS^T Q | dQ = dS K
    {
      constexpr uint32_t ID_BASE = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(AB_T >> 4) << 24);
      constexpr uint32_t ID_KV = ID_BASE | (1u << 16) | (static_cast<uint32_t>(AB_D >> 3) << 17);    // N=64, B MN-major
      constexpr uint32_t ID_Q = ID_KV | (1u << 15);                                                   // A MN-major too
      uint32_t g = 0, it = 0;
      for (int w = blockIdx.x; w < p.items; w += gridDim.x) {
        const int j = w % p.n_t;
        const int b = (w / p.n_t) / p.heads;
        const int klen = __shfl_sync(0xffffffffu, max(1, min(__ldg(p.key_len + b), p.T)), 0);
        if (j * AB_T >= klen) continue;
        mbar_wait(kv_full, it & 1);
        const uint32_t k_addr = smem_u32(sK);
        for (int i = 0; i < p.n_t; ++i, ++g) {
          const uint32_t buf = g & 1;
          const uint32_t q_addr = smem_u32(sQ + buf * AB_TILE), do_addr = smem_u32(sDO + buf * AB_TILE);
          const uint32_t pt_addr = smem_u32(sPT + buf * AB_PT), dst_addr = smem_u32(sDST + buf * AB_PT);
          mbar_wait(&qdo_full[buf], (g >> 1) & 1);
          mbar_wait(&pds_full[buf], (g >> 1) & 1);             // P^T / dS^T of this tile are in shared memory
          // the dQ accumulator is single: the compute warps must have read out the previous tile's dQ
          if (warp == AB_W_DQ && g > 0) mbar_wait(dq_empty, (g - 1) & 1);
          if (i == 0 && warp != AB_W_DQ) mbar_wait(dkv_empty, (it & 1) ^ 1);   // previous item's dK/dV were read out
          tc_fence_after();
          const int kq = ((min(AB_T, p.T - i * AB_T) + 15) & ~15) / 16;        // 16-query steps that exist
          const int kk = ((min(AB_T, klen - j * AB_T) + 15) & ~15) / 16;       // 16-key steps that exist
          if (elect_one()) {
            if (warp == AB_W_DV) {
              for (int k = 0; k < kq; ++k)
                umma_bf16(tmem_base + TB_DV, umma_desc_sw128(pt_addr + (k >> 2) * (AB_PT / 2) + (k & 3) * 32),
                          ab_desc_mn(do_addr + k * 2048, 1024), ID_KV, (i | k) != 0 ? 1u : 0u);
            } else if (warp == AB_W_DK) {
              for (int k = 0; k < kq; ++k)
                umma_bf16(tmem_base + TB_DK, umma_desc_sw128(dst_addr + (k >> 2) * (AB_PT / 2) + (k & 3) * 32),
                          ab_desc_mn(q_addr + k * 2048, 1024), ID_KV, (i | k) != 0 ? 1u : 0u);
            } else {
              for (int k = 0; k < kk; ++k)
                umma_bf16(tmem_base + TB_DQ, ab_desc_mn(dst_addr + k * 2048, AB_PT / 2),
                          ab_desc_mn(k_addr + k * 2048, 1024), ID_Q, k != 0 ? 1u : 0u);
              umma_commit(dq_full);
            }
            umma_commit(&qdo_empty[buf]);
            if (i == p.n_t - 1) {
              if (warp != AB_W_DQ) umma_commit(dkv_full);
              else umma_commit(kv_empty);        // S^T / dP^T of the last tile completed long before (s_full)
            }
          }
          __syncwarp();
        }
        ++it;
      }
    }
  } else if (warp < 8) {
    const int q4 = warp & 3, hf = warp >> 2;
    const int row = q4 * 32 + lane;                       // TMEM lane: key row (phase 1) / query row (dQ)
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16);
    const int ctid = threadIdx.x;                         // 0..255
    uint32_t g = 0, it = 0;
    for (int w = blockIdx.x; w < p.items; w += gridDim.x) {
      const int j = w % p.n_t;
      const int bh = w / p.n_t;
      const int h = bh % p.heads, b = bh / p.heads;
      const int klen = max(1, min(__ldg(p.key_len + b), p.T));
      const long long row0 = static_cast<long long>(b) * p.T;
      const int key = j * AB_T + row;
      __nv_bfloat16* dk_out = p.dqkv + (row0 + key) * 3 * p.H + p.H + h * AB_D + hf * 32;
      __nv_bfloat16* dv_out = dk_out + p.H;
      if (j * AB_T >= klen) {
        // no valid key in this block: its dK / dV rows are zero
        if (key < p.T) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            *reinterpret_cast<uint4*>(dk_out + i) = make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(dv_out + i) = make_uint4(0u, 0u, 0u, 0u);
          }
        }
        continue;
      }
      const bool key_ok = key < klen;
      const uint32_t seed_bh = attn_drop_seed_bh(p.drop_seed, static_cast<uint32_t>(bh));
      const float* lse_bh = p.lse + static_cast<long long>(bh) * p.T;
      const float* d_bh = p.dvec + static_cast<long long>(bh) * p.T;
      // dQ partial of (query tile qi_tile, this key block): TMEM -> fp32 vector reductions into dq32
      auto dq_out = [&](int qi_tile, uint32_t gg) {
        mbar_wait(dq_full, gg & 1);
        tc_fence_after();
        uint32_t r[32];
        tmem_ld32(t_lane + TB_DQ + hf * 32, r);
        tmem_ld_wait();
        const int qi = qi_tile * AB_T + row;
        if (qi < p.T) {
          float* dst = p.dq32 + (row0 + qi) * p.H + h * AB_D + hf * 32;
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + e), "f"(__uint_as_float(r[e])),
                         "f"(__uint_as_float(r[e + 1])), "f"(__uint_as_float(r[e + 2])),
                         "f"(__uint_as_float(r[e + 3]))
                         : "memory");
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dq_empty);
      };
      for (int i = 0; i < p.n_t; ++i, ++g) {
        float* st = stats + (g & 1) * 2 * AB_T;
        {
          const int qi = i * AB_T + (ctid & 127);
          if (ctid < 128) st[ctid] = qi < p.T ? __ldg(lse_bh + qi) : INFINITY;
          else st[ctid] = qi < p.T ? __ldg(d_bh + qi) : 0.f;
        }
        named_bar_sync(1, 256);
        // s_full(g) implies that Q/dO of tile g were loaded, i.e. that the phase-2 MMAs of tile g-2 (the previous
        // users of buffer g & 1, P^T / dS^T included) have completed
        mbar_wait(s_full, g & 1);
        tc_fence_after();
        uint8_t* pt_row = sPT + (g & 1) * AB_PT + hf * (AB_PT / 2) + row * 128;
        uint8_t* ds_row = sDST + (g & 1) * AB_PT + hf * (AB_PT / 2) + row * 128;
        const int nq16 = (min(AB_T, p.T - i * AB_T) + 15) & ~15;
        const int nk16 = (min(AB_T, klen - j * AB_T) + 15) & ~15;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          // padding: query columns beyond the utterance are never read by dV / dK (their k-steps stop at nq16), key
          // rows beyond nk16 are never read by dQ; key rows in [klen, nk16) must be zero
          // (all decisions warp-uniform: tcgen05.ld below is a .sync.aligned instruction)
          if (hf * 64 + c * 32 >= nq16 || q4 * 32 >= nk16) continue;
          if (j * AB_T + q4 * 32 >= klen) {           // every key row of this warp is padding
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int unit = (c * 4 + u) ^ (row & 7);
              *reinterpret_cast<uint4*>(pt_row + (unit << 4)) = make_uint4(0u, 0u, 0u, 0u);
              *reinterpret_cast<uint4*>(ds_row + (unit << 4)) = make_uint4(0u, 0u, 0u, 0u);
            }
            continue;
          }
          uint32_t s[32], dp[32];
          tmem_ld32(t_lane + TB_ST + hf * 64 + c * 32, s);
          tmem_ld32(t_lane + TB_DPT + hf * 64 + c * 32, dp);
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float pv[8], dv[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int qc = hf * 64 + c * 32 + u * 8 + e;
              float pe = ab_ex2(fmaf(__uint_as_float(s[u * 8 + e]), AB_LOG2E, -st[qc]));
              if (!key_ok) pe = 0.f;
              float dpe = __uint_as_float(dp[u * 8 + e]);
              float pd = pe;                      // what fed P V in the forward: the dropped, rescaled probability
              if (DROP) {
                const float mk = attn_drop_keep(seed_bh, static_cast<uint32_t>(i * AB_T + qc), static_cast<uint32_t>(key),
                                                static_cast<uint32_t>(p.T), p.drop_thresh24) ? p.drop_inv_keep : 0.f;
                pd *= mk;
                dpe *= mk;
              }
              pv[e] = pd;
              dv[e] = pe * (dpe - st[AB_T + qc]);
            }
            const int unit = (c * 4 + u) ^ (row & 7);
            *reinterpret_cast<uint4*>(pt_row + (unit << 4)) =
                make_uint4(pack_bf16(pv[0], pv[1]), pack_bf16(pv[2], pv[3]), pack_bf16(pv[4], pv[5]),
                           pack_bf16(pv[6], pv[7]));
            *reinterpret_cast<uint4*>(ds_row + (unit << 4)) =
                make_uint4(pack_bf16(dv[0], dv[1]), pack_bf16(dv[2], dv[3]), pack_bf16(dv[4], dv[5]),
                           pack_bf16(dv[6], dv[7]));
          }
        }
        tc_fence_before();
        fence_async_proxy();
        __syncwarp();
        if (lane == 0) mbar_arrive(&pds_full[g & 1]);
        // dQ of the PREVIOUS query tile: its MMAs ran on the tensor pipe while this tile's P^T / dS^T were computed
        if (i > 0) dq_out(i - 1, g - 1);
      }
      dq_out(p.n_t - 1, g - 1);
      // dK_j, dV_j
      mbar_wait(dkv_full, it & 1);
      tc_fence_after();
      {
        uint32_t rk[32], rv[32];
        tmem_ld32(t_lane + TB_DK + hf * 32, rk);
        tmem_ld32(t_lane + TB_DV + hf * 32, rv);
        tmem_ld_wait();
        if (key < p.T && !key_ok) {       // padded key: its accumulator rows are undefined (P^T rows not written)
#pragma unroll
          for (int e = 0; e < 32; e += 8) {
            *reinterpret_cast<uint4*>(dk_out + e) = make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(dv_out + e) = make_uint4(0u, 0u, 0u, 0u);
          }
        } else if (key < p.T) {
#pragma unroll
          for (int e = 0; e < 32; e += 8) {
            *reinterpret_cast<uint4*>(dk_out + e) =
                make_uint4(pack_bf16(__uint_as_float(rk[e]), __uint_as_float(rk[e + 1])),
                           pack_bf16(__uint_as_float(rk[e + 2]), __uint_as_float(rk[e + 3])),
                           pack_bf16(__uint_as_float(rk[e + 4]), __uint_as_float(rk[e + 5])),
                           pack_bf16(__uint_as_float(rk[e + 6]), __uint_as_float(rk[e + 7])));
            *reinterpret_cast<uint4*>(dv_out + e) =
                make_uint4(pack_bf16(__uint_as_float(rv[e]), __uint_as_float(rv[e + 1])),
                           pack_bf16(__uint_as_float(rv[e + 2]), __uint_as_float(rv[e + 3])),
                           pack_bf16(__uint_as_float(rv[e + 4]), __uint_as_float(rv[e + 5])),
                           pack_bf16(__uint_as_float(rv[e + 6]), __uint_as_float(rv[e + 7])));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dkv_empty);
      ++it;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == AB_W_ALLOC) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace aptai

using namespace aptai;

static int attention_bwd_impl(const void* qkv, const void* d_ctx, const float* lse, const float* dvec,
                              const int32_t* key_len, int B, int T, int heads, float* dq32, void* dqkv, float drop_p,
                              uint64_t drop_seed, void* stream);

extern "C" int aptai_attention_bwd(const void* qkv, const void* d_ctx, const float* lse, const float* dvec,
                                   const int32_t* key_len, int B, int T, int heads, float* dq32, void* dqkv,
                                   void* stream) {
  return attention_bwd_impl(qkv, d_ctx, lse, dvec, key_len, B, T, heads, dq32, dqkv, 0.f, 0, stream);
}

extern "C" int aptai_attention_bwd_dropout(const void* qkv, const void* d_ctx, const float* lse, const float* dvec,
                                           const int32_t* key_len, int B, int T, int heads, float* dq32, void* dqkv,
                                           float drop_p, uint64_t drop_seed, void* stream) {
  return attention_bwd_impl(qkv, d_ctx, lse, dvec, key_len, B, T, heads, dq32, dqkv, drop_p, drop_seed, stream);
}

static int attention_bwd_impl(const void* qkv, const void* d_ctx, const float* lse, const float* dvec,
                              const int32_t* key_len, int B, int T, int heads, float* dq32, void* dqkv, float drop_p,
                              uint64_t drop_seed, void* stream) {
  if (int rc = check_arch()) return rc;
  APTAI_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "attention_bwd: dropout p must be in [0, 1)");
  APTAI_REQUIRE(qkv && d_ctx && lse && dvec && key_len && dq32 && dqkv, "attention_bwd: null pointer");
  APTAI_REQUIRE(B >= 1 && T >= 1 && heads >= 1, "attention_bwd: bad shape");
  APTAI_REQUIRE(((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(d_ctx) |
                  reinterpret_cast<uintptr_t>(dq32) | reinterpret_cast<uintptr_t>(dqkv)) & 15) == 0,
                "attention_bwd: buffers must be 16-byte aligned");
  const int H = heads * AB_D;
  CUtensorMap tmq, tmdo;
  uint32_t box[2] = {AB_D, AB_T};
  {
    uint64_t dims[2] = {static_cast<uint64_t>(3) * H, static_cast<uint64_t>(B) * T};
    uint64_t strides[1] = {static_cast<uint64_t>(3) * H * 2};
    if (int rc = encode_tmap_bf16(&tmq, qkv, 2, dims, strides, box, 1)) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(H), static_cast<uint64_t>(B) * T};
    uint64_t strides[1] = {static_cast<uint64_t>(H) * 2};
    if (int rc = encode_tmap_bf16(&tmdo, d_ctx, 2, dims, strides, box, 1)) return rc;
  }
  AttnBwdParams p;
  p.key_len = key_len;
  p.lse = lse;
  p.dvec = dvec;
  p.dq32 = dq32;
  p.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv);
  p.B = B; p.T = T; p.heads = heads; p.H = H;
  p.n_t = (T + AB_T - 1) / AB_T;
  p.items = B * heads * p.n_t;
  p.drop_thresh24 = static_cast<uint32_t>(static_cast<double>(drop_p) * 16777216.0);
  p.drop_inv_keep = 1.0f / (1.0f - drop_p);
  p.drop_seed = drop_seed;
  static bool attr_set = false;
  if (!attr_set) {
    for (int v = 0; v < 2; ++v) {
      const void* fn = v ? reinterpret_cast<const void*>(attention_bwd_kernel<true>)
                         : reinterpret_cast<const void*>(attention_bwd_kernel<false>);
      cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM);
      if (e != cudaSuccess) {
        set_error("attention_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return static_cast<int>(e);
      }
    }
    attr_set = true;
  }
  const int grid = p.items < num_sms() ? p.items : num_sms();
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (drop_p > 0.f) attention_bwd_kernel<true><<<grid, AB_THREADS, AB_SMEM, st>>>(tmq, tmdo, p);
  else attention_bwd_kernel<false><<<grid, AB_THREADS, AB_SMEM, st>>>(tmq, tmdo, p);
  return after_launch("attention_bwd");
}
