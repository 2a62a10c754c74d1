// Grouped positional conv (HF:329-379: Conv1d(H, H, k = 128, groups = 16, 'same') + weight-norm + GELU, added to the
// hidden stream) for 64-channel groups, as a tcgen05 kernel that loads every input row ONCE.
//
// The generic implicit-GEMM path (gemm_tc.cu, BN = 64) streams one 128-row x 64-channel A tile per tap: 128 taps of
// the same rows shifted by one frame each, 16 KB of A + 8 KB of B per 128 cycles of MMA work = 192 B/clk per SM from
// L2 — the launch ran at the L2 -> SM crossbar limit (0.64 PFLOP/s, 3.4 % of the inference step for 2 % of its FLOPs).
// Here a work item is (utterance, 256-frame block, group) on a CTA PAIR (cta_group::2, M = 256):
//   * each CTA loads the 255-row x 64-channel SLAB of the padded input its 128 output frames read (32 KB, once,
//     double-buffered across items) and addresses tap t by moving the A descriptor's start address down t rows (128 B
//     each).  The 128-byte swizzle is a function of the ABSOLUTE shared-memory address bits (measured: with the slab
//     1024-byte aligned a start address t rows into it reads exactly what TMA wrote, descriptor base offset 0 —
//     bit-identical to the generic path for every tap; setting the base-offset field to (start >> 7) & 7 is wrong);
//   * the group's weights (64 x 8192 bf16 = 1 MB per item) stream through a 6-stage ring of 4-tap slices, half of the
//     64 output channels per CTA: 16 KB per CTA per 512 cycles of MMA work = 32 B/clk per SM;
//   * the 64-column accumulator is double-buffered in TMEM; the epilogue adds the bias, applies the erf-GELU and adds
//     the result to the fp32 hidden stream IN PLACE with TMA reduce-add stores (the residual is never loaded).
// Same arithmetic as the generic path (bf16 operands, fp32 accumulation over taps in the same order per k-block).
#include "common.h"
#include "ptx.cuh"

namespace aptai {

constexpr int PC_TAPS = 128, PC_GW = 64;            // taps, channels per group
constexpr int PC_ROWS = 128;                        // output frames per CTA (256 per pair)
constexpr int PC_SLAB_ROWS = PC_ROWS + PC_TAPS - 1; // 255 input rows per CTA
constexpr int PC_SLAB_BYTES = 256 * 128;            // 32 KB reserved per slab (255 rows used)
constexpr int PC_SLAB_TX = PC_SLAB_ROWS * 128;
constexpr int PC_TPS = 4;                           // taps per weight-ring stage
constexpr int PC_BH_BYTES = 32 * 128;               // one tap's half B tile: 32 output channels x 64 inputs
constexpr int PC_STAGE_BYTES = PC_TPS * PC_BH_BYTES;   // 16 KB per CTA
constexpr int PC_STAGES = 6;
constexpr int PC_OFF_B = 2 * PC_SLAB_BYTES;
constexpr int PC_OFF_STG = PC_OFF_B + PC_STAGES * PC_STAGE_BYTES;
constexpr int PC_OFF_BAR = PC_OFF_STG + 8 * 4096;   // 4 KB staging slab per epilogue warp (32 rows x 32 fp32)
constexpr int PC_SMEM = PC_OFF_BAR + 256;
constexpr int PC_THREADS = 352;                     // warps 0..7 epilogue, 8 TMA + TMEM alloc, 9 MMA, 10 spare
constexpr int PC_W_TMA = 8, PC_W_MMA = 9;

struct PosconvParams {
  const float* bias;
  int B, T, m_blocks, groups, items;
  int fp16;                 // x_pad and w_fold are IEEE fp16 instead of bf16
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PC_THREADS, 1)
posconv_slab_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                    const __grid_constant__ CUtensorMap tmH, const PosconvParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("aptai posconv: dynamic shared memory base is not 1024-byte aligned\n");
    __trap();
  }
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + PC_OFF_BAR);
  uint64_t* slab_full = bars + 0;        // [2]
  uint64_t* slab_empty = bars + 2;       // [2]
  uint64_t* b_full = bars + 4;           // [STAGES]
  uint64_t* b_empty = bars + 4 + PC_STAGES;
  uint64_t* tfull = bars + 4 + 2 * PC_STAGES;      // [2]
  uint64_t* tempty = tfull + 2;                    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const uint32_t rank = __shfl_sync(0xffffffffu, cluster_ctarank(), 0);
  const int item0 = blockIdx.x >> 1, item_step = gridDim.x >> 1;

  if (warp == PC_W_TMA && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmH);
  }
  if (warp == PC_W_MMA && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&slab_full[i], 1);
      mbar_init(&slab_empty[i], 1);
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 16);                   // 8 epilogue warps in each CTA of the pair
    }
    for (int i = 0; i < PC_STAGES; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == PC_W_TMA) tmem_alloc_cg2(tmem_slot, 128);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int per_b = p.m_blocks * p.groups;

  if (warp == PC_W_TMA) {
    // ------------------------------------------------------------------ TMA producer (whole warp, elected issue)
    uint32_t it = 0, stage = 0, phase = 0;
    for (int item = item0; item < p.items; item += item_step, ++it) {
      const int b = item / per_b, rem = item - b * per_b;
      const int mb = rem / p.groups, g = rem - mb * p.groups;
      const uint32_t sb = it & 1;
      mbar_wait_backoff(&slab_empty[sb], ((it >> 1) & 1) ^ 1, 64);
      if (elect_one()) {
        // both CTAs' slabs are counted on the LEADER's barrier (the leader issues every MMA)
        if (rank == 0) mbar_expect_tx(&slab_full[sb], 2 * PC_SLAB_TX);
        tma_load_3d_cg2(&tmX, map_to_cta(&slab_full[sb], 0), smem + sb * PC_SLAB_BYTES, g * PC_GW,
                        mb * 2 * PC_ROWS + static_cast<int>(rank) * PC_ROWS, b);
      }
      __syncwarp();
      for (int s = 0; s < PC_TAPS / PC_TPS; ++s) {
        mbar_wait_backoff(&b_empty[stage], phase ^ 1, 64);
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx(&b_full[stage], 2 * PC_STAGE_BYTES);
          const uint32_t fb = map_to_cta(&b_full[stage], 0);
          uint8_t* dst = smem + PC_OFF_B + stage * PC_STAGE_BYTES;
#pragma unroll
          for (int j = 0; j < PC_TPS; ++j)
            tma_load_2d_cg2(&tmW, fb, dst + j * PC_BH_BYTES, (s * PC_TPS + j) * PC_GW,
                            g * PC_GW + static_cast<int>(rank) * 32);
        }
        __syncwarp();
        if (++stage == PC_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == PC_W_MMA) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA; whole warp, elected)
    if (rank == 0) {
      const uint32_t IDESC = p.fp16 ? umma_idesc_f16(2 * PC_ROWS, PC_GW) : umma_idesc_bf16(2 * PC_ROWS, PC_GW);
      const uint32_t smem_base = smem_u32(smem);
      uint32_t it = 0, stage = 0, phase = 0;
      for (int item = item0; item < p.items; item += item_step, ++it) {
        const uint32_t sb = it & 1, acc = it & 1;
        mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
        mbar_wait(&slab_full[sb], (it >> 1) & 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * PC_GW;
        const uint32_t slab = smem_base + sb * PC_SLAB_BYTES;
        for (int s = 0; s < PC_TAPS / PC_TPS; ++s) {
          mbar_wait(&b_full[stage], phase);
          tc_fence_after();
          const uint32_t b_base = smem_base + PC_OFF_B + stage * PC_STAGE_BYTES;
          if (elect_one()) {
#pragma unroll
            for (int j = 0; j < PC_TPS; ++j) {
              const int tap = s * PC_TPS + j;
              const uint32_t a_row = slab + tap * 128;
              const uint64_t bdesc0 = umma_desc_sw128(b_base + j * PC_BH_BYTES);
#pragma unroll
#pragma unroll
              for (int k = 0; k < PC_GW / 16; ++k)
                umma_bf16_cg2(d_tmem, umma_desc_sw128(a_row + k * 32), bdesc0 + 2 * k, IDESC, (tap | k) != 0 ? 1u : 0u);
            }
            umma_commit_cg2(&b_empty[stage], 0x3);
            if (s == PC_TAPS / PC_TPS - 1) {
              umma_commit_cg2(&slab_empty[sb], 0x3);
              umma_commit_cg2(&tfull[acc], 0x3);
            }
          }
          __syncwarp();
          if (++stage == PC_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------------ epilogue: h += gelu(acc + bias)
    const int q = warp & 3, half = warp >> 2;          // TMEM lane quadrant, 32-column half of the 64-column tile
    const uint32_t tempty_leader0 = map_to_cta(&tempty[0], 0);
    float* stg = reinterpret_cast<float*>(smem + PC_OFF_STG) + warp * 1024;
    uint32_t it = 0;
    for (int item = item0; item < p.items; item += item_step, ++it) {
      const int b = item / per_b, rem = item - b * per_b;
      const int mb = rem / p.groups, g = rem - mb * p.groups;
      const uint32_t acc = it & 1;
      const int r_first = mb * 2 * PC_ROWS + static_cast<int>(rank) * PC_ROWS + q * 32;
      const int c0 = g * PC_GW + half * 32;
      mbar_wait(&tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      uint32_t r[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * PC_GW + half * 32, r);
      tmem_ld_wait();
      // the accumulator is in registers: hand it back before the GELU
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(tempty_leader0 + acc * 8);
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 b4 = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + c0 + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[i] = __uint_as_float(r[i]) + b4.x;
        v[i + 1] = __uint_as_float(r[i + 1]) + b4.y;
        v[i + 2] = __uint_as_float(r[i + 2]) + b4.z;
        v[i + 3] = __uint_as_float(r[i + 3]) + b4.w;
      }
#pragma unroll
      for (int i = 0; i < 32; i += 2) gelu_erf2(v[i], v[i + 1]);
      if (lane == 0) tma_store_wait_read<0>();
      __syncwarp();
#pragma unroll
      for (int u = 0; u < 8; ++u)
        *reinterpret_cast<float4*>(stg + lane * 32 + ((u ^ (lane & 7)) << 2)) =
            make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
      fence_async_proxy();
      __syncwarp();
      if (lane == 0) {
        tma_reduce_add_3d(&tmH, stg, c0, r_first, b);      // clipped at the utterance's last frame
        tma_store_commit();
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == PC_W_TMA) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, 128);
  }
}

}  // namespace aptai

using namespace aptai;

extern "C" int aptai_posconv_slab_fmt(const void* x_pad, const void* w_fold, const float* bias, float* h, int B, int T,
                                      int H, int half_fmt, void* stream) {
  if (int rc = check_arch()) return rc;
  APTAI_REQUIRE(x_pad && w_fold && h, "posconv_slab: null pointer");
  APTAI_REQUIRE(B >= 1 && T >= 1 && H >= PC_GW && H % PC_GW == 0, "posconv_slab: H must be a multiple of 64");
  APTAI_REQUIRE((reinterpret_cast<uintptr_t>(x_pad) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_fold) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(h) & 15) == 0,
                "posconv_slab: buffers must be 16-byte aligned");
  const int groups = H / PC_GW;
  CUtensorMap tx, tw, th;
  {
    // padded input [B][T + 128][H]: the slab of 255 rows x 64 channels; rows beyond the buffer are zero-filled
    uint64_t dims[3] = {static_cast<uint64_t>(H), static_cast<uint64_t>(T + PC_TAPS), static_cast<uint64_t>(B)};
    uint64_t strides[2] = {static_cast<uint64_t>(H) * 2, static_cast<uint64_t>(H) * 2 * (T + PC_TAPS)};
    uint32_t box[3] = {PC_GW, PC_SLAB_ROWS, 1};
    if (int rc = encode_tmap_bf16(&tx, x_pad, 3, dims, strides, box, 1)) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(PC_TAPS) * PC_GW, static_cast<uint64_t>(H)};
    uint64_t strides[1] = {static_cast<uint64_t>(PC_TAPS) * PC_GW * 2};
    uint32_t box[2] = {PC_GW, 32};
    if (int rc = encode_tmap_bf16(&tw, w_fold, 2, dims, strides, box, 1)) return rc;
  }
  {
    uint64_t dims[3] = {static_cast<uint64_t>(H), static_cast<uint64_t>(T), static_cast<uint64_t>(B)};
    uint64_t strides[2] = {static_cast<uint64_t>(H) * 4, static_cast<uint64_t>(H) * 4 * T};
    uint32_t box[3] = {32, 32, 1};
    if (int rc = encode_tmap_f32(&th, h, 3, dims, strides, box, 1)) return rc;
  }
  PosconvParams p;
  p.bias = bias;
  p.fp16 = half_fmt ? 1 : 0;
  p.B = B; p.T = T;
  p.m_blocks = (T + 2 * PC_ROWS - 1) / (2 * PC_ROWS);
  p.groups = groups;
  p.items = B * p.m_blocks * groups;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(posconv_slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PC_SMEM);
    if (e != cudaSuccess) {
      set_error("posconv_slab: cudaFuncSetAttribute(%d bytes): %s", PC_SMEM, cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    attr_set = true;
  }
  int clusters = num_sms() / 2;
  if (p.items < clusters) clusters = p.items;
  posconv_slab_kernel<<<2 * clusters, PC_THREADS, PC_SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(tx, tw, th, p);
  return after_launch("posconv_slab");
}

extern "C" int aptai_posconv_slab(const void* x_pad, const void* w_fold, const float* bias, float* h, int B, int T,
                                  int H, void* stream) {
  return aptai_posconv_slab_fmt(x_pad, w_fold, bias, h, B, T, H, 0, stream);
}
