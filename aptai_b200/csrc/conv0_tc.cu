// Conv layer 0 (1 -> 512 channels, k = 10, s = 5) + LayerNorm / GroupNorm + GELU on the tensor cores.
//
// The SIMT kernel (frontend.cu: conv0_kernel) is bound by the FMA pipe and the issue slots: 10 packed FMAs + 10 shared
// loads per output pair before the normalisation and the GELU even start (ncu, profiles/r02_conv0.md: FMA pipe 59 %,
// 0.35 of the HBM floor).  Here the K = 10 contraction runs as ONE tcgen05 K-block per 128-frame tile on SPLIT
// operands, so that the fp32 result is kept:
//     x = xh + xm + xl,  w = wh + wm + wl   (three bf16 terms each: 24 mantissa bits)
//     y = xh wh + (xh wm + xm wh) + (xh wl + xm wm + xl wh)          (the terms below 2^-24 are dropped)
// = six products x ten taps = 60 of the 64 K slots of one SWIZZLE_128B row; three more slots carry the conv bias
// (A = 1, B = bias split in three), the last one is zero.  A rows are BUILT in shared memory from the waveform tile
// (which arrives by a 1-D TMA bulk copy, two tiles ahead), B (512 x 64 bf16, 64 KB) stays resident for the life of the
// persistent CTA.  The 128 x 512 fp32 tile lives in TMEM as two 256-column halves that ping-pong between the MMA
// issuer and the epilogue; the epilogue is what is left of the old kernel's inner loop: normalise (LayerNorm
// statistics are the closed form in the frame's 10 samples, frontend.cu), GELU, pack, and leave through SWIZZLE_128B
// staging slabs + TMA stores (clipped at T0 by the tensor map).  HBM floor: T0 x 512 x 2 B written per utterance.
//
// Reference: HF:281-323 (conv layer 0 + LayerNorm over channels / GroupNorm over time + GELU), as conv0_kernel.
#include "common.h"
#include "ptx.cuh"

namespace aptai {

namespace c0tc {
constexpr int C0 = 512, K0 = 10, S0 = 5;
constexpr int TF = 128;                         // frames per tile (= TMEM lanes)
constexpr int NX = 648;                         // samples staged per tile: (TF - 1) * S0 + K0 = 645, padded to 16 B
constexpr int X_BYTES = 2688;                   // NX * 4 = 2592, padded to 128 B
constexpr int EPI_W = 8, BLD_W = 4;
constexpr int W_BLD = EPI_W, W_MMA = EPI_W + BLD_W;
constexpr int THREADS = (EPI_W + BLD_W + 1) * 32;
constexpr int B_BYTES = C0 * 128;               // 512 rows x 64 bf16
constexpr int A_BYTES = TF * 128;
constexpr int SLAB = 4096;                      // 32 rows x 128 B staging slab (64 channels of 32 frames)
constexpr int OFF_B = 0;
constexpr int OFF_A = OFF_B + B_BYTES;          // 2 buffers
constexpr int OFF_STG = OFF_A + 2 * A_BYTES;    // EPI_W warps x 2 slabs
constexpr int OFF_X = OFF_STG + EPI_W * 2 * SLAB;
constexpr int OFF_VEC = OFF_X + 2 * X_BYTES;    // 2 buffers x (G[512] | E[512]) fp32
constexpr int OFF_STAT = OFF_VEC + 2 * 2 * C0 * 4;   // 2 buffers x float2[TF]
constexpr int OFF_CONST = OFF_STAT + 2 * TF * 8;     // LayerNorm closed-form constants (132 floats)
constexpr int OFF_BAR = OFF_CONST + 136 * 4;
constexpr int SMEM_BYTES = OFF_BAR + 128;
static_assert(OFF_A % 1024 == 0 && OFF_STG % 1024 == 0, "SWIZZLE_128B buffers must be 1024-byte aligned");
static_assert(OFF_X % 128 == 0 && OFF_BAR % 8 == 0, "alignment");
}  // namespace c0tc

__device__ __forceinline__ void bulk_load_1d(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Split operand B[c][64] (bf16) from the fp32 conv weight [512][10] and bias: see the K layout above.
__global__ void conv0_tc_prep_kernel(const float* __restrict__ w, const float* __restrict__ bias, int use_bias,
                                     __nv_bfloat16* __restrict__ bsplit) {
  using namespace c0tc;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C0) return;
  __nv_bfloat16 row[64];
  auto split3 = [](float v, __nv_bfloat16& h, __nv_bfloat16& m, __nv_bfloat16& l) {
    h = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(h);
    m = __float2bfloat16_rn(r1);
    l = __float2bfloat16_rn(r1 - __bfloat162float(m));
  };
#pragma unroll
  for (int j = 0; j < K0; ++j) {
    __nv_bfloat16 h, m, l;
    split3(w[c * K0 + j], h, m, l);
    row[0 * K0 + j] = h;      // x xh
    row[1 * K0 + j] = m;      // x xh
    row[2 * K0 + j] = h;      // x xm
    row[3 * K0 + j] = l;      // x xh
    row[4 * K0 + j] = m;      // x xm
    row[5 * K0 + j] = h;      // x xl
  }
  __nv_bfloat16 bh, bm, bl;
  split3((use_bias && bias) ? bias[c] : 0.f, bh, bm, bl);
  row[60] = bh; row[61] = bm; row[62] = bl; row[63] = __float2bfloat16_rn(0.f);
  uint4* dst = reinterpret_cast<uint4*>(bsplit + static_cast<size_t>(c) * 64);
  const uint4* src = reinterpret_cast<const uint4*>(row);
#pragma unroll
  for (int i = 0; i < 8; ++i) dst[i] = src[i];
}

struct Conv0TcParams {
  const float* wav;
  long long L;
  int B, T0, tiles_per_utt, num_tiles;
  const __nv_bfloat16* bsplit;   // [512][64]
  const float* consts;           // NORM 1: closed-form LayerNorm constants (conv0_ln_consts_kernel layout)
  const float* gamma;            // NORM 1
  const float* beta;             // NORM 1
  const float* affine;           // NORM 2: [B][512][2] scale / shift
  float eps;
  int out_fp16;
  int x_tma;                     // waveform tiles by cp.async.bulk (L % 4 == 0 and a 16-byte aligned base)
};

// NORM: 0 none, 1 LayerNorm over channels per frame, 2 GroupNorm affine per (utterance, channel)
template <int NORM>
__global__ void __launch_bounds__(c0tc::THREADS, 1)
conv0_tc_kernel(const __grid_constant__ CUtensorMap tmOut, const Conv0TcParams p) {
  using namespace c0tc;
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("aptai conv0_tc: dynamic shared memory base is not 1024-byte aligned\n");
    __trap();
  }
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);   // [2] builders -> MMA
  uint64_t* a_empty = a_full + 2;                                   // [2] MMA commit -> builders
  uint64_t* acc_full = a_empty + 2;                                 // [2] MMA commit -> epilogue (one per column half)
  uint64_t* acc_empty = acc_full + 2;                               // [2] epilogue -> MMA
  uint64_t* st_empty = acc_empty + 2;                               // [2] epilogue -> builders (stat / vec buffers)
  uint64_t* x_full = st_empty + 2;                                  // [2] bulk copy -> builders
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_full + 2);
  float* consts_s = reinterpret_cast<float*>(smem + OFF_CONST);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  if (warp == W_MMA && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], BLD_W * 32);
      mbar_init(&a_empty[i], 1);
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], EPI_W);          // every epilogue warp reads its 128 columns of BOTH halves
      mbar_init(&st_empty[i], EPI_W);
      mbar_init(&x_full[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == W_MMA) tmem_alloc(tmem_slot, 512);
  // resident B operand: 512 rows x 8 chunks of 16 B, chunk j of row r at position j ^ (r & 7) (SWIZZLE_128B)
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.bsplit);
    for (int i = threadIdx.x; i < C0 * 8; i += THREADS) {
      const int r = i >> 3, j = i & 7;
      *reinterpret_cast<uint4*>(smem + OFF_B + r * 128 + ((j ^ (r & 7)) << 4)) = __ldg(src + i);
    }
    if (NORM == 1)
      for (int i = threadIdx.x; i < 132; i += THREADS) consts_s[i] = __ldg(p.consts + i);
  }
  fence_async_proxy();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tile0 = blockIdx.x, tstep = gridDim.x;

  if (warp >= W_BLD && warp < W_MMA) {
    // ------------------------------------------------------------ builders: waveform tile -> A rows + statistics
    const int bt = threadIdx.x - W_BLD * 32;             // frame of the tile this thread builds (0..127)
    auto x_src = [&](int tile, uint32_t& bytes) -> const float* {
      const int b = tile / p.tiles_per_utt, ft = tile - b * p.tiles_per_utt;
      const long long x0 = static_cast<long long>(ft) * TF * S0;
      const long long left = p.L - x0;
      bytes = static_cast<uint32_t>((left < NX ? left : NX) * 4);
      return p.wav + static_cast<long long>(b) * p.L + x0;
    };
    if (p.x_tma && bt == 0) {
      for (int i = 0; i < 2; ++i) {
        const int tile = tile0 + i * tstep;
        if (tile < p.num_tiles) {
          uint32_t bytes;
          const float* src = x_src(tile, bytes);
          mbar_expect_tx(&x_full[i], bytes);
          bulk_load_1d(smem + OFF_X + i * X_BYTES, src, bytes, &x_full[i]);
        }
      }
    }
    int k = 0;
    for (int tile = tile0; tile < p.num_tiles; tile += tstep, ++k) {
      const int s = k & 1;
      const uint32_t u = static_cast<uint32_t>(k >> 1) & 1u;
      const int b = tile / p.tiles_per_utt, ft = tile - b * p.tiles_per_utt;
      float* xs = reinterpret_cast<float*>(smem + OFF_X + s * X_BYTES);
      if (p.x_tma) {
        mbar_wait_backoff(&x_full[s], u, 100);
      } else {
        uint32_t bytes;
        const float* src = x_src(tile, bytes);
        const int n = static_cast<int>(bytes >> 2);
        for (int i = bt; i < NX; i += BLD_W * 32) xs[i] = i < n ? __ldg(src + i) : 0.f;
        named_bar_sync(2, BLD_W * 32);
      }
      float f[K0];
#pragma unroll
      for (int j = 0; j < K0; ++j) f[j] = xs[bt * S0 + j];
      // the tail tile of an utterance: frames at or beyond T0 read stale / foreign samples; their rows are clipped by
      // the TMA store, but keep them finite
      if (ft * TF + bt >= p.T0) {
#pragma unroll
        for (int j = 0; j < K0; ++j) f[j] = 0.f;
      }
      named_bar_sync(3, BLD_W * 32);                      // every builder has read this waveform buffer
      if (p.x_tma && bt == 0) {
        const int nt = tile + 2 * tstep;
        if (nt < p.num_tiles) {
          uint32_t bytes;
          const float* src = x_src(nt, bytes);
          mbar_expect_tx(&x_full[s], bytes);
          bulk_load_1d(xs, src, bytes, &x_full[s]);
        }
      }
      // per-frame statistics {rstd, -mean * rstd}
      float rs = 1.f, nmr = 0.f;
      if (NORM == 1) {
        float mean = consts_s[10];
        float var = consts_s[21];
#pragma unroll
        for (int j = 0; j < K0; ++j) {
          mean = fmaf(consts_s[j], f[j], mean);
          float acc = 2.f * consts_s[11 + j];
#pragma unroll
          for (int q = 0; q < K0; ++q) acc = fmaf(consts_s[32 + j * K0 + q], f[q], acc);
          var = fmaf(acc, f[j], var);
        }
        rs = rsqrtf(fmaxf(var, 0.f) + p.eps);
        nmr = -mean * rs;
      }
      // split the ten samples and lay the row out in the K order of conv0_tc_prep_kernel
      uint32_t row[32];
      {
        __nv_bfloat16 h[K0], m[K0], l[K0];
#pragma unroll
        for (int j = 0; j < K0; ++j) {
          h[j] = __float2bfloat16_rn(f[j]);
          const float r1 = f[j] - __bfloat162float(h[j]);
          m[j] = __float2bfloat16_rn(r1);
          l[j] = __float2bfloat16_rn(r1 - __bfloat162float(m[j]));
        }
        __nv_bfloat16 e[64];
#pragma unroll
        for (int j = 0; j < K0; ++j) {
          e[0 * K0 + j] = h[j];
          e[1 * K0 + j] = h[j];
          e[2 * K0 + j] = m[j];
          e[3 * K0 + j] = h[j];
          e[4 * K0 + j] = m[j];
          e[5 * K0 + j] = l[j];
        }
        const __nv_bfloat16 one = __float2bfloat16_rn(1.f);
        e[60] = one; e[61] = one; e[62] = one; e[63] = __float2bfloat16_rn(0.f);
#pragma unroll
        for (int i = 0; i < 32; ++i)
          row[i] = static_cast<uint32_t>(__bfloat16_as_ushort(e[2 * i])) |
                   (static_cast<uint32_t>(__bfloat16_as_ushort(e[2 * i + 1])) << 16);
      }
      // the builders run up to two tiles ahead of the epilogue: wait with a back-off, their spinning would take issue
      // slots from the epilogue warps (12 % of all issued instructions in the first capture)
      mbar_wait_backoff(&a_empty[s], u ^ 1, 200);
      mbar_wait_backoff(&st_empty[s], u ^ 1, 200);
      uint8_t* arow = smem + OFF_A + s * A_BYTES + bt * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4*>(arow + ((j ^ (bt & 7)) << 4)) =
            make_uint4(row[4 * j], row[4 * j + 1], row[4 * j + 2], row[4 * j + 3]);
      reinterpret_cast<float2*>(smem + OFF_STAT)[s * TF + bt] = make_float2(rs, nmr);
      // scale / shift vectors of this tile's utterance: G[c] | E[c]
      {
        float4 g4, e4;
        if (NORM == 1) {
          g4 = __ldg(reinterpret_cast<const float4*>(p.gamma) + bt);
          e4 = __ldg(reinterpret_cast<const float4*>(p.beta) + bt);
        } else if (NORM == 2) {
          // {scale, shift} pairs; the array follows B * 65 doubles, so it is only 8-byte aligned
          const float2* af = reinterpret_cast<const float2*>(p.affine + (static_cast<long long>(b) * C0 + 4 * bt) * 2);
          const float2 a0 = __ldg(af), a1 = __ldg(af + 1), a2 = __ldg(af + 2), a3 = __ldg(af + 3);
          g4 = make_float4(a0.x, a1.x, a2.x, a3.x);
          e4 = make_float4(a0.y, a1.y, a2.y, a3.y);
        } else {
          g4 = make_float4(1.f, 1.f, 1.f, 1.f);
          e4 = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // pre-halved: the epilogue's GELU takes x / 2 (gelu_fast2_half)
        float4* vec = reinterpret_cast<float4*>(smem + OFF_VEC + s * 2 * C0 * 4);
        vec[bt] = make_float4(0.5f * g4.x, 0.5f * g4.y, 0.5f * g4.z, 0.5f * g4.w);
        vec[C0 / 4 + bt] = make_float4(0.5f * e4.x, 0.5f * e4.y, 0.5f * e4.z, 0.5f * e4.w);
      }
      fence_async_proxy();                               // the A row is read by the tensor core (async proxy)
      mbar_arrive(&a_full[s]);
    }
  } else if (warp == W_MMA) {
    // ------------------------------------------------------------ MMA issuer: one K block, two 256-column halves
    constexpr uint32_t IDESC = umma_idesc_bf16(TF, 256);
    const uint32_t smem_base = smem_u32(smem);
    int k = 0;
    for (int tile = tile0; tile < p.num_tiles; tile += tstep, ++k) {
      const int s = k & 1;
      const uint32_t u = static_cast<uint32_t>(k >> 1) & 1u;
      mbar_wait_backoff(&a_full[s], u, 32);
      for (int h = 0; h < 2; ++h) {
        mbar_wait_backoff(&acc_empty[h], (static_cast<uint32_t>(k) & 1u) ^ 1u, 32);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc0 = umma_desc_sw128(smem_base + OFF_A + s * A_BYTES);
          const uint64_t bdesc0 = umma_desc_sw128(smem_base + OFF_B + h * 256 * 128);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16(tmem_base + h * 256, adesc0 + (kk * 32 >> 4), bdesc0 + (kk * 32 >> 4), IDESC, kk != 0 ? 1u : 0u);
          umma_commit(&acc_full[h]);
          if (h == 1) umma_commit(&a_empty[s]);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue: TMEM -> normalise -> GELU -> TMA store
    const int q = warp & 3;                      // TMEM lane quadrant
    const int cp = warp >> 2;                    // column part inside each 256-column half
    const int row = q * 32 + lane;
    uint8_t* stg = smem + OFF_STG + warp * 2 * SLAB;
    int slab = 0;
    int k = 0;
    for (int tile = tile0; tile < p.num_tiles; tile += tstep, ++k) {
      const int s = k & 1;
      const int b = tile / p.tiles_per_utt, ft = tile - b * p.tiles_per_utt;
      const int t_first = ft * TF + q * 32;
      const float* vecG = reinterpret_cast<const float*>(smem + OFF_VEC + s * 2 * C0 * 4);
      const float* vecE = vecG + C0;
      uint64_t rs2 = 0, nmr2 = 0;
      for (int h = 0; h < 2; ++h) {
        mbar_wait(&acc_full[h], static_cast<uint32_t>(k) & 1u);
        tc_fence_after();
        if (h == 0) {
          // a_full (observed by the MMA warp before the commit this wait saw) orders the builders' writes
          const float2 st = reinterpret_cast<const float2*>(smem + OFF_STAT)[s * TF + row];
          rs2 = f32x2_pack(st.x, st.x);
          nmr2 = f32x2_pack(st.y, st.y);
        }
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + h * 256 + cp * 128;
        // one 32-column chunk: normalise (the scale / shift vectors are pre-halved: the GELU wants x / 2), GELU, pack,
        // stage; every second chunk completes a 32-row x 128-byte slab that one TMA store writes
        auto process = [&](uint32_t (&r)[32], int ch) {
          const int c0 = h * 256 + cp * 128 + ch * 32;
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 g4 = *reinterpret_cast<const float4*>(vecG + c0 + i);
            const float4 e4 = *reinterpret_cast<const float4*>(vecE + c0 + i);
            uint64_t y01 = f32x2_fma(f32x2_pack(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), rs2, nmr2);
            uint64_t y23 = f32x2_fma(f32x2_pack(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])), rs2, nmr2);
            y01 = f32x2_fma(y01, f32x2_pack(g4.x, g4.y), f32x2_pack(e4.x, e4.y));
            y23 = f32x2_fma(y23, f32x2_pack(g4.z, g4.w), f32x2_pack(e4.z, e4.w));
            f32x2_unpack(y01, v[i], v[i + 1]);
            f32x2_unpack(y23, v[i + 2], v[i + 3]);
            gelu_fast2_half(v[i], v[i + 1]);
            gelu_fast2_half(v[i + 2], v[i + 3]);
          }
          const int hsel = ch & 1;               // which 64-byte half of the slab's 128-byte rows this chunk fills
          uint4* sb = reinterpret_cast<uint4*>(stg + slab * SLAB);
          if (hsel == 0) {
            if (lane == 0) tma_store_wait_read<1>();      // the store that last read THIS slab (two stores ago) is done
            __syncwarp();
          }
          uint4 o16[4];
          pack32_h16(v, o16, p.out_fp16);
#pragma unroll
          for (int uu = 0; uu < 4; ++uu) sb[lane * 8 + ((hsel * 4 + uu) ^ (lane & 7))] = o16[uu];
          if (hsel == 1) {
            fence_async_proxy();
            __syncwarp();
            if (lane == 0) {
              if (t_first < p.T0) tma_store_3d(&tmOut, sb, c0 - 32, t_first, b);
              tma_store_commit();
            }
            slab ^= 1;
          }
        };
        // two register arrays alternate: the tcgen05.ld of chunk c + 1 is in flight while chunk c is processed
        uint32_t ra[32], rb[32];
        tmem_ld32(t_row, ra);
        tmem_ld_wait_regs(ra);
        tmem_ld32(t_row + 32, rb);
        process(ra, 0);
        tmem_ld_wait_regs(rb);
        tmem_ld32(t_row + 64, ra);
        process(rb, 1);
        tmem_ld_wait_regs(ra);
        tmem_ld32(t_row + 96, rb);
        process(ra, 2);
        tmem_ld_wait_regs(rb);
        // the last chunk of this half is in registers: hand the accumulator half back to the MMA issuer
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[h]);
        process(rb, 3);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&st_empty[s]);          // statistics and vectors of this tile have been consumed
    }
    if (lane == 0) tma_store_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem_base, 512);
}

}  // namespace aptai

namespace aptai {
// Launch on prepared inputs (called by aptai_conv0_norm_gelu, frontend.cu).  ws_b: 64 KB for the split operand;
// consts (norm 1) / affine (norm 2) as produced by frontend.cu's helper kernels.
int conv0_tc_launch(const float* wav, int B, int64_t L, int T0, const float* w, const float* bias,
                    const float* gamma, const float* beta, int norm, float eps, const float* consts, const float* affine,
                    void* ws_b, void* out16, int out_fp16, cudaStream_t st) {
  using namespace c0tc;
  __nv_bfloat16* bsplit = reinterpret_cast<__nv_bfloat16*>(ws_b);
  conv0_tc_prep_kernel<<<C0 / 128, 128, 0, st>>>(w, bias, norm != 2, bsplit);
  if (int rc = after_launch("conv0_tc_prep")) return rc;
  CUtensorMap tm;
  const uint64_t dims[3] = {static_cast<uint64_t>(C0), static_cast<uint64_t>(T0), static_cast<uint64_t>(B)};
  const uint64_t strides[2] = {static_cast<uint64_t>(C0) * 2, static_cast<uint64_t>(T0) * C0 * 2};
  const uint32_t box[3] = {64, 32, 1};
  if (int rc = encode_tmap_bf16(&tm, out16, 3, dims, strides, box, 1)) return rc;
  Conv0TcParams p;
  p.wav = wav; p.L = L; p.B = B; p.T0 = T0;
  p.tiles_per_utt = (T0 + TF - 1) / TF;
  p.num_tiles = B * p.tiles_per_utt;
  p.bsplit = bsplit; p.consts = consts; p.gamma = gamma; p.beta = beta; p.affine = affine;
  p.eps = eps; p.out_fp16 = out_fp16;
  p.x_tma = (L % 4 == 0 && (reinterpret_cast<uintptr_t>(wav) & 15) == 0) ? 1 : 0;
  const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  cudaError_t e = cudaSuccess;
#define APTAI_C0TC_LAUNCH(N)                                                                                      \
  do {                                                                                                            \
    e = cudaFuncSetAttribute(conv0_tc_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);        \
    if (e == cudaSuccess) conv0_tc_kernel<N><<<grid, THREADS, SMEM_BYTES, st>>>(tm, p);                           \
  } while (0)
  if (norm == 1) APTAI_C0TC_LAUNCH(1);
  else if (norm == 2) APTAI_C0TC_LAUNCH(2);
  else APTAI_C0TC_LAUNCH(0);
#undef APTAI_C0TC_LAUNCH
  if (e != cudaSuccess) {
    set_error("conv0_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return after_launch("conv0_tc");
}
}  // namespace aptai
