// Inverted dropout for the training path (HF:434 feat_proj_dropout, HF:546/603-607/647-653 hidden_dropout,
// HF:570 activation_dropout, HF:694/766 encoder dropout, models/aptai.py:44,52 head dropouts, models/w2v2_pr.py:56).
//
// Counter-based: the keep decision of element i of site s in step k is a pure function of (seed(s, k), i), so the
// backward pass regenerates the mask instead of storing it, and a test can materialise the very mask a step used.
// Standalone streaming kernels (HBM-bound) in this round; the sites sit right behind GEMM epilogues, fusing them in
// is the next step.
#include "common.h"
#include "ptx.cuh"

namespace aptai {

__device__ __forceinline__ bool dropout_keep(unsigned long long seed, unsigned long long idx, unsigned int thresh24) {
  unsigned long long x = idx + seed * 0x9E3779B97F4A7C15ULL;     // murmur3 fmix64 of a Weyl-shifted counter
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return (static_cast<unsigned int>(x) >> 8) >= thresh24;
}

// out = residual + keep(i) * x[i] / (1 - p)   for 4 consecutive elements per thread
template <bool IN_BF16>
__global__ void __launch_bounds__(256)
dropout_kernel(const void* __restrict__ xin, const float* __restrict__ residual, long long n4, unsigned int thresh24,
               float inv_keep, unsigned long long seed, float* __restrict__ out_f32,
               __nv_bfloat16* __restrict__ out_bf16) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 v;
    if (IN_BF16) {
      const uint2 u = __ldg(reinterpret_cast<const uint2*>(xin) + i);
      const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
      const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
      v = make_float4(a.x, a.y, b.x, b.y);
    } else {
      v = __ldg(reinterpret_cast<const float4*>(xin) + i);
    }
    const unsigned long long e = static_cast<unsigned long long>(i) * 4;
    v.x = dropout_keep(seed, e, thresh24) ? v.x * inv_keep : 0.f;
    v.y = dropout_keep(seed, e + 1, thresh24) ? v.y * inv_keep : 0.f;
    v.z = dropout_keep(seed, e + 2, thresh24) ? v.z * inv_keep : 0.f;
    v.w = dropout_keep(seed, e + 3, thresh24) ? v.w * inv_keep : 0.f;
    if (residual) {
      const float4 r = __ldg(reinterpret_cast<const float4*>(residual) + i);
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    if (out_f32) reinterpret_cast<float4*>(out_f32)[i] = v;
    if (out_bf16) reinterpret_cast<uint2*>(out_bf16)[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
}

}  // namespace aptai

using namespace aptai;

extern "C" int aptai_dropout(const void* x, int x_bf16, const float* residual, int64_t n, float p, uint64_t seed,
                             float* out_f32, void* out_bf16, void* stream) {
  if (int rc = check_arch()) return rc;
  APTAI_REQUIRE(x && (out_f32 || out_bf16) && n >= 4 && n % 4 == 0, "dropout: bad arguments (n must be a multiple of 4)");
  APTAI_REQUIRE(p >= 0.f && p < 1.f, "dropout: p must be in [0, 1)");
  const unsigned int thresh24 = static_cast<unsigned int>(static_cast<double>(p) * 16777216.0);
  const float inv_keep = 1.0f / (1.0f - p);
  const long long n4 = n / 4;
  long long gx = (n4 + 255) / 256;
  if (gx > 16LL * num_sms()) gx = 16LL * num_sms();
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  if (x_bf16) dropout_kernel<true><<<static_cast<unsigned>(gx), 256, 0, st>>>(x, residual, n4, thresh24, inv_keep, seed, out_f32, ob);
  else dropout_kernel<false><<<static_cast<unsigned>(gx), 256, 0, st>>>(x, residual, n4, thresh24, inv_keep, seed, out_f32, ob);
  return after_launch("dropout");
}
