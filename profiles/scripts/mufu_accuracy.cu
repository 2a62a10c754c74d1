// Accuracy of the MUFU-based GELU candidates against the fp64 definition x * Phi(x) (profiles/scripts, not product code).
// nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/mufu_accuracy profiles/scripts/mufu_accuracy.cu && /tmp/mufu_accuracy
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__device__ float gelu_sig(float x) {          // current gelu_fast2 form: x / (1 + 2^(x q(min(x^2, 64))))
  float t = fminf(x * x, 64.f);
  float q = fmaf(t, 0.0010148165747523308f, -0.10677912831306458f);
  q = fmaf(q, t, -2.3011176586151123f);
  float a = q * x, e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
  return x * r;
}
__device__ float gelu_tanh1(float x) {        // one MUFU: 0.5 x (1 + tanh(-ln2/2 * x q))
  float t = fminf(x * x, 64.f);
  const float k = -0.34657359027997264f;      // -ln(2)/2
  float q = fmaf(t, 0.0010148165747523308f * k, -0.10677912831306458f * k);
  q = fmaf(q, t, -2.3011176586151123f * k);
  float a = q * x, th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(a));
  float hx = 0.5f * x;
  return fmaf(hx, th, hx);
}
__global__ void k(int n, float lo, float hi, double* out) {
  // out: [0] max abs err sig, [1] max abs err tanh, [2] sum sq sig, [3] sum sq tanh, [4] max rel-to-|x| tanh
  double m0 = 0, m1 = 0, s0 = 0, s1 = 0, m2 = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float x = lo + (hi - lo) * (float)i / (float)n;
    double ref = 0.5 * (double)x * erfc(-(double)x * 0.70710678118654752440);
    double e0 = fabs((double)gelu_sig(x) - ref), e1 = fabs((double)gelu_tanh1(x) - ref);
    m0 = fmax(m0, e0); m1 = fmax(m1, e1); s0 += e0 * e0; s1 += e1 * e1;
    if (fabsf(x) > 1e-3f) m2 = fmax(m2, e1 / fabs((double)x));
  }
  __shared__ double sh[5][256];
  sh[0][threadIdx.x] = m0; sh[1][threadIdx.x] = m1; sh[2][threadIdx.x] = s0; sh[3][threadIdx.x] = s1; sh[4][threadIdx.x] = m2;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int j = 1; j < 256; ++j) {
      sh[0][0] = fmax(sh[0][0], sh[0][j]); sh[1][0] = fmax(sh[1][0], sh[1][j]); sh[4][0] = fmax(sh[4][0], sh[4][j]);
      sh[2][0] += sh[2][j]; sh[3][0] += sh[3][j];
    }
    for (int q = 0; q < 5; ++q) out[blockIdx.x * 5 + q] = sh[q][0];
  }
}
int main() {
  double* d; cudaMalloc(&d, 5 * 64 * sizeof(double));
  const float ranges[4][2] = {{-8, 8}, {-3, 3}, {-1, 1}, {0, 4}};
  for (auto& r : ranges) {
    int n = 1 << 24;
    k<<<64, 256>>>(n, r[0], r[1], d);
    double h[5 * 64]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double m0 = 0, m1 = 0, s0 = 0, s1 = 0, m2 = 0;
    for (int b = 0; b < 64; ++b) { m0 = fmax(m0, h[b*5]); m1 = fmax(m1, h[b*5+1]); s0 += h[b*5+2]; s1 += h[b*5+3]; m2 = fmax(m2, h[b*5+4]); }
    printf("x in [%g,%g]: sigmoid form max abs %.3e rms %.3e | tanh form max abs %.3e rms %.3e max err/|x| %.3e\n", r[0], r[1], m0,
           sqrt(s0 / n), m1, sqrt(s1 / n), m2);
  }
  return 0;
}
