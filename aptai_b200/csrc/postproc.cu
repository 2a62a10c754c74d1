// The callers' data formats either side of the hot path, on the device:
//   input side  — ragged -> padded collate (train/train_aptai.py:268-332), polyphase sinc resampling to 16 kHz
//                 (torchaudio.functional.resample as used by data/dataset_hprc.py:68-72), linear interpolation of
//                 the articulatory targets to the 49 Hz frame rate (data/dataset_hprc.py:2307-2313);
//   output side — frame labels -> (start, end, phoneme) segments (utility.py:539-566), per-channel RMSE / Pearson
//                 of the trajectories (utility.py:393-444), boundary precision/recall counters (utility.py:589-611),
//                 frame overlap (utility.py:614-622).
// All of it is HBM- or latency-bound integer / fp64 bookkeeping: one thread or one warp per utterance or channel,
// sequential fp64 sums where the reference's Python `sum()` is sequential (so RMSE is bit-exact).
#include "common.h"

#include <math.h>

namespace aptai {

// ---------------------------------------------------------------------------------------------- collate
template <typename T>
__global__ void collate_pad_kernel(const T* __restrict__ flat, const long long* __restrict__ offsets, int B,
                                   long long Lmax, T pad, T* __restrict__ out) {
  const int b = blockIdx.y;
  const long long o0 = offsets[b], n = offsets[b + 1] - o0;
  T* row = out + static_cast<long long>(b) * Lmax;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < Lmax;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    row[i] = i < n ? flat[o0 + i] : pad;
}

// ---------------------------------------------------------------------------------------------- resample
// y[b][j] = sum_k kern[j % new][k] * xpad[b][(j / new) * orig + k],  xpad[n] = x[n - width] (zero outside [0, len))
__global__ void resample_fir_kernel(const float* __restrict__ x, const long long* __restrict__ in_len, long long in_ld,
                                    const float* __restrict__ kern, int orig, int nw, int width, int klen,
                                    float* __restrict__ y, long long out_ld) {
  const int b = blockIdx.y;
  const long long L = in_len[b];
  const long long target = (static_cast<long long>(nw) * L + orig - 1) / orig;     // ceil(new * L / orig)
  const float* xb = x + static_cast<long long>(b) * in_ld;
  float* yb = y + static_cast<long long>(b) * out_ld;
  for (long long j = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; j < out_ld;
       j += static_cast<long long>(gridDim.x) * blockDim.x) {
    float acc = 0.f;
    if (j < target) {
      const int p = static_cast<int>(j % nw);
      const long long base = (j / nw) * orig - width;
      const float* kp = kern + static_cast<long long>(p) * klen;
      for (int k = 0; k < klen; ++k) {
        const long long n = base + k;
        if (n >= 0 && n < L) acc = fmaf(__ldg(kp + k), __ldg(xb + n), acc);
      }
    }
    yb[j] = acc;
  }
}

// ---------------------------------------------------------------------------------------------- linear interpolation
// scipy.interpolate.interp1d(arange(n), sig, 'linear', axis=0)(linspace(0, n-1, m)), fp64, same operation order
// per_channel = 1: every column is interpolated as a 1-D signal, where scipy dispatches to numpy.interp
// (slope * (x - x_lo) + y_lo, exact grid hits returned as is) — the reference's call pattern (dataset_hprc.py:2370)
__global__ void interp_linear_kernel(const double* __restrict__ sig, int n, int C, int m, double step, int per_channel,
                                     double* __restrict__ out) {
  const long long total = static_cast<long long>(m) * C;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / C), c = static_cast<int>(i - static_cast<long long>(r) * C);
    const double xn = (r == m - 1 && m > 1) ? static_cast<double>(n - 1) : __dmul_rn(static_cast<double>(r), step);
    if (per_channel) {
      const int j = min(static_cast<int>(floor(xn)), n - 2);
      const double yj = sig[static_cast<long long>(j) * C + c];
      if (xn == static_cast<double>(n - 1)) out[i] = sig[static_cast<long long>(n - 1) * C + c];
      else if (xn == static_cast<double>(j)) out[i] = yj;
      else
        out[i] = __dadd_rn(__dmul_rn(__ddiv_rn(__dsub_rn(sig[static_cast<long long>(j + 1) * C + c], yj), 1.0),
                                     __dsub_rn(xn, static_cast<double>(j))), yj);
      continue;
    }
    int hi = static_cast<int>(ceil(xn));            // searchsorted(arange(n), xn, 'left')
    hi = min(max(hi, 1), n - 1);
    const int lo = hi - 1;
    const double ylo = sig[static_cast<long long>(lo) * C + c], yhi = sig[static_cast<long long>(hi) * C + c];
    // scipy _call_linear: (x_new - x_lo)/(x_hi - x_lo) * y_hi + (x_hi - x_new)/(x_hi - x_lo) * y_lo, x_hi - x_lo = 1
    const double whi = __dsub_rn(xn, static_cast<double>(lo)), wlo = __dsub_rn(static_cast<double>(hi), xn);
    out[i] = __dadd_rn(__dmul_rn(whi, yhi), __dmul_rn(wlo, ylo));
  }
}

// ---------------------------------------------------------------------------------------------- segments (RLE)
// one warp per utterance: run starts found with ballots, written in order
__global__ void frames_to_segments_kernel(const long long* __restrict__ frames, const int* __restrict__ lens, int B,
                                          int T, int* __restrict__ seg_start, int* __restrict__ seg_end,
                                          long long* __restrict__ seg_phn, int* __restrict__ nseg, int max_seg) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int n = min(max(lens[b], 0), T);
  const long long* f = frames + static_cast<long long>(b) * T;
  int count = 0;
  for (int t0 = 0; t0 < n; t0 += 32) {
    const int t = t0 + lane;
    const bool start = t < n && (t == 0 || f[t] != f[t - 1]);
    const unsigned m = __ballot_sync(0xffffffffu, start);
    if (start) {
      const int idx = count + __popc(m & ((1u << lane) - 1));
      if (idx < max_seg) {
        seg_start[static_cast<long long>(b) * max_seg + idx] = t;
        seg_phn[static_cast<long long>(b) * max_seg + idx] = f[t];
        if (idx > 0) seg_end[static_cast<long long>(b) * max_seg + idx - 1] = t;
      }
    }
    count += __popc(m);
  }
  if (lane == 0) {
    if (count > 0 && count <= max_seg) seg_end[static_cast<long long>(b) * max_seg + count - 1] = n;
    nseg[b] = count;
  }
}

// ---------------------------------------------------------------------------------------------- TV metrics
// one thread per (utterance, channel): sequential fp64 sums over the valid frames, like the reference's Python
// `sum(se) / len(se)`; Pearson as scipy.stats.pearsonr (centred vectors normalised by their 2-norms).
__global__ void tv_metrics_kernel(const float* __restrict__ gt, const float* __restrict__ pred,
                                  const int* __restrict__ lens, int B, int T, int C, double* __restrict__ rmse,
                                  double* __restrict__ pcc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, c = i - b * C;
  const int n = min(max(lens[b], 0), T);
  const float* g = gt + static_cast<long long>(b) * T * C + c;
  const float* p = pred + static_cast<long long>(b) * T * C + c;
  double se = 0.0, sg = 0.0, sp = 0.0;
  for (int t = 0; t < n; ++t) {
    const double a = g[static_cast<long long>(t) * C], q = p[static_cast<long long>(t) * C];
    const double d = __dsub_rn(a, q);
    se = __dadd_rn(se, __dmul_rn(d, d));
    sg += a;
    sp += q;
  }
  rmse[i] = n > 0 ? sqrt(se / n) : nan("");
  const double mg = sg / n, mp = sp / n;
  double gg = 0.0, pp = 0.0, gp = 0.0;
  for (int t = 0; t < n; ++t) {
    const double a = g[static_cast<long long>(t) * C] - mg, q = p[static_cast<long long>(t) * C] - mp;
    gg += a * a;
    pp += q * q;
    gp += a * q;
  }
  double r = gp / (sqrt(gg) * sqrt(pp));
  r = fmax(fmin(r, 1.0), -1.0);
  pcc[i] = n > 1 ? r : nan("");
}

// ---------------------------------------------------------------------------------------------- boundary statistics
// per utterance: precision_counter = #{yhat_i : min_j |y_j - yhat_i| <= tol}, recall_counter likewise (utility.py:589-611)
__global__ void boundary_stats_kernel(const double* __restrict__ y, const int* __restrict__ ny,
                                      const double* __restrict__ yhat, const int* __restrict__ nyhat, int B, int maxn,
                                      double tol, int* __restrict__ counters /* [B][4]: prec, rec, n_pred, n_gt */) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const double* yb = y + static_cast<long long>(b) * maxn;
  const double* hb = yhat + static_cast<long long>(b) * maxn;
  const int n = ny[b], m = nyhat[b];
  int pc = 0, rc = 0;
  for (int i = lane; i < m; i += 32) {
    double md = INFINITY;
    for (int j = 0; j < n; ++j) md = fmin(md, fabs(__dsub_rn(yb[j], hb[i])));
    pc += (md <= tol);
  }
  for (int j = lane; j < n; j += 32) {
    double md = INFINITY;
    for (int i = 0; i < m; ++i) md = fmin(md, fabs(__dsub_rn(hb[i], yb[j])));
    rc += (md <= tol);
  }
  for (int o = 16; o; o >>= 1) {
    pc += __shfl_xor_sync(0xffffffffu, pc, o);
    rc += __shfl_xor_sync(0xffffffffu, rc, o);
  }
  if (lane == 0) {
    counters[b * 4 + 0] = pc;
    counters[b * 4 + 1] = rc;
    counters[b * 4 + 2] = m;
    counters[b * 4 + 3] = n;
  }
}

// hits / counts of equal frame labels over the valid frames (utility.py:614-622)
__global__ void frame_overlap_kernel(const long long* __restrict__ a, const long long* __restrict__ b_,
                                     const int* __restrict__ lens, int B, int T, unsigned long long* __restrict__ hc) {
  unsigned long long hits = 0, cnt = 0;
  const long long total = static_cast<long long>(B) * T;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / T), t = static_cast<int>(i - static_cast<long long>(b) * T);
    if (t < lens[b]) {
      hits += (a[i] == b_[i]);
      ++cnt;
    }
  }
  for (int o = 16; o; o >>= 1) {
    hits += __shfl_xor_sync(0xffffffffu, hits, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(hc, hits);
    atomicAdd(hc + 1, cnt);
  }
}

}  // namespace aptai

using namespace aptai;

extern "C" int aptai_collate_pad(const void* flat, int elem_bytes, const int64_t* offsets, int B, int64_t Lmax,
                                 const void* pad_value_host, void* out, void* stream) {
  if (int rc = check_arch()) return rc;
  APTAI_REQUIRE(flat && offsets && out && pad_value_host && B >= 1 && Lmax >= 1, "collate_pad: bad arguments");
  APTAI_REQUIRE(elem_bytes == 4 || elem_bytes == 8, "collate_pad: element size must be 4 or 8 bytes");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int gx = static_cast<int>((Lmax + 255) / 256);
  if (gx > 2048) gx = 2048;
  dim3 grid(gx, B);
  if (elem_bytes == 4)
    collate_pad_kernel<uint32_t><<<grid, 256, 0, st>>>(reinterpret_cast<const uint32_t*>(flat),
                                                       reinterpret_cast<const long long*>(offsets), B, Lmax,
                                                       *reinterpret_cast<const uint32_t*>(pad_value_host),
                                                       reinterpret_cast<uint32_t*>(out));
  else
    collate_pad_kernel<unsigned long long><<<grid, 256, 0, st>>>(
        reinterpret_cast<const unsigned long long*>(flat), reinterpret_cast<const long long*>(offsets), B, Lmax,
        *reinterpret_cast<const unsigned long long*>(pad_value_host), reinterpret_cast<unsigned long long*>(out));
  return after_launch("collate_pad");
}

extern "C" int aptai_resample_fir(const float* x, const int64_t* in_len, int B, int64_t in_ld, const float* kernel,
                                  int orig, int nw, int width, float* y, int64_t out_ld, void* stream) {
  if (int rc = check_arch()) return rc;
  APTAI_REQUIRE(x && in_len && kernel && y && B >= 1 && orig >= 1 && nw >= 1 && width >= 0, "resample_fir: bad arguments");
  const int klen = 2 * width + orig;
  int gx = static_cast<int>((out_ld + 255) / 256);
  if (gx > 4096) gx = 4096;
  resample_fir_kernel<<<dim3(gx, B), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, reinterpret_cast<const long long*>(in_len), in_ld, kernel, orig, nw, width, klen, y, out_ld);
  return after_launch("resample_fir");
}

extern "C" int aptai_interp_linear_f64(const double* sig, int n, int C, int m, int per_channel, double* out,
                                       void* stream) {
  if (int rc = check_arch()) return rc;
  APTAI_REQUIRE(sig && out && n >= 2 && C >= 1 && m >= 1, "interp_linear: need at least 2 input rows");
  const double step = m > 1 ? static_cast<double>(n - 1) / static_cast<double>(m - 1) : 0.0;   // numpy.linspace
  const long long total = static_cast<long long>(m) * C;
  int gx = static_cast<int>((total + 255) / 256);
  if (gx > 4096) gx = 4096;
  interp_linear_kernel<<<gx, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(sig, n, C, m, step, per_channel, out);
  return after_launch("interp_linear");
}

extern "C" int aptai_frames_to_segments(const int64_t* frames, const int32_t* lens, int B, int T, int32_t* seg_start,
                                        int32_t* seg_end, int64_t* seg_phn, int32_t* nseg, int max_seg, void* stream) {
  if (int rc = check_arch()) return rc;
  APTAI_REQUIRE(frames && lens && seg_start && seg_end && seg_phn && nseg && B >= 1 && T >= 1 && max_seg >= 1,
                "frames_to_segments: bad arguments");
  frames_to_segments_kernel<<<(B + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(frames), lens, B, T, seg_start, seg_end,
      reinterpret_cast<long long*>(seg_phn), nseg, max_seg);
  return after_launch("frames_to_segments");
}

extern "C" int aptai_tv_metrics(const float* gt, const float* pred, const int32_t* lens, int B, int T, int C,
                                double* rmse, double* pcc, void* stream) {
  if (int rc = check_arch()) return rc;
  APTAI_REQUIRE(gt && pred && lens && rmse && pcc && B >= 1 && T >= 1 && C >= 1, "tv_metrics: bad arguments");
  tv_metrics_kernel<<<(B * C + 63) / 64, 64, 0, reinterpret_cast<cudaStream_t>(stream)>>>(gt, pred, lens, B, T, C, rmse,
                                                                                          pcc);
  return after_launch("tv_metrics");
}

extern "C" int aptai_boundary_stats(const double* y, const int32_t* ny, const double* yhat, const int32_t* nyhat, int B,
                                    int maxn, double tolerance, int32_t* counters, void* stream) {
  if (int rc = check_arch()) return rc;
  APTAI_REQUIRE(y && ny && yhat && nyhat && counters && B >= 1 && maxn >= 1, "boundary_stats: bad arguments");
  boundary_stats_kernel<<<(B + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(y, ny, yhat, nyhat, B, maxn,
                                                                                         tolerance, counters);
  return after_launch("boundary_stats");
}

extern "C" int aptai_frame_overlap(const int64_t* a, const int64_t* b, const int32_t* lens, int B, int T,
                                   uint64_t* hits_counts, void* stream) {
  if (int rc = check_arch()) return rc;
  APTAI_REQUIRE(a && b && lens && hits_counts && B >= 1 && T >= 1, "frame_overlap: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(hits_counts, 0, 2 * sizeof(uint64_t), st);
  if (e != cudaSuccess) {
    set_error("frame_overlap: memset: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  const long long total = static_cast<long long>(B) * T;
  int gx = static_cast<int>((total + 255) / 256);
  if (gx > 1024) gx = 1024;
  frame_overlap_kernel<<<gx, 256, 0, st>>>(reinterpret_cast<const long long*>(a), reinterpret_cast<const long long*>(b),
                                           lens, B, T, reinterpret_cast<unsigned long long*>(hits_counts));
  return after_launch("frame_overlap");
}
