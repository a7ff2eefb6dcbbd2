// Log-mel front-end: waveform -> framed, windowed STFT magnitude -> mel filterbank ->
// log(clamp), the step before the re-encode path (reference convert_spectrograms.py:14-35, i.e.
// torchaudio MelSpectrogram(power=1) + log(clamp(min=1e-5))).
//
// HBM-bound byte work, so no GEMM: one CTA transforms TWO consecutive frames of one utterance as the
// real and imaginary part of a single complex FFT (Stockham radix-2 autosort in shared memory, fp32,
// twiddles from a host-computed table), separates the two spectra by conjugate symmetry, takes
// magnitudes, applies the mel filterbank in its sparse triangular form (each mel bin touches ~16
// frequency bins, not 1025) and writes log(max(mel, clip)).  The waveform is read once per frame
// overlap through L2 (hop * 4 bytes of new data per frame) and n_mels * 4 bytes are written per frame;
// the reflect padding of center=True is index arithmetic, never a padded copy.
#include <string.h>

#include "../../include/mqgan_b200.h"
#include "common.cuh"

namespace mq {

constexpr int kMelThreads = 256;

struct MelArgs {
  const float* wav;
  long long wav_ld;
  const long long* lengths;
  int n_fft, log2n, hop, n_mels, n_freqs;
  const float* window;
  const float2* twiddle;
  const int* fb_start;
  const int* fb_count;
  const int* fb_off;
  const float* fb_w;
  float clip;
  float* out;
  long long out_frames;
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

__global__ void __launch_bounds__(kMelThreads) log_mel_kernel(const MelArgs a) {
  extern __shared__ float2 sm[];
  const int N = a.n_fft;
  float2* buf0 = sm;
  float2* buf1 = sm + N;
  float* magA = reinterpret_cast<float*>(sm + 2 * N);
  float* magB = magA + ((a.n_freqs + 3) & ~3);
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const long long T = a.lengths[b];
  // torch.stft(center=True, pad_mode="reflect") needs more than n_fft/2 samples; shorter inputs have no frames
  const long long nfr = T > N / 2 ? 1 + T / a.hop : 0;
  const long long f0 = 2LL * blockIdx.x;
  if (f0 >= a.out_frames) return;
  float* orow = a.out + (static_cast<long long>(b) * a.out_frames + f0) * a.n_mels;
  const bool second = f0 + 1 < a.out_frames;
  if (f0 >= nfr) {                                   // rows past this utterance's last frame: zero fill
    for (int m = tid; m < a.n_mels; m += kMelThreads) {
      orow[m] = 0.0f;
      if (second) orow[a.n_mels + m] = 0.0f;
    }
    return;
  }
  const bool haveB = f0 + 1 < nfr;
  const float* w = a.wav + static_cast<long long>(b) * a.wav_ld;
  for (int n = tid; n < N; n += kMelThreads) {
    const float win = a.window[n];
    long long i = f0 * a.hop + n - N / 2;
    long long ia = i < 0 ? -i : i;
    if (ia >= T) ia = 2 * (T - 1) - ia;
    float xb = 0.0f;
    if (haveB) {
      long long ib = i + a.hop;
      if (ib < 0) ib = -ib;
      if (ib >= T) ib = 2 * (T - 1) - ib;
      xb = w[ib] * win;
    }
    buf0[n] = make_float2(w[ia] * win, xb);
  }
  __syncthreads();
  float2* in = buf0;
  float2* out = buf1;
  const int half = N >> 1;
  // Stockham autosort: radix-4 passes (sub-transform size Ns = 1, 4, 16, ...) and one radix-2 pass when
  // log2(n_fft) is odd.  Pass (Ns, R): thread j combines in[j + r N/R], twiddled by exp(-2 pi i k r / (R Ns))
  // with k = j mod Ns (twiddle[t] = exp(-2 pi i t / N), t < N), and writes out[(j - k) R + k + r Ns].
  int Ns = 1;
  for (; Ns * 4 <= N; Ns *= 4) {
    const int quarter = N >> 2;
    const int tstep = quarter / Ns;                   // k r / (4 Ns) = (k r tstep) / N
    for (int j = tid; j < quarter; j += kMelThreads) {
      const int k = j & (Ns - 1);
      const float2 v0 = in[j];
      const float2 v1 = cmul(in[j + quarter], __ldg(a.twiddle + k * tstep));
      const float2 v2 = cmul(in[j + 2 * quarter], __ldg(a.twiddle + 2 * k * tstep));
      const float2 v3 = cmul(in[j + 3 * quarter], __ldg(a.twiddle + 3 * k * tstep));
      const float2 a0 = make_float2(v0.x + v2.x, v0.y + v2.y), a1 = make_float2(v0.x - v2.x, v0.y - v2.y);
      const float2 a2 = make_float2(v1.x + v3.x, v1.y + v3.y);
      const float2 a3 = make_float2(v1.y - v3.y, v3.x - v1.x);          // (v1 - v3) * (-i)
      const int j0 = ((j - k) << 2) + k;
      out[j0] = make_float2(a0.x + a2.x, a0.y + a2.y);
      out[j0 + Ns] = make_float2(a1.x + a3.x, a1.y + a3.y);
      out[j0 + 2 * Ns] = make_float2(a0.x - a2.x, a0.y - a2.y);
      out[j0 + 3 * Ns] = make_float2(a1.x - a3.x, a1.y - a3.y);
    }
    __syncthreads();
    float2* tmp = in;
    in = out;
    out = tmp;
  }
  if (Ns < N) {                                       // one radix-2 pass left (Ns == N/2)
    for (int j = tid; j < half; j += kMelThreads) {
      const int k = j & (Ns - 1);
      const float2 v0 = in[j];
      const float2 t = cmul(in[j + half], __ldg(a.twiddle + k * (half / Ns)));
      const int j0 = ((j - k) << 1) + k;
      out[j0] = make_float2(v0.x + t.x, v0.y + t.y);
      out[j0 + Ns] = make_float2(v0.x - t.x, v0.y - t.y);
    }
    __syncthreads();
    float2* tmp = in;
    in = out;
    out = tmp;
  }
  // Z = FFT(xa + i xb):  A[k] = (Z[k] + conj(Z[N-k])) / 2,  B[k] = (Z[k] - conj(Z[N-k])) / (2i)
  for (int k = tid; k <= half; k += kMelThreads) {
    const float2 zk = in[k];
    const float2 zn = in[(N - k) & (N - 1)];
    const float ar = 0.5f * (zk.x + zn.x), ai = 0.5f * (zk.y - zn.y);
    const float br = 0.5f * (zk.y + zn.y), bi = -0.5f * (zk.x - zn.x);
    magA[k] = sqrtf(fmaf(ar, ar, ai * ai));
    magB[k] = sqrtf(fmaf(br, br, bi * bi));
  }
  __syncthreads();
  for (int m = tid; m < a.n_mels; m += kMelThreads) {
    const int s0 = a.fb_start[m], cnt = a.fb_count[m];
    const float* fw = a.fb_w + a.fb_off[m];
    float accA = 0.0f, accB = 0.0f;
    for (int i = 0; i < cnt; ++i) {
      const float wv = __ldg(fw + i);
      accA = fmaf(wv, magA[s0 + i], accA);
      accB = fmaf(wv, magB[s0 + i], accB);
    }
    orow[m] = logf(fmaxf(accA, a.clip));
    if (second) orow[a.n_mels + m] = haveB ? logf(fmaxf(accB, a.clip)) : 0.0f;
  }
}

}  // namespace mq

using namespace mq;

extern "C" int mq_log_mel(const mq_melspec_params* p, mq_stream_t stream) {
  MQ_REQUIRE(p != nullptr, "mq_log_mel: null params");
  MQ_REQUIRE(p->wav && p->lengths && p->window && p->twiddle && p->fb_start && p->fb_count && p->fb_off && p->fb_w && p->out,
             "mq_log_mel: null pointer argument");
  MQ_REQUIRE(p->B > 0 && p->B <= 65535 && p->hop > 0 && p->n_mels > 0 && p->out_frames > 0, "mq_log_mel: bad B/hop/n_mels/out_frames");
  int log2n = 0;
  while ((1 << log2n) < p->n_fft) ++log2n;
  MQ_REQUIRE((1 << log2n) == p->n_fft && p->n_fft >= 64 && p->n_fft <= 4096, "mq_log_mel: n_fft=%d must be a power of two in [64, 4096]", p->n_fft);
  MQ_REQUIRE(p->n_freqs == p->n_fft / 2 + 1, "mq_log_mel: n_freqs must be n_fft/2 + 1");
  MelArgs a;
  memset(&a, 0, sizeof(a));
  a.wav = p->wav; a.wav_ld = p->wav_ld; a.lengths = reinterpret_cast<const long long*>(p->lengths);
  a.n_fft = p->n_fft; a.log2n = log2n; a.hop = p->hop; a.n_mels = p->n_mels; a.n_freqs = p->n_freqs;
  a.window = p->window; a.twiddle = reinterpret_cast<const float2*>(p->twiddle);
  a.fb_start = p->fb_start; a.fb_count = p->fb_count; a.fb_off = p->fb_off; a.fb_w = p->fb_w;
  a.clip = p->clip; a.out = p->out; a.out_frames = p->out_frames;
  const size_t smem = 2 * static_cast<size_t>(p->n_fft) * sizeof(float2) + 2 * static_cast<size_t>((p->n_freqs + 3) & ~3) * sizeof(float);
  MQ_CUDA_OK(cudaFuncSetAttribute(log_mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  dim3 grid(static_cast<unsigned>((p->out_frames + 1) / 2), static_cast<unsigned>(p->B));
  log_mel_kernel<<<grid, kMelThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  MQ_CUDA_OK(cudaGetLastError());
  return 0;
}
