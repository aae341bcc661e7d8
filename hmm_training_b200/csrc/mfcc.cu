// MFCC front-end: the 13 coefficients the reference stores per 20 ms frame,
//   librosa.feature.mfcc(y=frame, sr, n_mfcc=13, n_fft=len(frame), hop_length=None, center=False, n_mels=26)
// (CodeVector/codevector_classes.py:226-250), batched over frames of one length.
//
// librosa is not available in the build image, so this follows its published algorithm (0.11.0) as restated
// in oracle/mfcc_oracle.py — parity with librosa itself is UNPINNED:
//   periodic Hann window -> DFT of the n_fft = L samples (one frame, center=False) -> |X|^2 ->
//   Slaney mel filter bank (26 bands, 0 .. sr/2, float32 weights with Slaney area normalisation) ->
//   10 log10(max(1e-10, .)) clipped at (max - 80 dB) -> DCT-II (ortho) -> first 13 coefficients.
//
// One warp per frame.  The DFT of the real frame is evaluated directly in fp64 on its even / odd folded samples
// (lane = bins lane, lane+32, ...; twiddles from an L-entry shared table indexed by (k*m mod L), kept
// incrementally): ~2 * 161 * 160 FMAs per frame — the front-end moves 2.5 KB of samples per frame over PCIe.
#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"

namespace hmmb {

constexpr int MFCC_WARPS = 8;
constexpr int MFCC_MELS = 26;
constexpr int MFCC_COEF = HMMB_DIM;  // 13
constexpr int MFCC_MAX_L = 1024;
constexpr int MFCC_BPL_SMALL = 6;                               // bins per lane for L <= 320 (161 bins)
constexpr int MFCC_BPL_LARGE = (MFCC_MAX_L / 2 + 1 + 31) / 32;  // ... at the largest frame length

__global__ void k_mfcc_tables(int L, double *__restrict__ tcos, double *__restrict__ tsin) {
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < L; n += gridDim.x * blockDim.x) {
        double s, c;
        sincospi(2.0 * (double)n / (double)L, &s, &c);
        tcos[n] = c;
        tsin[n] = s;
    }
}

// melw [26][nb] float (nb = L/2+1), mrange [26][2] = first / one-past-last non-zero bin, dct [13][26]
template <int BPL>
__global__ void __launch_bounds__(MFCC_WARPS * 32)
k_mfcc_frames(const double *__restrict__ Y, int64_t F, int L, const double *__restrict__ tcos,
              const double *__restrict__ tsin, const float *__restrict__ melw, const int *__restrict__ mrange,
              const double *__restrict__ dct, double *__restrict__ out) {
    extern __shared__ double sm[];
    const int nb = L / 2 + 1;
    double *sCos = sm, *sSin = sm + L;
    double *sX = sm + 2 * L + (size_t)(threadIdx.x >> 5) * (L + nb + MFCC_MELS);  // per warp: frame, power, dB
    double *sP = sX + L, *sDb = sP + nb;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int n = threadIdx.x; n < L; n += blockDim.x) {
        sCos[n] = tcos[n];
        sSin[n] = tsin[n];
    }
    __syncthreads();
    for (int64_t f = (int64_t)blockIdx.x * MFCC_WARPS + warp; f < F; f += (int64_t)gridDim.x * MFCC_WARPS) {
        // windowed frame: y * (0.5 - 0.5 cos(2 pi n / L))  (periodic Hann, as scipy's get_window(fftbins=True))
        for (int n = lane; n < L; n += 32) sX[n] = Y[f * L + n] * (0.5 - 0.5 * sCos[n]);
        __syncwarp();
        // Real input: cos(2 pi k n / L) is even and sin odd under n -> L - n, so
        //   Re X_k = x_0 [+ (-1)^k x_{L/2}] + sum_{m=1}^{H-1} (x_m + x_{L-m}) cos(2 pi k m / L)
        //   Im X_k =                        - sum_{m=1}^{H-1} (x_m - x_{L-m}) sin(2 pi k m / L),   H = ceil(L / 2):
        // half the multiply-adds of the plain sum.  The folded samples replace x_m (even part) and x_{L-m} (odd part).
        const int H = (L + 1) / 2;
        for (int m = 1 + lane; m < H; m += 32) {
            const double a = sX[m], b = sX[L - m];
            sX[m] = a + b;
            sX[L - m] = a - b;
        }
        __syncwarp();
        // DFT bins k = lane + 32 j
        double re[BPL], im[BPL];
        int idx[BPL];
        const double x0 = sX[0], xh = (L & 1) ? 0.0 : sX[L / 2];
#pragma unroll
        for (int j = 0; j < BPL; ++j) {
            const int k = lane + 32 * j;
            re[j] = x0 + ((k & 1) ? -xh : xh);
            im[j] = 0.0;
            idx[j] = k;  // k * m mod L at m = 1 (k <= L / 2 < L)
        }
        for (int m = 1; m < H; ++m) {
            const double xe = sX[m], xo = sX[L - m];
#pragma unroll
            for (int j = 0; j < BPL; ++j) {
                const int k = lane + 32 * j;
                if (k < nb) {
                    re[j] = fma(xe, sCos[idx[j]], re[j]);
                    im[j] = fma(-xo, sSin[idx[j]], im[j]);
                    idx[j] += k;
                    if (idx[j] >= L) idx[j] -= L;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < BPL; ++j) {
            const int k = lane + 32 * j;
            if (k < nb) sP[k] = fma(re[j], re[j], im[j] * im[j]);
        }
        __syncwarp();
        // mel bands (lane = band), dB with the 80 dB floor below the frame's maximum
        double db = -INFINITY;
        if (lane < MFCC_MELS) {
            double acc = 0.0;
            const int k0 = mrange[2 * lane], k1 = mrange[2 * lane + 1];
            for (int k = k0; k < k1; ++k) acc = fma((double)melw[(size_t)lane * nb + k], sP[k], acc);
            db = 10.0 * log10(fmax(1e-10, acc));
        }
        double mx = db;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane < MFCC_MELS) sDb[lane] = fmax(db, mx - 80.0);
        __syncwarp();
        if (lane < MFCC_COEF) {
            double acc = 0.0;
#pragma unroll
            for (int m = 0; m < MFCC_MELS; ++m) acc = fma(dct[lane * MFCC_MELS + m], sDb[m], acc);
            out[f * MFCC_COEF + lane] = acc;
        }
        __syncwarp();
    }
}

// ---- host-side tables, following librosa.filters.mel / mel_frequencies step by step (Slaney scale)
static double hz_to_mel(double f) {
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz(double m) {
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}

static void mel_tables(double sr, int L, std::vector<float> &w, std::vector<int> &range) {
    const int nb = L / 2 + 1, nm = MFCC_MELS;
    std::vector<double> mel_f(nm + 2);
    const double lo = hz_to_mel(0.0), hi = hz_to_mel(sr / 2.0);
    // np.linspace(lo, hi, nm + 2): start + i * step, last point set to `hi` exactly
    const double step = (hi - lo) / (nm + 1);
    for (int i = 0; i < nm + 2; ++i) mel_f[i] = mel_to_hz(i == nm + 1 ? hi : lo + i * step);
    w.assign((size_t)nm * nb, 0.f);
    range.assign(2 * nm, 0);
    for (int i = 0; i < nm; ++i) {
        const double fd0 = mel_f[i + 1] - mel_f[i], fd1 = mel_f[i + 2] - mel_f[i + 1];
        const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
        int k0 = nb, k1 = 0;
        for (int k = 0; k < nb; ++k) {
            const double fk = (double)k * (1.0 / ((double)L * (1.0 / sr)));  // np.fft.rfftfreq(n, d = 1/sr) = k * (1 / (n d))
            const double lower = -(mel_f[i] - fk) / fd0, upper = (mel_f[i + 2] - fk) / fd1;
            float v = (float)std::max(0.0, std::min(lower, upper));  // stored into a float32 array
            v = (float)((double)v * enorm);                          // weights *= enorm (float32 result)
            w[(size_t)i * nb + k] = v;
            if (v != 0.f) { k0 = std::min(k0, k); k1 = std::max(k1, k + 1); }
        }
        range[2 * i] = k0 < k1 ? k0 : 0;
        range[2 * i + 1] = k0 < k1 ? k1 : 0;
    }
}

}  // namespace hmmb

using namespace hmmb;

extern "C" int hmmb_mfcc_frames(const double *Y, int64_t F, int L, int y_on_device, double sr, double *mfcc_out) {
    HMMB_TRY(require_init());
    if (F < 0 || !mfcc_out || (F > 0 && !Y) || !(sr > 0.0)) { set_error("hmmb_mfcc_frames: bad arguments"); return HMMB_ERR_ARG; }
    if (L < 2 || L > MFCC_MAX_L) { set_error("hmmb_mfcc_frames: frame length %d outside 2..%d", L, MFCC_MAX_L); return HMMB_ERR_UNSUPPORTED; }
    if (F == 0) return HMMB_OK;
    Ctx &c = ctx();
    const int nb = L / 2 + 1;
    std::vector<float> w;
    std::vector<int> range;
    mel_tables(sr, L, w, range);
    std::vector<double> dct((size_t)MFCC_COEF * MFCC_MELS);
    for (int k = 0; k < MFCC_COEF; ++k)
        for (int m = 0; m < MFCC_MELS; ++m) {
            // scipy.fftpack.dct(type=2, norm='ortho'): sqrt(2/N) cos(pi k (2m+1) / (2N)), row 0 scaled by sqrt(1/2)
            double v = std::cos(M_PI * k * (2 * m + 1) / (2.0 * MFCC_MELS)) * std::sqrt(2.0 / MFCC_MELS);
            if (k == 0) v *= std::sqrt(0.5);
            dct[(size_t)k * MFCC_MELS + m] = v;
        }
    struct Buf { void *p = nullptr; ~Buf() { dev_free(p); } } dY, dT, dW, dR, dD, dO;
    HMMB_TRY(dev_alloc(&dT.p, (size_t)2 * L * sizeof(double)));
    HMMB_TRY(dev_alloc(&dW.p, w.size() * sizeof(float)));
    HMMB_TRY(dev_alloc(&dR.p, range.size() * sizeof(int)));
    HMMB_TRY(dev_alloc(&dD.p, dct.size() * sizeof(double)));
    HMMB_TRY(dev_alloc(&dO.p, (size_t)F * MFCC_COEF * sizeof(double)));
    HMMB_CUDA(cudaMemcpyAsync(dW.p, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice, c.stream));
    HMMB_CUDA(cudaMemcpyAsync(dR.p, range.data(), range.size() * sizeof(int), cudaMemcpyHostToDevice, c.stream));
    HMMB_CUDA(cudaMemcpyAsync(dD.p, dct.data(), dct.size() * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    double *tcos = static_cast<double *>(dT.p), *tsin = tcos + L;
    HMMB_LAUNCH("mfcc", k_mfcc_tables, (L + 255) / 256, 256, 0, L, tcos, tsin);
    const size_t smem = ((size_t)2 * L + (size_t)MFCC_WARPS * (L + nb + MFCC_MELS)) * sizeof(double);
    const bool small = nb <= 32 * MFCC_BPL_SMALL;
    if (small) HMMB_CUDA(cudaFuncSetAttribute(k_mfcc_frames<MFCC_BPL_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else HMMB_CUDA(cudaFuncSetAttribute(k_mfcc_frames<MFCC_BPL_LARGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // frames [f0, f0 + n) of a device-resident sample matrix -> rows [f0, f0 + n) of dO
    auto launch = [&](const double *dYp, int64_t f0, int64_t n) -> int {
        const int64_t grid = std::min<int64_t>((n + MFCC_WARPS - 1) / MFCC_WARPS, (int64_t)c.sm_count * 8);
        double *o = static_cast<double *>(dO.p) + f0 * MFCC_COEF;
        if (small) {
            HMMB_LAUNCH("mfcc", k_mfcc_frames<MFCC_BPL_SMALL>, (unsigned)grid, MFCC_WARPS * 32, smem, dYp + f0 * L, n, L, tcos, tsin,
                        static_cast<const float *>(dW.p), static_cast<const int *>(dR.p), static_cast<const double *>(dD.p), o);
        } else {
            HMMB_LAUNCH("mfcc", k_mfcc_frames<MFCC_BPL_LARGE>, (unsigned)grid, MFCC_WARPS * 32, smem, dYp + f0 * L, n, L, tcos, tsin,
                        static_cast<const float *>(dW.p), static_cast<const int *>(dR.p), static_cast<const double *>(dD.p), o);
        }
        return HMMB_OK;
    };
    if (y_on_device) {
        HMMB_TRY(launch(Y, 0, F));
    } else {
        HMMB_TRY(dev_alloc(&dY.p, (size_t)F * L * sizeof(double)));
        const double *dYp = static_cast<const double *>(dY.p);
        const int64_t chunk = std::max<int64_t>(1, (int64_t(16) << 20) / ((int64_t)L * (int64_t)sizeof(double)));  // ~16 MB of samples
        if (host_is_pinned(Y) && F >= 2 * chunk) {
            // pinned samples go up in chunks on the copy stream and every chunk is transformed as soon as it has
            // landed: the kernel (30 M frames/s at L = 320) hides behind the PCIe transfer (21 M frames/s)
            cudaEvent_t fence = event_get();  // recycled device blocks may still be in use on the compute stream
            HMMB_CUDA(cudaEventRecord(fence, c.stream));
            HMMB_CUDA(cudaStreamWaitEvent(c.copy_stream, fence, 0));
            event_put(fence);
            for (int64_t f0 = 0; f0 < F; f0 += chunk) {
                const int64_t n = std::min<int64_t>(chunk, F - f0);
                HMMB_CUDA(cudaMemcpyAsync(static_cast<double *>(dY.p) + f0 * L, Y + f0 * L, (size_t)n * L * sizeof(double),
                                          cudaMemcpyHostToDevice, c.copy_stream));
                cudaEvent_t landed = event_get();
                HMMB_CUDA(cudaEventRecord(landed, c.copy_stream));
                HMMB_CUDA(cudaStreamWaitEvent(c.stream, landed, 0));
                event_put(landed);
                HMMB_TRY(launch(dYp, f0, n));
            }
        } else {
            HMMB_TRY(h2d_big(dY.p, Y, (size_t)F * L * sizeof(double), c.stream));
            HMMB_TRY(launch(dYp, 0, F));
        }
    }
    HMMB_TRY(d2h_big(mfcc_out, dO.p, (size_t)F * MFCC_COEF * sizeof(double), c.stream));
    return HMMB_OK;
}
