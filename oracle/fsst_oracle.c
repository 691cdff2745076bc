/* CPU oracle (C restatement) of the FSST hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load the library built from this file.  The product never links or calls it.
 *
 * PARITY UNPINNED: the reference's arithmetic is ssq.fsst (reference
 * hss/transforms/synchrosqueeze.py:48), i.e. the external libssq 0.1.0 (MATLAB-Coder C++ of
 * MATLAB's fsst, FFTW backed; reference pixi.lock:2598-2607,4609-4620) whose source is not in the
 * reference tree.  This file restates the published MATLAB fsst algorithm (see
 * oracle/fsst_oracle.py for the step list) in float64 and is validated against that numpy
 * restatement and the analytic KATs in tests/test_oracle_fsst.py.
 *
 * It is also the "port" CPU baseline timed by bench.py (OpenMP over windows x time blocks).
 *
 * Build: make -C oracle   ->  oracle/_build/libhss_oracle.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define HSSO_MAX_NFFT 4096

typedef struct { double re, im; } cplx;

/* iterative radix-2 DIT FFT, n power of two, twiddles tw[j] = exp(-2*pi*i*j/n), j < n/2 */
static void fft_pow2(cplx *a, int n, const cplx *tw)
{
    for (int i = 1, j = 0; i < n; ++i) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { cplx t = a[i]; a[i] = a[j]; a[j] = t; }
    }
    for (int len = 2; len <= n; len <<= 1) {
        int half = len >> 1, step = n / len;
        for (int i = 0; i < n; i += len) {
            for (int j = 0; j < half; ++j) {
                cplx w = tw[j * step];
                cplx u = a[i + j], v = a[i + j + half];
                cplx vw = { v.re * w.re - v.im * w.im, v.re * w.im + v.im * w.re };
                a[i + j].re = u.re + vw.re;        a[i + j].im = u.im + vw.im;
                a[i + j + half].re = u.re - vw.re; a[i + j + half].im = u.im - vw.im;
            }
        }
    }
}

/* direct DFT for non power-of-two window lengths (slow; parity cases only) */
static void dft_any(const cplx *in, cplx *out, int n, const cplx *tw_full)
{
    for (int k = 0; k < n; ++k) {
        double sr = 0.0, si = 0.0;
        for (int j = 0; j < n; ++j) {
            cplx w = tw_full[(int)(((int64_t)k * j) % n)];
            sr += in[j].re * w.re - in[j].im * w.im;
            si += in[j].re * w.im + in[j].im * w.re;
        }
        out[k].re = sr; out[k].im = si;
    }
}

static double round_half_away(double v) { return v < 0.0 ? -floor(-v + 0.5) : floor(v + 0.5); }

/* Thread count of the OpenMP loops over windows; n <= 0 leaves it alone.  (Under torchrun the environment carries
 * OMP_NUM_THREADS=1: the bench sets the count explicitly instead of inheriting it.) */
void hsso_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int hsso_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Restated ssq.fsst for a batch: x[B,N] float32 (what the reference feeds, synchrosqueeze.py:48 via
 * preprocess.py:31-32), g/dg float64[nwin]; s_out[B,K,N] interleaved (re,im) float64, K = nwin/2+1.
 * Steps 1-7 of MATLAB fsst.  Returns 0, or -1 on bad arguments. */
int hsso_fsst(const float *x, int64_t B, int64_t N, double fs, const double *g, const double *dg,
              int nwin, double *s_out)
{
    if (!x || !g || !dg || !s_out || nwin < 4 || nwin > HSSO_MAX_NFFT || B < 0 || N < 0) return -1;
    const int nfft = nwin, K = nfft / 2 + 1;
    const int pow2 = (nfft & (nfft - 1)) == 0;
    const int left = (nwin % 2) ? (nwin - 1) / 2 : nwin / 2;
    cplx *tw = (cplx *)malloc(sizeof(cplx) * nfft);
    cplx *ez = (cplx *)malloc(sizeof(cplx) * nfft);
    for (int j = 0; j < nfft; ++j) {
        double a = -2.0 * M_PI * j / nfft;
        tw[j].re = cos(a); tw[j].im = sin(a);
        if (nwin % 2 == 0) { ez[j].re = (j % 2) ? -1.0 : 1.0; ez[j].im = 0.0; }
        else { double p = -2.0 * M_PI * (nwin / 2) * (double)j / nfft; ez[j].re = cos(p); ez[j].im = sin(p); }
    }
    const double df = fs / nfft;
    const double fmin = 0.0, fmax = (nfft - 1) * df;
    memset(s_out, 0, sizeof(double) * 2 * (size_t)B * K * N);

#pragma omp parallel
    {
        cplx *z = (cplx *)malloc(sizeof(cplx) * nfft);
        cplx *zz = (cplx *)malloc(sizeof(cplx) * nfft);
#pragma omp for collapse(2) schedule(static)
        for (int64_t b = 0; b < B; ++b) {
            for (int64_t t0 = 0; t0 < N; t0 += 256) {
                int64_t t1 = t0 + 256 < N ? t0 + 256 : N;
                const float *xb = x + b * N;
                double *sb = s_out + 2 * (size_t)b * K * N;
                for (int64_t t = t0; t < t1; ++t) {
                    /* frame t = xp[t .. t+nwin-1], xp = [zeros(left); x; zeros(right)] */
                    for (int j = 0; j < nwin; ++j) {
                        int64_t idx = t + j - left;
                        double v = (idx >= 0 && idx < N) ? (double)xb[idx] : 0.0;
                        z[j].re = v * g[j];      /* packs both real FFTs into one complex FFT */
                        z[j].im = v * dg[j];
                    }
                    cplx *Z = z;
                    if (pow2) fft_pow2(z, nfft, tw); else { dft_any(z, zz, nfft, tw); Z = zz; }
                    for (int k = 0; k < nfft; ++k) {
                        int km = (nfft - k) % nfft;
                        /* split: Sg = (Z[k] + conj Z[-k])/2 ; Sdg = (Z[k] - conj Z[-k])/(2i) */
                        double gr = 0.5 * (Z[k].re + Z[km].re), gi = 0.5 * (Z[k].im - Z[km].im);
                        double dr = 0.5 * (Z[k].im + Z[km].im), di = -0.5 * (Z[k].re - Z[km].re);
                        /* fcorr = -imag(Sdg / Sg); non-finite -> 0 */
                        double den = gr * gr + gi * gi;
                        double fc = -((di * gr - dr * gi) / den);
                        if (!isfinite(fc)) fc = 0.0;
                        double inst = k * df + fc;
                        double r = round_half_away((inst - fmin) * (nfft - 1) / (fmax - fmin));
                        double m = fmod(r, (double)nfft);
                        if (m < 0) m += nfft;
                        int row = (int)m;
                        if (row < K) {
                            double vr = gr * ez[k].re - gi * ez[k].im;
                            double vi = gr * ez[k].im + gi * ez[k].re;
                            sb[2 * ((size_t)row * N + t)] += vr;
                            sb[2 * ((size_t)row * N + t) + 1] += vi;
                        }
                    }
                }
            }
        }
        free(z); free(zz);
    }
    free(tw); free(ez);
    return 0;
}

/* Restated FSST.__call__ post-processing (reference synchrosqueeze.py:50-89) on top of hsso_fsst.
 * mode 1: abs -> out float32 [B,N,Kt];  mode 2: stack -> out float32 [B,N,2*Kt].
 * s is rounded to complex64 first (synchrosqueeze.py:51); mean / unbiased std in float64. */
int hsso_fsst_features(const float *x, int64_t B, int64_t N, double fs, const double *g,
                       const double *dg, int nwin, int k_lo, int k_hi, int mode, float *out)
{
    const int K = nwin / 2 + 1;
    if (k_lo < 0 || k_hi >= K || k_hi < k_lo || (mode != 1 && mode != 2)) return -1;
    const int Kt = k_hi - k_lo + 1;
    /* bounded scratch: process the batch in chunks of windows */
    const int64_t chunk = 32;
    if (B > chunk) {
        for (int64_t b0 = 0; b0 < B; b0 += chunk) {
            int64_t nb = B - b0 < chunk ? B - b0 : chunk;
            int rc = hsso_fsst_features(x + b0 * N, nb, N, fs, g, dg, nwin, k_lo, k_hi, mode,
                                        out + (size_t)b0 * N * (mode == 1 ? Kt : 2 * Kt));
            if (rc) return rc;
        }
        return 0;
    }
    double *s = (double *)malloc(sizeof(double) * 2 * (size_t)B * K * N);
    if (!s) return -2;
    int rc = hsso_fsst(x, B, N, fs, g, dg, nwin, s);
    if (rc) { free(s); return rc; }
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; ++b) {
        const double *sb = s + 2 * (size_t)b * K * N;
        if (mode == 1) {
            for (int k = 0; k < Kt; ++k)
                for (int64_t t = 0; t < N; ++t) {
                    float re = (float)sb[2 * ((size_t)(k + k_lo) * N + t)];
                    float im = (float)sb[2 * ((size_t)(k + k_lo) * N + t) + 1];
                    out[((size_t)b * N + t) * Kt + k] = (float)hypot((double)re, (double)im);
                }
            continue;
        }
        double mean[2] = {0, 0}, m2[2] = {0, 0};
        int64_t cnt = 0;
        /* streaming mean / M2: the hss.moments recurrences (reference hss/moments/__init__.py:16,35-36) */
        for (int k = 0; k < Kt; ++k)
            for (int64_t t = 0; t < N; ++t) {
                ++cnt;
                for (int c = 0; c < 2; ++c) {
                    double v = (double)(float)sb[2 * ((size_t)(k + k_lo) * N + t) + c];
                    double delta = v - mean[c];
                    mean[c] += delta / (double)cnt;
                    m2[c] += delta * (v - mean[c]);
                }
            }
        double sd[2] = { sqrt(m2[0] / (double)(cnt - 1)), sqrt(m2[1] / (double)(cnt - 1)) };
        for (int k = 0; k < Kt; ++k)
            for (int64_t t = 0; t < N; ++t)
                for (int c = 0; c < 2; ++c) {
                    double v = (double)(float)sb[2 * ((size_t)(k + k_lo) * N + t) + c];
                    out[((size_t)b * N + t) * (2 * Kt) + c * Kt + k] = (float)((v - mean[c]) / sd[c]);
                }
    }
    free(s);
    return 0;
}
