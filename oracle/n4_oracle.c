/* Oracle (TEST INFRASTRUCTURE): CPU restatement of N4 bias-field correction as executed
 * by the reference's only call site, Vent_Analysis.py:316-334:
 *
 *     corrector = sitk.N4BiasFieldCorrectionImageFilter()      # all defaults
 *     corrected_image = corrector.Execute(image, mask)         # float32 image, uint8 mask
 *
 * **PARITY UNPINNED.**  The arithmetic lives in the third-party dependency
 * SimpleITK==2.3.1 (reference requirements.txt:12), a binary wheel wrapping ITK 5.3
 * (itkN4BiasFieldCorrectionImageFilter.hxx, itkBSplineScatteredDataPointSetToImageFilter.hxx,
 * itkBSplineControlPointImageFilter.hxx, itkBSplineKernelFunction.h, vnl_fft_1d).  It is not
 * vendored under /root/reference, not installed in the build image and there is no network, so
 * this file restates ITK's published algorithm (Tustison et al., IEEE TMI 29(6) 2010, and the
 * ITK sources named above) and could not be checked against SimpleITK output here.  The
 * reference has no tests or golden vectors for N4.  tests/test_n4_simpleitk_optin.py pins it
 * wherever SimpleITK is importable.
 *
 * Arithmetic follows ITK type by type: RealType = float for images, lattices, B-spline
 * weights and the convergence statistics; the B-spline kernel evaluates in double; the
 * histogram deconvolution runs in double complex.  Single thread, voxels visited in ITK
 * iterator order (= numpy C order of the [row, col, slice] array, because
 * sitk.GetImageFromArray maps numpy axis 2 -> ITK x, axis 1 -> y, axis 0 -> z and ITK
 * iterates x fastest).  ITK itself splits the lattice accumulation over threads, so its own
 * float sums depend on the thread count; this restatement is the 1-thread order.
 *
 * SimpleITK 2.3.1 defaults: ConvergenceThreshold 0.001, MaximumNumberOfIterations
 * [50,50,50,50] (=> 4 fitting levels), BiasFieldFullWidthAtHalfMaximum 0.15,
 * WienerFilterNoise 0.01, NumberOfHistogramBins 200, NumberOfControlPoints [4,4,4],
 * SplineOrder 3, UseMaskLabel true, MaskLabel 1.  Spacing is (1,1,1), origin 0
 * (Vent_Analysis.py:322-323 never set them).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXCP 64          /* max control points per axis handled (4 -> 5 -> 7 -> 11 by default) */
#define BSPLINE_EPS 1e-3f /* m_BSplineEpsilon in both ITK B-spline filters */

/* itk::BSplineKernelFunction<3>::Evaluate (double) */
static double bspline3(double u)
{
    const double a = fabs(u);
    if (a < 1.0) {
        const double s = a * a;
        return (4.0 - 6.0 * s + 3.0 * s * a) / 6.0;
    } else if (a < 2.0) {
        const double s = a * a;
        return (8.0 - 12.0 * a + 6.0 * s - s * a) / 6.0;
    }
    return 0.0;
}

/* unnormalised radix-2 complex FFT (vnl_fft_1d is unnormalised both ways; the scale cancels
 * in E = numerator / denominator). sign = -1 forward, +1 backward. */
static void fft_pow2(double *re, double *im, int n, int sign)
{
    for (int i = 1, j = 0; i < n; ++i) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) {
            double t = re[i]; re[i] = re[j]; re[j] = t;
            t = im[i]; im[i] = im[j]; im[j] = t;
        }
    }
    for (int len = 2; len <= n; len <<= 1) {
        const double ang = sign * 2.0 * M_PI / len;
        for (int i = 0; i < n; i += len)
            for (int k = 0; k < len / 2; ++k) {
                const double wr = cos(ang * k), wi = sin(ang * k);
                const int a = i + k, b = i + k + len / 2;
                const double xr = re[b] * wr - im[b] * wi, xi = re[b] * wi + im[b] * wr;
                re[b] = re[a] - xr; im[b] = im[a] - xi;
                re[a] += xr; im[a] += xi;
            }
    }
}

/* N4BiasFieldCorrectionImageFilter::SharpenImage.  U: log-uncorrected image, m: mask==label.
 * Writes S (sharpened) on masked voxels only; E_out[nbins], H_out[nbins], range optional. */
void n4_sharpen(const float *U, const uint8_t *m, int64_t V, int nbins, float fwhm, float wiener,
                float *S, float *E_out, float *H_out, float *range_out)
{
    float bmax = -FLT_MAX, bmin = FLT_MAX;
    for (int64_t v = 0; v < V; ++v)
        if (m[v]) {
            const float p = U[v];
            if (p > bmax) bmax = p;
            else if (p < bmin) bmin = p;
        }
    /* ITK really writes "if (pixel > max) max = pixel; else if (pixel < min) min = pixel;":
     * a voxel that raises the running maximum is never a candidate for the minimum.  The only
     * realistic consequence: when the FIRST masked voxel is the global minimum, binMinimum is
     * the second smallest value and that voxel drops out of the histogram.  Kept as is. */
    const float slope = (bmax - bmin) / (float)(nbins - 1);

    float *H = (float *)calloc((size_t)nbins, sizeof(float));
    for (int64_t v = 0; v < V; ++v)
        if (m[v]) {
            const float cidx = (U[v] - bmin) / slope;
            const unsigned idx = (unsigned)(int)floorf(cidx);
            const float offset = cidx - (float)idx;
            if (offset == 0.0f) {
                if (idx < (unsigned)nbins) H[idx] = (float)((double)H[idx] + 1.0);
            } else if (idx < (unsigned)nbins - 1) {
                H[idx] = (float)((double)H[idx] + (1.0 - (double)offset));
                H[idx + 1] += offset;
            }
        }

    const float exponent = ceilf((float)(log((double)(float)nbins) / log(2.0))) + 1.0f;
    const int N = (int)(unsigned)(pow(2.0, (double)exponent) + 0.5);
    const unsigned hoff = (unsigned)(0.5 * (N - nbins));

    double *buf = (double *)calloc((size_t)N * 12, sizeof(double));
    double *Vr = buf, *Vi = buf + N, *Fr = buf + 2 * N, *Fi = buf + 3 * N;
    double *Ur = buf + 4 * N, *Ui = buf + 5 * N, *Nr = buf + 6 * N, *Ni = buf + 7 * N;
    double *Dr = buf + 8 * N, *Di = buf + 9 * N, *G = buf + 10 * N;

    for (int n = 0; n < nbins; ++n) Vr[n + hoff] = H[n];
    fft_pow2(Vr, Vi, N, -1);

    const float scaledFWHM = fwhm / slope;
    const float expFactor = (float)(4.0 * log(2.0) / (double)(scaledFWHM * scaledFWHM));
    const float scaleFactor = (float)(2.0 * sqrt(log(2.0) / M_PI) / (double)scaledFWHM);
    Fr[0] = scaleFactor;
    const int half = (int)(0.5 * N);
    for (int n = 1; n <= half; ++n) {
        const float fn = (float)n;
        Fr[n] = Fr[N - n] = (double)(scaleFactor * expf(-(fn * fn) * expFactor));
    }
    if (N % 2 == 0) {
        const float fN = (float)N;
        Fr[half] = (double)scaleFactor * exp(0.25 * (double)(-(fN * fN)) * (double)expFactor);
    }
    fft_pow2(Fr, Fi, N, -1);

    const double noise = (double)wiener;
    for (int n = 0; n < N; ++n) {            /* Gf = conj(Ff) / (conj(Ff)*Ff + noise); keep Re */
        const double cr = Fr[n], ci = -Fi[n];
        const double dr = cr * Fr[n] - ci * Fi[n] + noise, di = cr * Fi[n] + ci * Fr[n];
        G[n] = (cr * dr + ci * di) / (dr * dr + di * di);
    }
    for (int n = 0; n < N; ++n) { Ur[n] = Vr[n] * G[n]; Ui[n] = Vi[n] * G[n]; }
    fft_pow2(Ur, Ui, N, +1);
    for (int n = 0; n < N; ++n) { Ur[n] = Ur[n] > 0.0 ? Ur[n] : 0.0; Ui[n] = 0.0; }

    for (int n = 0; n < N; ++n) {
        const float x = bmin + ((float)n - (float)hoff) * slope;
        Nr[n] = (double)x * Ur[n]; Ni[n] = 0.0;
        Dr[n] = Ur[n]; Di[n] = 0.0;
    }
    fft_pow2(Nr, Ni, N, -1);
    fft_pow2(Dr, Di, N, -1);
    for (int n = 0; n < N; ++n) {
        double a = Nr[n] * Fr[n] - Ni[n] * Fi[n], b = Nr[n] * Fi[n] + Ni[n] * Fr[n];
        Nr[n] = a; Ni[n] = b;
        a = Dr[n] * Fr[n] - Di[n] * Fi[n]; b = Dr[n] * Fi[n] + Di[n] * Fr[n];
        Dr[n] = a; Di[n] = b;
    }
    fft_pow2(Nr, Ni, N, +1);
    fft_pow2(Dr, Di, N, +1);

    float *E = (float *)malloc(sizeof(float) * (size_t)nbins);
    for (int n = 0; n < nbins; ++n) {
        const double d = Dr[n + hoff];
        E[n] = (d != 0.0) ? (float)(Nr[n + hoff] / d) : 0.0f;
    }

    if (S)
        for (int64_t v = 0; v < V; ++v)
            if (m[v]) {
                const float cidx = (U[v] - bmin) / slope;
                const unsigned idx = (unsigned)(int)floorf(cidx);
                if (idx < (unsigned)nbins - 1)
                    S[v] = E[idx] + (E[idx + 1] - E[idx]) * (cidx - (float)idx);
                else
                    S[v] = E[nbins - 1];
            }
    if (E_out) memcpy(E_out, E, sizeof(float) * (size_t)nbins);
    if (H_out) memcpy(H_out, H, sizeof(float) * (size_t)nbins);
    if (range_out) { range_out[0] = bmin; range_out[1] = bmax; }
    free(E); free(buf); free(H);
}

/* Per-axis parametric tables shared by fit and evaluation. */
typedef struct {
    int base[3][4096];       /* floor(p) */
    double B[3][4096][4];    /* kernel values for the 4 supporting control points */
} axis_tab;

/* BSplineScatteredDataPointSetToImageFilter::ThreadedGenerateDataForFitting, per axis part:
 * p = (point - origin) * r with r = (ncp-3)/((n-1)*spacing); clamp at the upper end by
 * epsilon = r*spacing*1e-3; u = p - floor(p) - k + 1. */
static void fit_axis_table(int n, int ncp, int *base, double (*B)[4])
{
    const int spans = ncp - 3;
    const float r = (n > 1) ? (float)((double)(float)spans / ((double)(float)(n - 1) * 1.0)) : 0.0f;
    const float eps = (float)((double)r * 1.0 * (double)BSPLINE_EPS);
    for (int i = 0; i < n; ++i) {
        float p = (float)(((double)(float)i - 0.0) * (double)r);
        if (fabsf(p - (float)spans) <= eps) p = (float)spans - eps;
        if (p < 0.0f && fabsf(p) <= eps) p = 0.0f;
        const unsigned ip = (unsigned)p;
        base[i] = (int)ip;
        for (int k = 0; k < 4; ++k) {
            const double u = (double)(float)(p - (float)ip - (float)k) + 0.5 * (double)(float)(3 - 1);
            B[i][k] = bspline3(u);
        }
    }
}

/* BSplineControlPointImageFilter::ThreadedGenerateData, per axis part:
 * U = spans * idx / (n-1) in float, same epsilon clamp; v = U - (floor(U)+k) + 1 (float). */
static void eval_axis_table(int n, int ncp, int *base, double (*B)[4])
{
    const int spans = ncp - 3;
    const float r = (n > 1) ? (float)((double)(float)spans / ((double)(float)(n - 1) * 1.0)) : 0.0f;
    const float eps = (float)((double)r * 1.0 * (double)BSPLINE_EPS);
    for (int i = 0; i < n; ++i) {
        float U = (n > 1) ? (float)spans * (float)i / (float)(n - 1) : 0.0f;
        if (fabsf(U - (float)spans) <= eps) U = (float)spans - eps;
        if (U < 0.0f && fabsf(U) <= eps) U = 0.0f;
        const unsigned ip = (unsigned)U;
        base[i] = (int)ip;
        for (int k = 0; k < 4; ++k) {
            const float v = (float)((double)(U - (float)(ip + (unsigned)k)) + 0.5 * (double)(float)(3 - 1));
            B[i][k] = (double)(float)bspline3((double)v);
        }
    }
}

/* One-level MBA fit of the masked residual (Lee-Wolberg-Shin), ITK accumulation order.
 * ncp[] in numpy axis order (row, col, slice); phi is [ncp0][ncp1][ncp2]. */
void n4_fit(const float *resid, const uint8_t *m, int n0, int n1, int n2, const int *ncp, float *phi)
{
    axis_tab *T = (axis_tab *)malloc(sizeof(axis_tab));
    const int n[3] = {n0, n1, n2};
    for (int a = 0; a < 3; ++a) fit_axis_table(n[a], ncp[a], T->base[a], T->B[a]);
    const int L = ncp[0] * ncp[1] * ncp[2];
    float *delta = (float *)calloc((size_t)L, sizeof(float));
    float *omega = (float *)calloc((size_t)L, sizeof(float));
    float w[4][4][4];
    for (int i = 0; i < n0; ++i)
        for (int j = 0; j < n1; ++j)
            for (int k = 0; k < n2; ++k) {
                const int64_t v = ((int64_t)i * n1 + j) * n2 + k;
                if (!m[v]) continue;
                const double *Br = T->B[0][i], *Bc = T->B[1][j], *Bs = T->B[2][k];
                float w2 = 0.0f;
                for (int a = 0; a < 4; ++a)          /* ITK z (row)   : slowest */
                    for (int b = 0; b < 4; ++b)      /* ITK y (col)            */
                        for (int c = 0; c < 4; ++c) {/* ITK x (slice) : fastest */
                            float B = 1.0f;
                            B = (float)((double)B * Bs[c]);   /* dimension 0 = x */
                            B = (float)((double)B * Bc[b]);   /* dimension 1 = y */
                            B = (float)((double)B * Br[a]);   /* dimension 2 = z */
                            w[a][b][c] = B;
                            w2 += B * B;
                        }
                const float z = resid[v];
                for (int a = 0; a < 4; ++a)
                    for (int b = 0; b < 4; ++b)
                        for (int c = 0; c < 4; ++c) {
                            const int idx = ((T->base[0][i] + a) * ncp[1] + (T->base[1][j] + b)) * ncp[2] +
                                            (T->base[2][k] + c);
                            const float t = w[a][b][c];
                            omega[idx] += 1.0f * t * t;
                            float d = z;
                            d *= (t * t * t * 1.0f / w2);
                            delta[idx] += d;
                        }
            }
    for (int q = 0; q < L; ++q)  /* Math::NotAlmostEquals(omega, 0): |omega| > 0.1*FLT_EPSILON */
        phi[q] = (fabsf(omega[q]) > 0.1f * FLT_EPSILON) ? delta[q] / omega[q] : 0.0f;
    free(delta); free(omega); free(T);
}

/* BSplineControlPointImageFilter: evaluate the lattice at every voxel, collapsing the
 * row axis (ITK z) first, then columns, then slices, float accumulation. */
void n4_eval(const float *lat, const int *ncp, int n0, int n1, int n2, float *field)
{
    axis_tab *T = (axis_tab *)malloc(sizeof(axis_tab));
    const int n[3] = {n0, n1, n2};
    for (int a = 0; a < 3; ++a) eval_axis_table(n[a], ncp[a], T->base[a], T->B[a]);
    float *plane = (float *)malloc(sizeof(float) * (size_t)ncp[1] * ncp[2]);
    float *line = (float *)malloc(sizeof(float) * (size_t)ncp[2]);
    for (int i = 0; i < n0; ++i) {
        for (int b = 0; b < ncp[1]; ++b)
            for (int c = 0; c < ncp[2]; ++c) {
                float d = 0.0f;
                for (int k = 0; k < 4; ++k)
                    d += lat[((T->base[0][i] + k) * ncp[1] + b) * ncp[2] + c] * (float)T->B[0][i][k];
                plane[b * ncp[2] + c] = d;
            }
        for (int j = 0; j < n1; ++j) {
            for (int c = 0; c < ncp[2]; ++c) {
                float d = 0.0f;
                for (int k = 0; k < 4; ++k)
                    d += plane[(T->base[1][j] + k) * ncp[2] + c] * (float)T->B[1][j][k];
                line[c] = d;
            }
            for (int s = 0; s < n2; ++s) {
                float d = 0.0f;
                for (int k = 0; k < 4; ++k) d += line[T->base[2][s] + k] * (float)T->B[2][s][k];
                field[((int64_t)i * n1 + j) * n2 + s] = d;
            }
        }
    }
    free(plane); free(line); free(T);
}

/* BSplineControlPointImageFilter::RefineControlPointLattice, one doubling of every axis
 * (n -> 2n-3).  Coefficients of the cubic two-scale relation ([1/2 1/2 0 0; 1/8 3/4 1/8 0],
 * which ITK obtains numerically from an SVD solve in float), 64-term float sum with the
 * ITK x (slice) axis fastest. */
void n4_refine(const float *lat, const int *ncp, float *out, int *ncp_out)
{
    static const float C[2][4] = {{0.5f, 0.5f, 0.0f, 0.0f}, {0.125f, 0.75f, 0.125f, 0.0f}};
    for (int a = 0; a < 3; ++a) ncp_out[a] = 2 * ncp[a] - 3;
    for (int i = 0; i < ncp_out[0]; ++i)
        for (int j = 0; j < ncp_out[1]; ++j)
            for (int k = 0; k < ncp_out[2]; ++k) {
                const int bi = i / 2, bj = j / 2, bk = k / 2, oi = i & 1, oj = j & 1, ok = k & 1;
                float sum = 0.0f;
                for (int a = 0; a < 4; ++a)
                    for (int b = 0; b < 4; ++b)
                        for (int c = 0; c < 4; ++c) {
                            if (bi + a >= ncp[0] || bj + b >= ncp[1] || bk + c >= ncp[2]) continue;
                            float coeff = 1.0f;
                            coeff *= C[ok][c];
                            coeff *= C[oj][b];
                            coeff *= C[oi][a];
                            float val = lat[((bi + a) * ncp[1] + (bj + b)) * ncp[2] + (bk + c)];
                            val *= coeff;
                            sum += val;
                        }
                out[(i * ncp_out[1] + j) * ncp_out[2] + k] = sum;
            }
}

/* N4BiasFieldCorrectionImageFilter::CalculateConvergenceMeasurement: coefficient of variation
 * of exp(field1 - field2) over the mask, float Welford in visiting order. */
float n4_convergence(const float *f1, const float *f2, const uint8_t *m, int64_t V)
{
    float mu = 0.0f, sigma = 0.0f, N = 0.0f;
    for (int64_t v = 0; v < V; ++v)
        if (m[v]) {
            const float pixel = expf(f1[v] - f2[v]);
            N = (float)((double)N + 1.0);
            if (N > 1.0f) {
                const float d = pixel - mu;
                sigma = (float)((double)sigma + (double)(d * d) * ((double)N - 1.0) / (double)N);
            }
            mu = (float)((double)mu * (1.0 - 1.0 / (double)N) + (double)(pixel / N));
        }
    sigma = (float)sqrt((double)sigma / ((double)N - 1.0));
    return sigma / mu;
}

/* Full filter.  in: float32 [n0][n1][n2]; mask: uint8 (voxels == mask_label take part).
 * max_iter[nlevels]; ncp0[3] initial control points per numpy axis.
 * Outputs: out (corrected, every voxel), logbias (every voxel), iters[nlevels],
 * conv_trace (optional, nlevels*max(max_iter) floats, -1 filled),
 * lattice_out (optional, capacity >= prod(final ncp)), ncp_final[3] (optional).
 * returns 0 ok, 2 bad arguments. */
int n4_oracle(const float *in, const uint8_t *mask, int n0, int n1, int n2, int mask_label,
              int nlevels, const int *max_iter, float conv_thr, float fwhm, float wiener, int nbins,
              const int *ncp0, float *out, float *logbias, int *iters, float *conv_trace,
              int conv_stride, float *lattice_out, int *ncp_final)
{
    const int64_t V = (int64_t)n0 * n1 * n2;
    if (n0 > 4096 || n1 > 4096 || n2 > 4096 || nlevels < 1) return 2;
    uint8_t *m = (uint8_t *)malloc((size_t)V);
    float *logI = (float *)malloc(sizeof(float) * (size_t)V);
    float *U = (float *)malloc(sizeof(float) * (size_t)V);
    float *S = (float *)calloc((size_t)V, sizeof(float));
    float *resid = (float *)malloc(sizeof(float) * (size_t)V);
    float *bias = (float *)calloc((size_t)V, sizeof(float));
    float *nbias = (float *)malloc(sizeof(float) * (size_t)V);
    int ncp[3] = {ncp0[0], ncp0[1], ncp0[2]};
    float *lat = NULL, *phi = NULL;

    for (int64_t v = 0; v < V; ++v) {
        m[v] = (mask[v] == (uint8_t)mask_label);
        logI[v] = in[v];
        if (m[v] && in[v] > 0.0f) logI[v] = logf(in[v]);
        U[v] = logI[v];
    }
    if (conv_trace)
        for (int q = 0; q < nlevels * conv_stride; ++q) conv_trace[q] = -1.0f;

    for (int level = 0; level < nlevels; ++level) {
        float conv = FLT_MAX;
        int it = 0;
        while (it++ < max_iter[level] && conv > conv_thr) {
            n4_sharpen(U, m, V, nbins, fwhm, wiener, S, NULL, NULL, NULL);
            for (int64_t v = 0; v < V; ++v) resid[v] = U[v] - S[v];
            const int L = ncp[0] * ncp[1] * ncp[2];
            phi = (float *)malloc(sizeof(float) * (size_t)L);
            n4_fit(resid, m, n0, n1, n2, ncp, phi);
            if (!lat) {
                lat = phi;
            } else {
                for (int q = 0; q < L; ++q) lat[q] = lat[q] + phi[q];
                free(phi);
            }
            n4_eval(lat, ncp, n0, n1, n2, nbias);
            conv = n4_convergence(bias, nbias, m, V);
            if (conv_trace && it - 1 < conv_stride) conv_trace[level * conv_stride + it - 1] = conv;
            memcpy(bias, nbias, sizeof(float) * (size_t)V);
            for (int64_t v = 0; v < V; ++v) U[v] = logI[v] - bias[v];
        }
        iters[level] = it - 1;
        if (level != nlevels - 1) {
            int nn[3];
            float *ref = (float *)malloc(sizeof(float) * (size_t)(2 * ncp[0] - 3) * (2 * ncp[1] - 3) * (2 * ncp[2] - 3));
            n4_refine(lat, ncp, ref, nn);
            free(lat);
            lat = ref;
            ncp[0] = nn[0]; ncp[1] = nn[1]; ncp[2] = nn[2];
        }
    }
    for (int64_t v = 0; v < V; ++v) out[v] = in[v] / expf(bias[v]);
    if (logbias) memcpy(logbias, bias, sizeof(float) * (size_t)V);
    if (lattice_out) memcpy(lattice_out, lat, sizeof(float) * (size_t)ncp[0] * ncp[1] * ncp[2]);
    if (ncp_final) { ncp_final[0] = ncp[0]; ncp_final[1] = ncp[1]; ncp_final[2] = ncp[2]; }
    free(m); free(logI); free(U); free(S); free(resid); free(bias); free(nbias); free(lat);
    return 0;
}
