// N4 bias-field correction, one persistent CTA per subject.
//
// Replaces Vent_Analysis.N4_bias_correction (Vent_Analysis.py:316-334), i.e.
// sitk.N4BiasFieldCorrectionImageFilter() with all defaults (ITK 5.3
// itkN4BiasFieldCorrectionImageFilter.hxx; restated on the CPU in oracle/n4_oracle.c).
//
// Design:
//   * A CTA pulls subjects from a global work counter and runs the WHOLE filter for one
//     subject -- 4 levels x <=50 iterations x {histogram, Wiener sharpening, B-spline fit,
//     update/convergence} -- with block barriers only; no host round trip, subjects converge
//     independently.
//   * Volume layout is the reference's [row][col][slice] C order.  A thread owns one
//     (col, slice) column and walks it along the row axis, so a warp touches 32 consecutive
//     floats per row step (coalesced) and the row-axis B-spline weights are warp uniform.
//     A per-column row bitmask (built once per subject) lets warps skip rows and whole
//     column tiles that hold no mask voxel: only masked sectors of U are ever read.
//   * Parzen histogram: every thread owns a private 200-bin column in shared memory
//     ([bin][thread], conflict free, plain LDS/FADD/STS instead of atomics), reduced by a
//     fixed warp tree => deterministic.
//   * Sharpening: 4 packed 512-point complex FFTs in double in shared memory.
//   * B-spline fit: the Lee-Wolberg-Shin scatter is separable (t = tr*tc*ts,
//     sum t^2 = Sr*Sc*Ss), so delta = sum w^3 z / sum w^2 is three small contractions:
//     rows (registers, sliding 4-tap window) -> cols (shared Q[a][b][s]) -> slices.
//     omega depends only on the mask and the lattice size: once per level.
//   * U = logI - bias is kept in the output buffer during the iterations; the corrected
//     image overwrites it in the final dense pass (the only HBM-streaming part).
#pragma once
#include <float.h>
#include "vb_common.cuh"

#define VB_N4_THREADS 256
#define VB_BSPLINE_EPS 1e-3f

struct N4Args {
    const float* image;
    const uint8_t* mask;
    float* out;
    float* logbias;          // may be null
    const VbSubject* subj;
    int n;
    vb_n4_params p;
    vb_subject_stats* stats; // may be null
    // per-CTA workspace (slot = blockIdx.x)
    float* ws_logI;          // [grid][Vmax]
    uint32_t* ws_colmask;    // [grid][cmw_max]
    float* ws_axis;          // [grid][axis_stride]  (4-byte words)
    int* work_counter;
    float* debug;            // may be null: [n][debug_stride]
    long long Vmax, cmw_max;
    int axis_stride, debug_stride;
    int hist_threads;        // HT: threads that own a private histogram column
    int Lmax;                // max lattice entries (after the last refinement)
    int fftN;                // padded histogram size
    int union_bytes;         // size of the shared-memory union region
};

// debug record layout (floats): [0]=bmin [1]=bmax [2..4]=ncp [5]=n_trace, then
// hist[nbins], emap[nbins], lattice[Lmax], trace[...]
#define VB_DBG_HDR 8

VB_DEV double vb_bspline3(double u) {
    const double a = fabs(u);
    if (a < 1.0) {
        const double s = __dmul_rn(a, a);
        // (4 - 6 s + 3 s a) / 6, evaluated left to right like the C oracle
        double t = __dsub_rn(4.0, __dmul_rn(6.0, s));
        t = __dadd_rn(t, __dmul_rn(__dmul_rn(3.0, s), a));
        return __ddiv_rn(t, 6.0);
    } else if (a < 2.0) {
        const double s = __dmul_rn(a, a);
        double t = __dsub_rn(8.0, __dmul_rn(12.0, a));
        t = __dadd_rn(t, __dmul_rn(6.0, s));
        t = __dsub_rn(t, __dmul_rn(s, a));
        return __ddiv_rn(t, 6.0);
    }
    return 0.0;
}

// Per-axis tables, 16 words per sample: fit3[4] | fit2[4] | evw[4] | bfit | bev | pad[2]
struct N4Axis {
    const float4* fit3;   // B^3 / sum(B^2)   (delta numerator weights)
    const float4* fit2;   // B^2              (omega weights)
    const float4* evw;    // B                (field evaluation)
    const int* bfit;      // first supporting control point, fit parameterisation
    const int* bev;       // first supporting control point, evaluation parameterisation
};
VB_DEV N4Axis n4_axis(float* base, int n) {
    N4Axis a;
    a.fit3 = (const float4*)base;
    a.fit2 = (const float4*)(base + 4 * (size_t)n);
    a.evw = (const float4*)(base + 8 * (size_t)n);
    a.bfit = (const int*)(base + 12 * (size_t)n);
    a.bev = (const int*)(base + 13 * (size_t)n);
    return a;
}

// BSplineScatteredDataPointSetToImageFilter / BSplineControlPointImageFilter parametrisation
// of one axis (oracle: fit_axis_table / eval_axis_table in n4_oracle.c).
VB_DEV void n4_build_axis(float* base, int n, int ncp) {
    float4* fit3 = (float4*)base;
    float4* fit2 = (float4*)(base + 4 * (size_t)n);
    float4* evw = (float4*)(base + 8 * (size_t)n);
    int* bfit = (int*)(base + 12 * (size_t)n);
    int* bev = (int*)(base + 13 * (size_t)n);
    const int spans = ncp - 3;
    const float fs = (float)spans;
    const float r = (n > 1) ? (float)__ddiv_rn((double)fs, (double)(float)(n - 1)) : 0.0f;
    const float eps = (float)__dmul_rn((double)r, (double)VB_BSPLINE_EPS);
    for (int i = (int)threadIdx.x; i < n; i += (int)blockDim.x) {
        float p = (float)__dmul_rn((double)(float)i, (double)r);
        if (fabsf(p - fs) <= eps) p = fs - eps;
        if (p < 0.0f && fabsf(p) <= eps) p = 0.0f;
        const unsigned ip = (unsigned)p;
        float b[4], S = 0.0f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double u = __dadd_rn((double)__fsub_rn(__fsub_rn(p, (float)ip), (float)k), 1.0);
            b[k] = (float)vb_bspline3(u);
            S = __fadd_rn(S, __fmul_rn(b[k], b[k]));
        }
        fit3[i] = make_float4(b[0] * b[0] * b[0] / S, b[1] * b[1] * b[1] / S, b[2] * b[2] * b[2] / S,
                              b[3] * b[3] * b[3] / S);
        fit2[i] = make_float4(b[0] * b[0], b[1] * b[1], b[2] * b[2], b[3] * b[3]);
        bfit[i] = (int)ip;

        float U = (n > 1) ? __fdiv_rn(__fmul_rn(fs, (float)i), (float)(n - 1)) : 0.0f;
        if (fabsf(U - fs) <= eps) U = fs - eps;
        if (U < 0.0f && fabsf(U) <= eps) U = 0.0f;
        const unsigned iu = (unsigned)U;
        float e[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float v = (float)__dadd_rn((double)__fsub_rn(U, (float)(iu + (unsigned)k)), 1.0);
            e[k] = (float)vb_bspline3((double)v);
        }
        evw[i] = make_float4(e[0], e[1], e[2], e[3]);
        bev[i] = (int)iu;
    }
}

struct N4Ctx {
    // subject
    int n0, n1, n2, ncols, W;
    long long V;
    const float* img;
    const uint8_t* msk;
    float* U;        // aliases the output buffer
    float* logI;
    uint32_t* colmask;
    float* axis_base;
    // shared memory
    float* un_f;     // union region
    double* un_d;
    float* L;        // lattice [ncp0][ncp1][ncp2]
    float* Om;       // omega lattice of the current level
    float* E;        // E(u|v) table
    float* H;        // histogram
    double* red;     // 64 doubles reduction scratch
    int ncp[3];
    int nbins, HT, fftN;
};

VB_DEV N4Axis n4_ax(const N4Ctx& c, int a) {
    const int n[3] = {c.n0, c.n1, c.n2};
    size_t off = 0;
    for (int i = 0; i < a; ++i) off += 16 * (size_t)n[i];
    return n4_axis(c.axis_base + off, n[a]);
}

// ---------------------------------------------------------------------------------------------
// setup: per-column row bitmask, logI, U, range and count of the mask
VB_DEV void n4_setup(N4Ctx& c, int label, float& bmin, float& bmax, int& M) {
    const int nt = (int)blockDim.x, tid = (int)threadIdx.x;
    float tmin = FLT_MAX, tmax = -FLT_MAX;
    int cnt = 0;
    for (int q = tid; q < c.ncols; q += nt) {
        for (int wi = 0; wi < c.W; ++wi) {
            uint32_t bits = 0;
            const int rend = min(32, c.n0 - wi * 32);
            for (int b = 0; b < rend; ++b) {
                const size_t v = (size_t)(wi * 32 + b) * c.ncols + q;
                if ((int)vb_ld_stream(c.msk + v) == label) {
                    bits |= 1u << b;
                    const float x = vb_ld_stream(c.img + v);
                    const float l = (x > 0.0f) ? logf(x) : x;
                    c.logI[v] = l;
                    c.U[v] = l;
                    tmin = fminf(tmin, l);
                    tmax = fmaxf(tmax, l);
                    ++cnt;
                }
            }
            c.colmask[(size_t)wi * c.ncols + q] = bits;
        }
    }
    bmin = vb_block_reduce(tmin, VbMinF(), FLT_MAX, (float*)c.red);
    bmax = vb_block_reduce(tmax, VbMaxF(), -FLT_MAX, (float*)c.red);
    M = vb_block_reduce(cnt, VbAddI(), 0, (int*)c.red);
}

// ---------------------------------------------------------------------------------------------
// P1: triangular Parzen histogram of U over the mask (SharpenImage, first half)
VB_DEV void n4_histogram(N4Ctx& c, float bmin, float bmax) {
    const int nt = (int)blockDim.x, tid = (int)threadIdx.x, HT = c.HT;
    float* priv = c.un_f;
    for (int i = tid; i < c.nbins * HT; i += nt) priv[i] = 0.0f;
    __syncthreads();
    const float slope = (bmax - bmin) / (float)(c.nbins - 1);
    if (tid < HT) {
        for (int q0 = 0; q0 < c.ncols; q0 += HT) {
            const int q = q0 + tid;
            const bool valid = q < c.ncols;
            for (int wi = 0; wi < c.W; ++wi) {
                const uint32_t mw = valid ? c.colmask[(size_t)wi * c.ncols + q] : 0u;
                const uint32_t wm = __reduce_or_sync(VB_FULL, mw);
                if (!wm) continue;
#pragma unroll 1
                for (int b0 = 0; b0 < 32; b0 += 8) {
                    if (((wm >> b0) & 0xFFu) == 0) continue;
                    float u[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        u[j] = ((mw >> (b0 + j)) & 1u) ? c.U[(size_t)(wi * 32 + b0 + j) * c.ncols + q] : 0.0f;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if ((mw >> (b0 + j)) & 1u) {
                            const float cidx = __fdiv_rn(u[j] - bmin, slope);
                            const float fl = floorf(cidx);
                            const int idx = (int)fl;
                            const float off = cidx - fl;
                            if (off == 0.0f) {
                                if (idx >= 0 && idx < c.nbins) priv[idx * HT + tid] += 1.0f;
                            } else if (idx >= 0 && idx < c.nbins - 1) {
                                priv[idx * HT + tid] += 1.0f - off;
                                priv[(idx + 1) * HT + tid] += off;
                            }
                        }
                }
            }
        }
    }
    __syncthreads();
    // fixed-order reduction: one warp per bin
    const int nw = nt >> 5, w = tid >> 5, lane = tid & 31;
    for (int bin = w; bin < c.nbins; bin += nw) {
        float s = 0.0f;
        for (int t = lane; t < HT; t += 32) s += priv[bin * HT + t];
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) s += __shfl_xor_sync(VB_FULL, s, m);
        if (lane == 0) c.H[bin] = s;
    }
    __syncthreads();
}

// in-place radix-2 FFT on shared (re, im); tw = exp(-2 pi i k / N); sign<0 forward, >0 backward
VB_DEV void n4_fft(double* re, double* im, int N, int logN, int sign, const double* twr, const double* twi) {
    const int nt = (int)blockDim.x, tid = (int)threadIdx.x;
    for (int i = tid; i < N; i += nt) {
        const int j = (int)(__brev((unsigned)i) >> (32 - logN));
        if (i < j) {
            double t = re[i]; re[i] = re[j]; re[j] = t;
            t = im[i]; im[i] = im[j]; im[j] = t;
        }
    }
    __syncthreads();
    for (int len = 2; len <= N; len <<= 1) {
        const int half = len >> 1, step = N / len;
        for (int b = tid; b < (N >> 1); b += nt) {
            const int grp = b / half, k = b - grp * half;
            const int i0 = grp * len + k, i1 = i0 + half;
            const double wr = twr[k * step], wi = (sign < 0) ? twi[k * step] : -twi[k * step];
            const double xr = re[i1] * wr - im[i1] * wi, xi = re[i1] * wi + im[i1] * wr;
            re[i1] = re[i0] - xr; im[i1] = im[i0] - xi;
            re[i0] += xr; im[i0] += xi;
        }
        __syncthreads();
    }
}

// SharpenImage, second half: Wiener-deconvolve the histogram with the Gaussian bias model and
// build E(u|v).  Two real sequences share one complex FFT (4 FFTs instead of ITK's 7).
VB_DEV void n4_sharpen(N4Ctx& c, float bmin, float bmax, float fwhm, float wiener) {
    const int nt = (int)blockDim.x, tid = (int)threadIdx.x;
    const int N = c.fftN;
    int logN = 0;
    while ((1 << logN) < N) ++logN;
    const int hoff = (int)(0.5 * (double)(N - c.nbins));
    double* Ar = c.un_d;          double* Ai = Ar + N;
    double* Br = Ai + N;          double* Bi = Br + N;
    double* Fr = Bi + N;          double* Fi = Fr + N;     // Ff
    double* twr = Fi + N;         double* twi = twr + (N >> 1);
    const float slope = (bmax - bmin) / (float)(c.nbins - 1);
    const float sF = fwhm / slope;
    const float expFactor = (float)(4.0 * 0.69314718055994530942 / (double)(sF * sF));
    const float scaleFactor = (float)(2.0 * sqrt(0.69314718055994530942 / 3.14159265358979323846) / (double)sF);
    const int half = N >> 1;
    for (int n = tid; n < N; n += nt) {
        const int h = n - hoff;
        Ar[n] = (h >= 0 && h < c.nbins) ? (double)c.H[h] : 0.0;
        float f;
        if (n == 0) {
            f = scaleFactor;
            Ai[n] = (double)f;
        } else if (n == half) {
            const float fN = (float)N;
            Ai[n] = (double)scaleFactor * exp(0.25 * (double)(-(fN * fN)) * (double)expFactor);
        } else {
            const float fn = (float)(n < half ? n : N - n);
            f = scaleFactor * expf(-(fn * fn) * expFactor);
            Ai[n] = (double)f;
        }
        if (n < half) {
            double s, co;
            vb_sincospi(-2.0 * (double)n / (double)N, &s, &co);
            twr[n] = co; twi[n] = s;
        }
    }
    __syncthreads();
    n4_fft(Ar, Ai, N, logN, -1, twr, twi);
    const double noise = (double)wiener;
    for (int k = tid; k < N; k += nt) {
        const int kk = (N - k) & (N - 1);
        const double ar = Ar[k], ai = Ai[k], br = Ar[kk], bi = Ai[kk];
        const double vr = 0.5 * (ar + br), vi = 0.5 * (ai - bi);       // Vf
        const double fr = 0.5 * (ai + bi), fi = -0.5 * (ar - br);      // Ff
        Fr[k] = fr; Fi[k] = fi;
        const double g = fr / (fr * fr + fi * fi + noise);             // Re(conj(Ff)/(|Ff|^2+noise))
        Br[k] = vr * g; Bi[k] = vi * g;
    }
    __syncthreads();
    n4_fft(Br, Bi, N, logN, +1, twr, twi);
    for (int n = tid; n < N; n += nt) {
        const double u = Br[n] > 0.0 ? Br[n] : 0.0;
        const float x = bmin + ((float)n - (float)hoff) * slope;
        Ar[n] = (double)x * u;
        Ai[n] = u;
    }
    __syncthreads();
    n4_fft(Ar, Ai, N, logN, -1, twr, twi);
    for (int k = tid; k < N; k += nt) {
        const int kk = (N - k) & (N - 1);
        const double ar = Ar[k], ai = Ai[k], br = Ar[kk], bi = Ai[kk];
        const double nr = 0.5 * (ar + br), ni = 0.5 * (ai - bi);       // FFT(x*U)
        const double dr = 0.5 * (ai + bi), di = -0.5 * (ar - br);      // FFT(U)
        const double fr = Fr[k], fi = Fi[k];
        const double nfr = nr * fr - ni * fi, nfi = nr * fi + ni * fr;
        const double dfr = dr * fr - di * fi, dfi = dr * fi + di * fr;
        Br[k] = nfr - dfi; Bi[k] = nfi + dfr;                          // NF + i*DF
    }
    __syncthreads();
    n4_fft(Br, Bi, N, logN, +1, twr, twi);
    for (int n = tid; n < c.nbins; n += nt) {
        const double den = Bi[n + hoff];
        c.E[n] = (den != 0.0) ? (float)(Br[n + hoff] / den) : 0.0f;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Separable B-spline contraction over the mask.
//   OMEGA = true : value 1, weights B^2            -> c.Om
//   OMEGA = false: value U - E(U), weights B^3/S   -> phi = delta/omega, c.L += phi
template <bool OMEGA>
VB_DEV void n4_contract(N4Ctx& c, float bmin, float bmax) {
    const int nt = (int)blockDim.x, tid = (int)threadIdx.x;
    const int ncp0 = c.ncp[0], ncp1 = c.ncp[1], ncp2 = c.ncp[2];
    const int n2 = c.n2, ncols = c.ncols;
    const N4Axis ar = n4_ax(c, 0), ac = n4_ax(c, 1), as = n4_ax(c, 2);
    float* Q = c.un_f;                                   // [ncp0][ncp1][n2]
    const int Qn = ncp0 * ncp1 * n2;
    float* Pt = Q + ((Qn + 3) & ~3);                     // [ncp0][nt]
    for (int i = tid; i < Qn; i += nt) Q[i] = 0.0f;
    const float slope = (bmax - bmin) / (float)(c.nbins - 1);
    const int nbm1 = c.nbins - 1;
    __syncthreads();

    for (int q0 = 0; q0 < ncols; q0 += nt) {
        const int q = q0 + tid;
        const bool valid = q < ncols;
        for (int a = 0; a < ncp0; ++a) Pt[a * nt + tid] = 0.0f;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int cur = -1;
        for (int wi = 0; wi < c.W; ++wi) {
            const uint32_t mw = valid ? c.colmask[(size_t)wi * ncols + q] : 0u;
            const uint32_t wm = __reduce_or_sync(VB_FULL, mw);
            if (!wm) continue;
#pragma unroll 1
            for (int b0 = 0; b0 < 32; b0 += 8) {
                if (((wm >> b0) & 0xFFu) == 0) continue;
                float u[8];
                if (!OMEGA) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        u[j] = ((mw >> (b0 + j)) & 1u) ? c.U[(size_t)(wi * 32 + b0 + j) * ncols + q] : 0.0f;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (!((wm >> (b0 + j)) & 1u)) continue;       // warp uniform
                    const int r = wi * 32 + b0 + j;
                    const int base = ar.bfit[r];
                    if (base != cur) {                            // warp uniform
                        if (cur >= 0) {
                            while (cur < base) {
                                Pt[cur * nt + tid] = a0;
                                a0 = a1; a1 = a2; a2 = a3; a3 = 0.f;
                                ++cur;
                            }
                        } else {
                            cur = base;
                        }
                    }
                    if ((mw >> (b0 + j)) & 1u) {
                        float val;
                        float4 w;
                        if (OMEGA) {
                            val = 1.0f;
                            w = ar.fit2[r];
                        } else {
                            const float cidx = __fdiv_rn(u[j] - bmin, slope);
                            const float fl = floorf(cidx);
                            const int idx = (int)fl;
                            float s;
                            if (idx >= 0 && idx < nbm1) {
                                const float e0 = c.E[idx], e1 = c.E[idx + 1];
                                s = e0 + (e1 - e0) * (cidx - fl);
                            } else {
                                s = c.E[nbm1];
                            }
                            val = u[j] - s;
                            w = ar.fit3[r];
                        }
                        a0 += w.x * val; a1 += w.y * val; a2 += w.z * val; a3 += w.w * val;
                    }
                }
            }
        }
        if (cur >= 0) {
            Pt[cur * nt + tid] = a0;
            Pt[(cur + 1) * nt + tid] = a1;
            Pt[(cur + 2) * nt + tid] = a2;
            Pt[(cur + 3) * nt + tid] = a3;
        }
        __syncthreads();
        // columns -> Q[a][b][s]: one thread per (a, s) of this tile
        const int tile_cols = min(nt, ncols - q0);
        const int ns_t = min(n2, tile_cols);
        const int s_first = q0 % n2;
        for (int pi = tid; pi < ncp0 * ns_t; pi += nt) {
            const int a = pi / ns_t;
            int s = s_first + (pi - a * ns_t);
            if (s >= n2) s -= n2;
            int d = s - s_first;
            if (d < 0) d += n2;
            for (int ql = d; ql < tile_cols; ql += n2) {
                const float val = Pt[a * nt + ql];
                if (val != 0.0f) {
                    const int cc = (q0 + ql) / n2;
                    const int bc = ac.bfit[cc];
                    const float4 w = OMEGA ? ac.fit2[cc] : ac.fit3[cc];
                    float* qp = Q + ((size_t)a * ncp1 + bc) * n2 + s;
                    qp[0] += w.x * val;
                    qp[n2] += w.y * val;
                    qp[2 * n2] += w.z * val;
                    qp[3 * n2] += w.w * val;
                }
            }
        }
        __syncthreads();
    }
    // slices -> lattice
    const int Ln = ncp0 * ncp1 * ncp2;
    for (int e = tid; e < Ln; e += nt) {
        const int g = e % ncp2, ab = e / ncp2;
        float sum = 0.0f;
        for (int s = 0; s < n2; ++s) {
            const int k = g - as.bfit[s];
            if (k >= 0 && k < 4) {
                const float4 w = OMEGA ? as.fit2[s] : as.fit3[s];
                const float wk = (k == 0) ? w.x : (k == 1) ? w.y : (k == 2) ? w.z : w.w;
                sum += Q[(size_t)ab * n2 + s] * wk;
            }
        }
        if (OMEGA) {
            c.Om[e] = sum;
        } else {
            const float om = c.Om[e];
            // itk::Math::NotAlmostEquals(omega, 0): |omega| > 0.1 * FLT_EPSILON
            const float phi = (fabsf(om) > 0.1f * FLT_EPSILON) ? sum / om : 0.0f;
            c.L[e] += phi;
        }
    }
    __syncthreads();
}

// collapsed column coefficients G[a] = sum_{b,g} wc[b] ws[g] L[a][bc+b][bs+g]  ->  Gs[a][tid]
VB_DEV void n4_column_coeffs(const N4Ctx& c, const N4Axis& ac, const N4Axis& as, int q, float* Gs) {
    const int nt = (int)blockDim.x, tid = (int)threadIdx.x;
    const int cc = q / c.n2, ss = q - cc * c.n2;
    const int bc = ac.bev[cc], bs = as.bev[ss];
    const float4 wc = ac.evw[cc], ws = as.evw[ss];
    const float wcv[4] = {wc.x, wc.y, wc.z, wc.w};
    for (int a = 0; a < c.ncp[0]; ++a) {
        float g = 0.0f;
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) {
            const float* lp = c.L + ((size_t)a * c.ncp[1] + bc + k1) * c.ncp[2] + bs;
            g += wcv[k1] * (ws.x * lp[0] + ws.y * lp[1] + ws.z * lp[2] + ws.w * lp[3]);
        }
        Gs[a * nt + tid] = g;
    }
}

#define VB_N4_WINDOW_ADVANCE(base)                                                   \
    if ((base) != cur) {                                                             \
        if (cur < 0 || (base) - cur >= 4 || (base) < cur) {                          \
            g0 = Gs[(base) * nt + tid]; g1 = Gs[((base) + 1) * nt + tid];            \
            g2 = Gs[((base) + 2) * nt + tid]; g3 = Gs[((base) + 3) * nt + tid];      \
            cur = (base);                                                            \
        } else {                                                                     \
            while (cur < (base)) {                                                   \
                g0 = g1; g1 = g2; g2 = g3; g3 = Gs[(cur + 4) * nt + tid]; ++cur;     \
            }                                                                        \
        }                                                                            \
    }

// ---------------------------------------------------------------------------------------------
// P3: evaluate the new bias on the mask, U = logI - bias, convergence statistics and new range
VB_DEV float n4_update(N4Ctx& c, int M, float& bmin, float& bmax) {
    const int nt = (int)blockDim.x, tid = (int)threadIdx.x;
    const int ncols = c.ncols;
    const N4Axis ar = n4_ax(c, 0), ac = n4_ax(c, 1), as = n4_ax(c, 2);
    float* Gs = c.un_f;
    double s1 = 0.0, s2 = 0.0;
    float tmin = FLT_MAX, tmax = -FLT_MAX;
    for (int q0 = 0; q0 < ncols; q0 += nt) {
        const int q = q0 + tid;
        const bool valid = q < ncols;
        uint32_t any = 0;
        for (int wi = 0; wi < c.W; ++wi) any |= valid ? c.colmask[(size_t)wi * ncols + q] : 0u;
        if (any) n4_column_coeffs(c, ac, as, q, Gs);
        float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
        int cur = -1;
        for (int wi = 0; wi < c.W; ++wi) {
            const uint32_t mw = valid ? c.colmask[(size_t)wi * ncols + q] : 0u;
            const uint32_t wm = __reduce_or_sync(VB_FULL, mw);
            if (!wm) continue;
#pragma unroll 1
            for (int b0 = 0; b0 < 32; b0 += 8) {
                if (((wm >> b0) & 0xFFu) == 0) continue;
                float lg[8], uo[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const bool p = (mw >> (b0 + j)) & 1u;
                    const size_t v = (size_t)(wi * 32 + b0 + j) * ncols + q;
                    lg[j] = p ? c.logI[v] : 0.0f;
                    uo[j] = p ? c.U[v] : 0.0f;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (!((wm >> (b0 + j)) & 1u)) continue;       // warp uniform
                    const int r = wi * 32 + b0 + j;
                    const int base = ar.bev[r];
                    if ((mw >> (b0 + j)) & 1u) {
                        VB_N4_WINDOW_ADVANCE(base)
                        const float4 w = ar.evw[r];
                        const float bias = w.x * g0 + w.y * g1 + w.z * g2 + w.w * g3;
                        const float un = lg[j] - bias;
                        const float e = expf(un - uo[j]);         // exp(bias_old - bias_new)
                        const double em = (double)e - 1.0;
                        s1 += em;
                        s2 += em * em;
                        tmin = fminf(tmin, un);
                        tmax = fmaxf(tmax, un);
                        c.U[(size_t)r * ncols + q] = un;
                    }
                }
            }
        }
    }
    __syncthreads();     // Gs (union region) is reused by the caller next
    s1 = vb_block_reduce(s1, VbAddD(), 0.0, c.red);
    s2 = vb_block_reduce(s2, VbAddD(), 0.0, c.red);
    bmin = vb_block_reduce(tmin, VbMinF(), FLT_MAX, (float*)c.red);
    bmax = vb_block_reduce(tmax, VbMaxF(), -FLT_MAX, (float*)c.red);
    // CalculateConvergenceMeasurement: sample std / mean of exp(old - new) over the mask
    const double N = (double)M;
    const double mean = 1.0 + s1 / N;
    double var = (s2 - s1 * s1 / N) / (N - 1.0);
    if (var < 0.0) var = 0.0;
    return (float)(sqrt(var) / mean);
}

// final dense pass: bias at EVERY voxel, out = in / exp(bias)   (N4 GenerateData tail)
VB_DEV void n4_finalize(N4Ctx& c, float* logbias) {
    const int nt = (int)blockDim.x, tid = (int)threadIdx.x;
    const int ncols = c.ncols;
    const N4Axis ar = n4_ax(c, 0), ac = n4_ax(c, 1), as = n4_ax(c, 2);
    float* Gs = c.un_f;
    for (int q0 = 0; q0 < ncols; q0 += nt) {
        const int q = q0 + tid;
        const bool valid = q < ncols;
        if (valid) n4_column_coeffs(c, ac, as, q, Gs);
        float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
        int cur = -1;
        if (valid) {
            for (int r0 = 0; r0 < c.n0; r0 += 8) {
                float x[8];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    x[j] = (r0 + j < c.n0) ? vb_ld_stream(c.img + (size_t)(r0 + j) * ncols + q) : 0.0f;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int r = r0 + j;
                    if (r < c.n0) {
                        const int base = ar.bev[r];
                        VB_N4_WINDOW_ADVANCE(base)
                        const float4 w = ar.evw[r];
                        const float bias = w.x * g0 + w.y * g1 + w.z * g2 + w.w * g3;
                        const size_t v = (size_t)r * ncols + q;
                        vb_st_stream(c.U + v, __fdiv_rn(x[j], expf(bias)));
                        if (logbias) vb_st_stream(logbias + v, bias);
                    }
                }
            }
        }
    }
    __syncthreads();
}

// BSplineControlPointImageFilter::RefineControlPointLattice (every axis n -> 2n-3)
VB_DEV void n4_refine(N4Ctx& c) {
    const int nt = (int)blockDim.x, tid = (int)threadIdx.x;
    const float C[2][4] = {{0.5f, 0.5f, 0.0f, 0.0f}, {0.125f, 0.75f, 0.125f, 0.0f}};
    const int o0 = c.ncp[0], o1 = c.ncp[1], o2 = c.ncp[2];
    const int m0 = 2 * o0 - 3, m1 = 2 * o1 - 3, m2 = 2 * o2 - 3;
    float* tmp = c.un_f;
    for (int e = tid; e < m0 * m1 * m2; e += nt) {
        const int k = e % m2, j = (e / m2) % m1, i = e / (m2 * m1);
        const int bi = i >> 1, bj = j >> 1, bk = k >> 1, oi = i & 1, oj = j & 1, ok = k & 1;
        float sum = 0.0f;
        for (int a = 0; a < 4; ++a)
            for (int b = 0; b < 4; ++b)
                for (int g = 0; g < 4; ++g) {
                    if (bi + a >= o0 || bj + b >= o1 || bk + g >= o2) continue;
                    float coeff = 1.0f;
                    coeff = __fmul_rn(coeff, C[ok][g]);
                    coeff = __fmul_rn(coeff, C[oj][b]);
                    coeff = __fmul_rn(coeff, C[oi][a]);
                    sum = __fadd_rn(sum, __fmul_rn(c.L[((bi + a) * o1 + (bj + b)) * o2 + (bk + g)], coeff));
                }
        tmp[e] = sum;
    }
    __syncthreads();
    for (int e = tid; e < m0 * m1 * m2; e += nt) c.L[e] = tmp[e];
    c.ncp[0] = m0; c.ncp[1] = m1; c.ncp[2] = m2;
    __syncthreads();
}

VB_DEV void n4_build_tables(N4Ctx& c) {
    const int n[3] = {c.n0, c.n1, c.n2};
    size_t off = 0;
    for (int a = 0; a < 3; ++a) {
        n4_build_axis(c.axis_base + off, n[a], c.ncp[a]);
        off += 16 * (size_t)n[a];
    }
    __syncthreads();
}

VB_KERNEL void VB_LAUNCH_BOUNDS(VB_N4_THREADS, 1) n4_subject_kernel(N4Args a) {
    VB_DYN_SMEM(smem);
    VB_SHARED int s_next;
    const int tid = (int)threadIdx.x;
    N4Ctx c;
    c.un_f = (float*)smem;
    c.un_d = (double*)smem;
    unsigned char* p = smem + a.union_bytes;
    c.L = (float*)p;  p += sizeof(float) * (size_t)((a.Lmax + 3) & ~3);
    c.Om = (float*)p; p += sizeof(float) * (size_t)((a.Lmax + 3) & ~3);
    c.E = (float*)p;  p += sizeof(float) * (size_t)((a.p.n_bins + 3) & ~3);
    c.H = (float*)p;  p += sizeof(float) * (size_t)((a.p.n_bins + 3) & ~3);
    c.red = (double*)p;
    c.nbins = a.p.n_bins; c.HT = a.hist_threads; c.fftN = a.fftN;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_next = atomicAdd(a.work_counter, 1);
        __syncthreads();
        const int si = s_next;
        if (si >= a.n) break;
        const VbSubject S = a.subj[si];
        c.n0 = S.n0; c.n1 = S.n1; c.n2 = S.n2; c.ncols = S.n1 * S.n2; c.W = (S.n0 + 31) >> 5;
        c.V = S.V;
        c.img = a.image + S.off; c.msk = a.mask + S.off; c.U = a.out + S.off;
        c.logI = a.ws_logI + (size_t)blockIdx.x * a.Vmax;
        c.colmask = a.ws_colmask + (size_t)blockIdx.x * a.cmw_max;
        c.axis_base = a.ws_axis + (size_t)blockIdx.x * a.axis_stride;
        c.ncp[0] = a.p.ncp[0]; c.ncp[1] = a.p.ncp[1]; c.ncp[2] = a.p.ncp[2];

        float bmin, bmax;
        int M;
        n4_setup(c, a.p.mask_label, bmin, bmax, M);
        for (int e = tid; e < a.Lmax; e += (int)blockDim.x) c.L[e] = 0.0f;
        __syncthreads();

        float conv = FLT_MAX;
        int total = 0, ntrace = 0;
        float* dbg = a.debug ? a.debug + (size_t)si * a.debug_stride : nullptr;
        float* dbg_trace = dbg ? dbg + VB_DBG_HDR + 2 * c.nbins + a.Lmax : nullptr;
        const int trace_cap = dbg ? a.debug_stride - (VB_DBG_HDR + 2 * c.nbins + a.Lmax) : 0;
        bool capped = false;
        for (int level = 0; level < a.p.n_levels; ++level) {
            int it = 0;
            if (M >= 2 && !capped) {
                n4_build_tables(c);
                n4_contract<true>(c, bmin, bmax);
                conv = FLT_MAX;
                while (it < a.p.max_iter[level] && conv > a.p.conv_threshold) {
                    if (a.p.max_total_iters > 0 && total >= a.p.max_total_iters) { capped = true; break; }
                    ++it; ++total;
                    if (dbg && tid == 0) { dbg[0] = bmin; dbg[1] = bmax; }
                    n4_histogram(c, bmin, bmax);
                    n4_sharpen(c, bmin, bmax, a.p.fwhm, a.p.wiener_noise);
                    n4_contract<false>(c, bmin, bmax);
                    conv = n4_update(c, M, bmin, bmax);
                    if (dbg_trace && tid == 0 && ntrace < trace_cap) dbg_trace[ntrace] = conv;
                    ++ntrace;
                }
            }
            if (a.stats && tid == 0) a.stats[si].n4_iters[level] = it;
            if (level != a.p.n_levels - 1 && M >= 2 && !capped) n4_refine(c);
        }
        if (M < 2) n4_build_tables(c);     // degenerate mask: zero lattice => out = in
        if (dbg) {
            for (int i = tid; i < c.nbins; i += (int)blockDim.x) {
                dbg[VB_DBG_HDR + i] = c.H[i];
                dbg[VB_DBG_HDR + c.nbins + i] = c.E[i];
            }
            for (int i = tid; i < c.ncp[0] * c.ncp[1] * c.ncp[2]; i += (int)blockDim.x)
                dbg[VB_DBG_HDR + 2 * c.nbins + i] = c.L[i];
            if (tid == 0) {
                dbg[2] = (float)c.ncp[0]; dbg[3] = (float)c.ncp[1]; dbg[4] = (float)c.ncp[2];
                dbg[5] = (float)min(ntrace, trace_cap);
            }
        }
        if (a.stats && tid == 0) {
            for (int l = a.p.n_levels; l < VB_MAX_LEVELS; ++l) a.stats[si].n4_iters[l] = 0;
            a.stats[si].n4_last_conv = conv;
            a.stats[si].n4_n_mask = M;
        }
        n4_finalize(c, a.logbias ? a.logbias + S.off : nullptr);
    }
}
