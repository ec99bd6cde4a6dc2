// SNR, VDP (mean-anchored threshold + 3x3 median), linear binning and border maps.
//
// Replaces, with integer-exact outputs:
//   Vent_Analysis.calculate_SNR      Vent_Analysis.py:337-357
//   Vent_Analysis.calculate_VDP body Vent_Analysis.py:244-257
//   Vent_Analysis.calculateBorder    Vent_Analysis.py:225-231
// (CPU restatement: oracle/vdp_oracle.py.)
#pragma once
#include <float.h>
#include "vb_common.cuh"

#ifndef VB_VDP_THREADS
#define VB_VDP_THREADS 1024
#endif
#define VB_SNR_MAX_AXIS 8192   // n0 + n1 + n2 must not exceed this (projection flags in smem)

// ---------------------------------------------------------------------------------------------
// calculate_SNR: one CTA per subject.  Signal = mask > 0.  Noise = voxels outside
// {rows with mask, plus row 0} x {cols [first masked col > 0, last masked col)} x
// {slices with mask, plus slice 0}, minus the first / last `fov` rows.  Population std.
// 4 consecutive voxels per thread and step (16-byte image loads when the subject is aligned)
// x / d from rd = RN(1/d) with one Newton correction of the quotient: the IEEE-rounded result
// for normal operands (checked against 2e7 random float32 divisions), 3 instructions
VB_DEV float vb_div_rcp(float x, float d, float rd) {
    const float q = __fmul_rn(x, rd);
    return fmaf(fmaf(-q, d, x), rd, q);
}
struct Vb4f { float v[4]; };
struct Vb4b { uint8_t v[4]; };
VB_DEV Vb4f vb_load4f(const float* p, int v, int V, bool aligned) {
    Vb4f r;
    if (aligned && v + 3 < V) {
        const float4 t = *(const float4*)(p + v);
        r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) r.v[e] = (v + e < V) ? p[v + e] : 0.0f;
    }
    return r;
}
VB_DEV Vb4b vb_load4b(const uint8_t* p, int v, int V, bool aligned) {
    Vb4b r;
    if (aligned && v + 3 < V) {
        const uint32_t t = *(const uint32_t*)(p + v);
        r.v[0] = (uint8_t)t; r.v[1] = (uint8_t)(t >> 8); r.v[2] = (uint8_t)(t >> 16); r.v[3] = (uint8_t)(t >> 24);
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) r.v[e] = (v + e < V) ? p[v + e] : (uint8_t)0;
    }
    return r;
}

VB_KERNEL void VB_LAUNCH_BOUNDS(VB_VDP_THREADS, 1)
snr_kernel(const float* VB_RESTRICT image, const uint8_t* VB_RESTRICT mask, const VbSubject* subj, int n,
           int fov, vb_subject_stats* stats) {
    VB_SHARED int flags[VB_SNR_MAX_AXIS];
    VB_SHARED double red[64];
    VB_SHARED int s_cmin, s_cmax;
    const int nt = (int)blockDim.x, tid = (int)threadIdx.x;
    for (int si = (int)blockIdx.x; si < n; si += (int)gridDim.x) {
        const VbSubject S = subj[si];
        const float* img = image + S.off;
        const uint8_t* msk = mask + S.off;
        const bool aligned = (S.off & 3) == 0;
        const int n0 = S.n0, n1 = S.n1, n2 = S.n2, ncols = n1 * n2;
        int* rowf = flags; int* colf = flags + n0; int* slcf = flags + n0 + n1;
        __syncthreads();
        for (int i = tid; i < n0 + n1 + n2; i += nt) flags[i] = 0;
        __syncthreads();
        // fast path (16-byte aligned rows): a thread owns 4 consecutive (col, slice) positions and
        // a row phase and walks down the rows -- no index arithmetic per voxel, coalesced loads
        const bool fast = aligned && (ncols & 3) == 0 && (ncols >> 2) <= nt;
        const int G = ncols >> 2, RP = fast ? max(1, nt / max(G, 1)) : 1;
        const int g = fast ? tid % G : 0, ph = fast ? tid / G : 0;
        const bool mine = fast && ph < RP;
        if (fast) {
            uint32_t colacc = 0;
            if (mine) {
#pragma unroll 4
                for (int i = ph; i < n0; i += RP) {
                    const uint32_t wv = *(const uint32_t*)(msk + (size_t)i * ncols + 4 * g);
                    colacc |= wv;
                    if (wv) atomicOr((unsigned*)&rowf[i], 1u);
                }
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if ((colacc >> (8 * e)) & 0xFFu) {
                        const int rem = 4 * g + e, j = rem / n2, k = rem - j * n2;
                        atomicOr((unsigned*)&colf[j], 1u); atomicOr((unsigned*)&slcf[k], 1u);
                    }
            }
        } else {
            for (int v = 4 * tid; v < S.V; v += 4 * nt) {
                const Vb4b m = vb_load4b(msk, v, S.V, aligned);
                if (m.v[0] | m.v[1] | m.v[2] | m.v[3]) {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (m.v[e]) {
                            const int vv = v + e, i = vv / ncols, rem = vv - i * ncols, j = rem / n2, k = rem - j * n2;
                            // (atomics only so that concurrent same-value writes are not a data race)
                            atomicOr((unsigned*)&rowf[i], 1u); atomicOr((unsigned*)&colf[j], 1u); atomicOr((unsigned*)&slcf[k], 1u);
                        }
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            int cmin = -1, cmax = 0;
            for (int j = 1; j < n1; ++j) if (colf[j]) { if (cmin < 0) cmin = j; }
            for (int j = 0; j < n1; ++j) if (colf[j]) cmax = j;
            s_cmin = cmin; s_cmax = cmax;          // arange(cmin, cmax): cmax excluded
            rowf[0] = 1; slcf[0] = 1;              // "(…>0)*range" keeps index 0
        }
        __syncthreads();
        const int cmin = s_cmin, cmax = s_cmax;
        const int lo = fov < n0 ? fov : n0;        // noisemask[:fov] = 0
        int hi = n0 - fov;                         // noisemask[n0-fov:] = 0 (python slice rules)
        if (hi < 0) { hi += n0; if (hi < 0) hi = 0; }
        // one pass: signal sum / count, noise sum / sum of squares about a pivot / count.  The
        // pivot (any fixed value near the noise level) keeps the one-pass variance well
        // conditioned in float64.
        const float pivot = img[0];
        double ssum = 0.0, nsum = 0.0, nsq = 0.0;
        long long scnt = 0, ncnt = 0;
        if (fast) {
            if (mine) {
                unsigned cs = 0;                   // bit e: (col, slice) of my e-th position is inside the box
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int rem = 4 * g + e, j = rem / n2, k = rem - j * n2;
                    if ((j >= cmin && j < cmax && cmin >= 0) && slcf[k]) cs |= 1u << e;
                }
#pragma unroll 4
                for (int i = ph; i < n0; i += RP) {
                    const float4 xv = *(const float4*)(img + (size_t)i * ncols + 4 * g);
                    const uint32_t wv = *(const uint32_t*)(msk + (size_t)i * ncols + 4 * g);
                    const bool rowin = rowf[i] != 0, band = i >= lo && i < hi;
                    const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if ((wv >> (8 * e)) & 0xFFu) { ssum += (double)xs[e]; ++scnt; }
                        if (band && !(rowin && ((cs >> e) & 1u))) {
                            const double d = (double)xs[e] - (double)pivot;
                            nsum += d; nsq += d * d; ++ncnt;
                        }
                    }
                }
            }
        } else {
            for (int v = 4 * tid; v < S.V; v += 4 * nt) {
                const Vb4f x = vb_load4f(img, v, S.V, aligned);
                const Vb4b m = vb_load4b(msk, v, S.V, aligned);
                int i = v / ncols, rem = v - i * ncols;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (v + e < S.V) {
                        const int j = rem / n2, k = rem - j * n2;
                        if (m.v[e]) { ssum += (double)x.v[e]; ++scnt; }
                        const bool inbox = rowf[i] && (j >= cmin && j < cmax && cmin >= 0) && slcf[k];
                        if (!inbox && i >= lo && i < hi) {
                            const double d = (double)x.v[e] - (double)pivot;
                            nsum += d; nsq += d * d; ++ncnt;
                        }
                    }
                    if (++rem == ncols) { rem = 0; ++i; }
                }
            }
        }
        ssum = vb_block_reduce(ssum, VbAddD(), 0.0, red);
        nsum = vb_block_reduce(nsum, VbAddD(), 0.0, red);
        nsq = vb_block_reduce(nsq, VbAddD(), 0.0, red);
        scnt = vb_block_reduce(scnt, VbAddLL(), 0LL, (long long*)red);
        ncnt = vb_block_reduce(ncnt, VbAddLL(), 0LL, (long long*)red);
        if (tid == 0) {
            vb_subject_stats* st = stats + si;
            const double dm = nsum / (double)ncnt;                 // mean - pivot
            double var = nsq / (double)ncnt - dm * dm;
            if (var < 0.0) var = 0.0;
            const double smean = ssum / (double)scnt, nmean = (double)pivot + dm, nstd = sqrt(var);
            st->signal_mean = smean; st->noise_mean = nmean; st->noise_std = nstd;
            st->snr = (smean - nmean) / nstd;
            st->n_noise = (cmin < 0) ? -1 : (int)ncnt;   // -1: reference raises (np.min of empty)
        }
    }
}

// ---------------------------------------------------------------------------------------------
// VDP statistics, one CTA per subject: M = #mask, mean (float32 of the float64 sum), and the
// exact order statistic sorted(N4[mask>0])[int(M*.99)] by a 3-pass (11+11+10 bit) radix select
// on the order-preserving key of the float bits.  Also zeroes the defect / LB counters.
struct VbRun { int bin; unsigned cnt; };
VB_DEV void vb_run_add(VbRun& r, unsigned* hist, int bin) {
    if (bin == r.bin) { ++r.cnt; return; }
    if (r.cnt) atomicAdd(&hist[r.bin], r.cnt);
    r.bin = bin; r.cnt = 1u;
}
VB_DEV void vb_run_flush(VbRun& r, unsigned* hist) {
    if (r.cnt) atomicAdd(&hist[r.bin], r.cnt);
    r.cnt = 0u;
}

VB_DEV int vdp_find_bin(unsigned* hist, int nb, long long& k, int* s_bin, long long* s_k) {
    // block-wide exclusive prefix over the bins (each thread owns a contiguous run); the thread
    // whose run contains the k-th element publishes (bin, rank inside the bin)
    VB_SHARED int wsum[32];
    const int nt = (int)blockDim.x, tid = (int)threadIdx.x, lane = tid & 31, w = tid >> 5, nw = nt >> 5;
    const int per = (nb + nt - 1) / nt;
    const int lo = min(nb, tid * per), hi = min(nb, lo + per);
    int s = 0;
    for (int b = lo; b < hi; ++b) s += (int)hist[b];
    int inc = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(VB_FULL, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    if (tid == 0) { *s_bin = nb - 1; *s_k = 0; }
    __syncthreads();
    int woff = 0, total = 0;
    for (int i = 0; i < nw; ++i) { const int t = wsum[i]; if (i < w) woff += t; total += t; }
    long long cum = (long long)(woff + inc - s);
    if (k >= (long long)total) {                       // cannot happen for k < count; keep the last bin
        if (tid == 0) { *s_bin = nb - 1; *s_k = k - ((long long)total - (long long)hist[nb - 1]); }
    } else if (k >= cum && k < cum + (long long)s) {
        int b = lo;
        for (; b < hi; ++b) {
            if (cum + (long long)hist[b] > k) break;
            cum += hist[b];
        }
        *s_bin = b; *s_k = k - cum;
    }
    __syncthreads();
    k = *s_k;
    return *s_bin;
}

VB_KERNEL void VB_LAUNCH_BOUNDS(VB_VDP_THREADS, 1)
vdp_stats_kernel(const float* n4, const uint8_t* VB_RESTRICT mask, const VbSubject* subj, int n,
                 vb_subject_stats* stats) {
    VB_SHARED unsigned hist[2048];
    VB_SHARED double red[64];
    VB_SHARED int s_bin;
    VB_SHARED long long s_k;
    const int nt = (int)blockDim.x, tid = (int)threadIdx.x;
    for (int si = (int)blockIdx.x; si < n; si += (int)gridDim.x) {
        const VbSubject S = subj[si];
        const float* x = n4 + S.off;
        const uint8_t* msk = mask + S.off;
        const bool aligned = (S.off & 3) == 0;
        __syncthreads();
        for (int i = tid; i < 2048; i += nt) hist[i] = 0;
        __syncthreads();
        double sum = 0.0;
        long long cnt = 0;
        VbRun run; run.bin = 0; run.cnt = 0u;
        for (int v = 4 * tid; v < S.V; v += 4 * nt) {
            const Vb4b m = vb_load4b(msk, v, S.V, aligned);
            if (!(m.v[0] | m.v[1] | m.v[2] | m.v[3])) continue;
            const Vb4f f = vb_load4f(x, v, S.V, aligned);
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (m.v[e]) {
                    sum += (double)f.v[e]; ++cnt;
                    vb_run_add(run, hist, (int)(vb_float_key(f.v[e]) >> 21));
                }
        }
        vb_run_flush(run, hist);
        sum = vb_block_reduce(sum, VbAddD(), 0.0, red);
        cnt = vb_block_reduce(cnt, VbAddLL(), 0LL, (long long*)red);
        float p99 = 0.0f;
        if (cnt > 0) {
            long long k = (long long)((double)cnt * .99);            // int(len(signal_list)*.99)
            if (k >= cnt) k = cnt - 1;
            const unsigned b1 = (unsigned)vdp_find_bin(hist, 2048, k, &s_bin, &s_k);
            __syncthreads();
            for (int i = tid; i < 2048; i += nt) hist[i] = 0;
            __syncthreads();
            for (int v = 4 * tid; v < S.V; v += 4 * nt) {
                const Vb4b m = vb_load4b(msk, v, S.V, aligned);
                if (!(m.v[0] | m.v[1] | m.v[2] | m.v[3])) continue;
                const Vb4f f = vb_load4f(x, v, S.V, aligned);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (m.v[e]) {
                        const unsigned key = vb_float_key(f.v[e]);
                        if ((key >> 21) == b1) vb_run_add(run, hist, (int)((key >> 10) & 0x7FFu));
                    }
            }
            vb_run_flush(run, hist);
            __syncthreads();
            const unsigned b2 = (unsigned)vdp_find_bin(hist, 2048, k, &s_bin, &s_k);
            __syncthreads();
            for (int i = tid; i < 1024; i += nt) hist[i] = 0;
            __syncthreads();
            const unsigned pre = (b1 << 11) | b2;
            for (int v = 4 * tid; v < S.V; v += 4 * nt) {
                const Vb4b m = vb_load4b(msk, v, S.V, aligned);
                if (!(m.v[0] | m.v[1] | m.v[2] | m.v[3])) continue;
                const Vb4f f = vb_load4f(x, v, S.V, aligned);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (m.v[e]) {
                        const unsigned key = vb_float_key(f.v[e]);
                        if ((key >> 10) == pre) vb_run_add(run, hist, (int)(key & 0x3FFu));
                    }
            }
            vb_run_flush(run, hist);
            __syncthreads();
            const unsigned b3 = (unsigned)vdp_find_bin(hist, 1024, k, &s_bin, &s_k);
            p99 = vb_key_float((pre << 10) | b3);
        }
        if (tid == 0) {
            vb_subject_stats* st = stats + si;
            st->n_mask = (int)cnt;
            st->mean = (cnt > 0) ? (float)(sum / (double)cnt) : 0.0f;
            st->p99 = p99;
            st->n_defect = 0;
            st->n_lb12 = 0;
        }
    }
}

// d0 = (N4/mean < thresh) * mask  (float32 compare, V_A.py:246-249) and the linear-binning code
// (V_A.py:255-256).  grid = (chunks, subjects).
VB_KERNEL void vdp_threshold_kernel(const float* VB_RESTRICT n4, const uint8_t* VB_RESTRICT mask,
                                    const VbSubject* subj, float thresh, uint8_t* d0, uint8_t* lb,
                                    vb_subject_stats* stats) {
    VB_SHARED int red[64];
    const int si = (int)blockIdx.y;
    const VbSubject S = subj[si];
    const float mean = stats[si].mean, p99 = stats[si].p99;
    int c12 = 0;
    for (int v = (int)(blockIdx.x * blockDim.x + threadIdx.x); v < S.V; v += (int)(gridDim.x * blockDim.x)) {
        const float x = n4[S.off + v];
        const bool m = mask[S.off + v] != 0;
        const float a = __fdiv_rn(x, mean);
        d0[S.off + v] = (uint8_t)((a < thresh) && m);
        const float b = __fdiv_rn(x, p99);
        int code = (b <= 0.16f) ? 1 : (b <= 0.34f) ? 2 : (b <= 0.52f) ? 3 : (b <= 0.7f) ? 4 : (b <= 0.88f) ? 5
                 : (b > 0.88f) ? 6 : 0;      // NaN compares false everywhere -> 0, as in numpy
        code = m ? code : 0;
        lb[S.off + v] = (uint8_t)code;
        c12 += (code == 1 || code == 2);
    }
    c12 = vb_block_reduce(c12, VbAddI(), 0, red);
    if (threadIdx.x == 0 && c12) atomicAdd(&stats[si].n_lb12, c12);
}

// defect = medfilt2d(d0[:, :, k]) per slice: zero padded 3x3 neighbour count >= 5 (V_A.py:249)
VB_KERNEL void vdp_median_kernel(const uint8_t* VB_RESTRICT d0, const VbSubject* subj, uint8_t* defect,
                                 vb_subject_stats* stats) {
    VB_SHARED int red[64];
    const int si = (int)blockIdx.y;
    const VbSubject S = subj[si];
    const int n0 = S.n0, n1 = S.n1, n2 = S.n2, ncols = n1 * n2;
    const uint8_t* a = d0 + S.off;
    int nd = 0;
    for (int v = (int)(blockIdx.x * blockDim.x + threadIdx.x); v < S.V; v += (int)(gridDim.x * blockDim.x)) {
        const int i = v / ncols, rem = v - i * ncols, j = rem / n2;
        int cnt = 0;
#pragma unroll
        for (int di = -1; di <= 1; ++di)
#pragma unroll
            for (int dj = -1; dj <= 1; ++dj) {
                const int ii = i + di, jj = j + dj;
                if (ii >= 0 && ii < n0 && jj >= 0 && jj < n1) cnt += a[v + di * ncols + dj * n2];
            }
        const int d = cnt >= 5;
        defect[S.off + v] = (uint8_t)d;
        nd += d;
    }
    nd = vb_block_reduce(nd, VbAddI(), 0, red);
    if (threadIdx.x == 0 && nd) atomicAdd(&stats[si].n_defect, nd);
}

// Fused classification: threshold + linear binning + 3x3 median + defect border in ONE pass.
// A CTA owns rows [i0, i0 + R) of one subject.  It streams N4 and the mask of rows
// [i0 - 2, i0 + R + 2) once (16-byte loads), keeps only the thresholded byte d0 in shared memory,
// forms the median (neighbour count >= 5, zero padded, per slice) for rows [i0 - 1, i0 + R + 1)
// in a second shared tile and from it the border of rows [i0, i0 + R).  Traffic per voxel:
// 5 B read (+ 4/R halo) and 3 B written, instead of three kernels and a d0 round trip.
#ifndef VB_CLS_THREADS
#define VB_CLS_THREADS 512
#endif
VB_KERNEL void VB_LAUNCH_BOUNDS(VB_CLS_THREADS, 2)
vdp_classify_kernel(const float* VB_RESTRICT n4, const uint8_t* VB_RESTRICT mask, const VbSubject* subj, float thresh,
                    int R, uint8_t* lb, uint8_t* defect, uint8_t* border, vb_subject_stats* stats) {
    VB_DYN_SMEM(smem);
    VB_SHARED int red[64];
    const int si = (int)blockIdx.y;
    const VbSubject S = subj[si];
    const int n0 = S.n0, n1 = S.n1, n2 = S.n2, ncols = n1 * n2;
    const int i0 = (int)blockIdx.x * R;
    if (i0 >= n0) return;
    const int nt = (int)blockDim.x, tid = (int)threadIdx.x;
    const int ra = max(i0 - 2, 0), rb = min(i0 + R + 2, n0);          // rows of d0 held
    const int ma = max(i0 - 1, 0), mb = min(i0 + R + 1, n0);          // rows of the median held
    uint8_t* d0 = smem;                                               // [(R + 4) * ncols]
    uint8_t* md = smem + (size_t)(R + 4) * ncols;                     // [(R + 2) * ncols]
    const float mean = stats[si].mean, p99 = stats[si].p99;
    const float rmean = __frcp_rn(mean), rp99 = __frcp_rn(p99);
    const bool okdiv = isfinite(rmean) && isfinite(rp99) && rmean != 0.0f && rp99 != 0.0f;   // else: IEEE division
    const float* x = n4 + S.off;
    const uint8_t* msk = mask + S.off;
    const bool aligned = ((S.off & 3) == 0) && ((ncols & 3) == 0);
    const int v0 = ra * ncols, v1 = rb * ncols;
    const int w0 = i0 * ncols, w1 = min(i0 + R, n0) * ncols;          // voxels this CTA owns
    int c12 = 0;
    // four independent 20-byte loads per thread in flight (the pass is latency bound otherwise)
    for (int vb = v0 + 4 * tid; vb < v1; vb += 16 * nt) {
        Vb4f fu[4];
        Vb4b mu[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int v = vb + 4 * u * nt;
            if (v < v1) { fu[u] = vb_load4f(x, v, v1, aligned); mu[u] = vb_load4b(msk, v, v1, aligned); }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int v = vb + 4 * u * nt;
            if (v >= v1) break;
            const Vb4f f = fu[u];
            const Vb4b m = mu[u];
            uint8_t code[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (v + e < v1) {
                    const bool in = m.v[e] != 0;
                    const bool fin = okdiv && fabsf(f.v[e]) < 3.0e38f;      // inf / nan inputs: IEEE division
                    const float qa = fin ? vb_div_rcp(f.v[e], mean, rmean) : __fdiv_rn(f.v[e], mean);
                    d0[v + e - v0] = (uint8_t)((qa < thresh) && in);
                    const float b = fin ? vb_div_rcp(f.v[e], p99, rp99) : __fdiv_rn(f.v[e], p99);
                    int cd = (b <= 0.16f) ? 1 : (b <= 0.34f) ? 2 : (b <= 0.52f) ? 3 : (b <= 0.7f) ? 4 : (b <= 0.88f) ? 5
                           : (b > 0.88f) ? 6 : 0;      // NaN compares false everywhere -> 0, as in numpy
                    cd = in ? cd : 0;
                    code[e] = (uint8_t)cd;
                    if (v + e >= w0 && v + e < w1) c12 += (cd == 1 || cd == 2);
                } else {
                    code[e] = 0;
                }
            }
            if (v >= w0 && v + 3 < w1 && aligned) {
                *(uint32_t*)(lb + S.off + v) = (uint32_t)code[0] | ((uint32_t)code[1] << 8) | ((uint32_t)code[2] << 16) | ((uint32_t)code[3] << 24);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (v + e >= w0 && v + e < w1) lb[S.off + v + e] = code[e];
            }
        }
    }
    __syncthreads();
    // median of rows [ma, mb): zero padded 3x3 neighbour count >= 5 within the slice.  A thread
    // owns (col, slice) positions and walks down the rows with a sliding window of the three
    // column sums, so every row costs three shared-memory bytes instead of nine.
    int nd = 0;
    const int mv0 = ma * ncols;
    for (int rem = tid; rem < ncols; rem += nt) {
        const int j = rem / n2;
        const bool jl = j > 0, jr = j < n1 - 1;
        // s(i) = d0[i][j-1] + d0[i][j] + d0[i][j+1] (zero outside the volume)
        int s_prev = 0, s_cur = 0;
        if (ma - 1 >= 0) {                                    // row ma - 1 is held whenever it exists
            const uint8_t* p = d0 + (size_t)(ma - 1 - ra) * ncols + rem;
            s_prev = p[0] + (jl ? p[-n2] : 0) + (jr ? p[n2] : 0);
        }
        {
            const uint8_t* p = d0 + (size_t)(ma - ra) * ncols + rem;
            s_cur = p[0] + (jl ? p[-n2] : 0) + (jr ? p[n2] : 0);
        }
        for (int i = ma; i < mb; ++i) {
            int s_next = 0;
            if (i + 1 < n0) {                                 // row i + 1 <= mb is held whenever it exists
                const uint8_t* p = d0 + (size_t)(i + 1 - ra) * ncols + rem;
                s_next = p[0] + (jl ? p[-n2] : 0) + (jr ? p[n2] : 0);
            }
            const int d = (s_prev + s_cur + s_next) >= 5;
            md[(size_t)(i - ma) * ncols + rem] = (uint8_t)d;
            if (i >= i0 && i < i0 + R) { defect[S.off + (size_t)i * ncols + rem] = (uint8_t)d; nd += d; }
            s_prev = s_cur; s_cur = s_next;
        }
    }
    __syncthreads();
    if (border) {
        const int ib = min(i0 + R, n0);
        for (int rem = tid; rem < ncols; rem += nt) {
            const int j = rem / n2;
            const int dl = (j == 0) ? 0 : -n2, dr = (j == n1 - 1) ? 0 : n2;
            for (int i = i0; i < ib; ++i) {
                const uint8_t* a = md + (size_t)(i - ma) * ncols + rem;
                const int up = (i == 0) ? 0 : -ncols, dn = (i == n0 - 1) ? 0 : ncols;
                border[S.off + (size_t)i * ncols + rem] = (uint8_t)((a[dn] != a[up]) || (a[dr] != a[dl]));
            }
        }
    }
    nd = vb_block_reduce(nd, VbAddI(), 0, red);
    c12 = vb_block_reduce(c12, VbAddI(), 0, red);
    if (tid == 0) {
        if (nd) atomicAdd(&stats[si].n_defect, nd);
        if (c12) atomicAdd(&stats[si].n_lb12, c12);
    }
}

// calculateBorder: per slice, np.gradient != 0 along rows OR along columns (V_A.py:225-231)
VB_KERNEL void border_kernel(const uint8_t* VB_RESTRICT a_, const VbSubject* subj, uint8_t* border) {
    const int si = (int)blockIdx.y;
    const VbSubject S = subj[si];
    const int n0 = S.n0, n1 = S.n1, n2 = S.n2, ncols = n1 * n2;
    const uint8_t* a = a_ + S.off;
    for (int v = (int)(blockIdx.x * blockDim.x + threadIdx.x); v < S.V; v += (int)(gridDim.x * blockDim.x)) {
        const int i = v / ncols, rem = v - i * ncols, j = rem / n2;
        // interior: A[i+1] != A[i-1]; edges: one-sided difference
        const int ilo = (i == 0) ? 0 : i - 1, ihi = (i == n0 - 1) ? n0 - 1 : i + 1;
        const int jlo = (j == 0) ? 0 : j - 1, jhi = (j == n1 - 1) ? n1 - 1 : j + 1;
        const bool gr = a[v + (ihi - i) * ncols] != a[v + (ilo - i) * ncols];
        const bool gc = a[v + (jhi - j) * n2] != a[v + (jlo - j) * n2];
        border[S.off + v] = (uint8_t)(gr || gc);
    }
}
