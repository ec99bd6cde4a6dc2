// k-space reconstruction of Vent_Analysis.process_RAW (Vent_Analysis.py:536-540):
//   for every slice k:  img[:, :, k] = fftshift(fft2(fftshift(K[:, :, k])))
//   raw_HPvent = transpose(img, (1, 0, 2))[:, ::-1, :]
// i.e. with F = fft2 of the shifted slice, s0 = n0 // 2, s1 = n1 // 2:
//   out[c][rr][k] = F[(n0 - 1 - rr - s0) mod n0][(c - s1) mod n1].
// (CPU restatement: oracle/recon_oracle.py.)  Two passes, one per axis; all three index
// permutations (input shift, output shift, transpose + flip) are folded into the loads of pass 1 /
// pass 2 and the store of pass 2, so the data cross HBM four times in all (32 bytes per voxel).
// Slices are the fastest axis of the volume.  Pass 1 (along rows) does not care which (col, slice) an
// inner index belongs to: a CTA takes `nl` CONSECUTIVE inner positions, so every row it touches
// is one contiguous run of 8 nl bytes.  Pass 2 (along columns) takes several frequency rows x all
// slices per CTA (runs of 8 n2 bytes).  The nl 1-D transforms of a tile run side by side in shared
// memory -- radix-2 for power-of-two lengths, a direct DFT with a shared twiddle table otherwise
// (any length; DICOM matrices like 208 are not powers of two).
#pragma once
#include "vb_common.cuh"

#define VB_RECON_THREADS 256
#define VB_RECON_NL_MAX 64         /* transforms per CTA at most */
struct ReconArgs {
    const float2* src;            // pass 1: k-space [n0][n1][n2]; pass 2: the axis-0 transform
    float2* dst;                  // pass 1: temp [n0][n1][n2]; pass 2: image [n1][n0][n2]
    const VbSubject* subj;
    int pass;                     // 0: along rows (axis 0); 1: along columns (axis 1) + output permutation
    int nl;                       // transforms per CTA
    int need_y;                   // a second tile buffer follows the first (non power-of-two lengths)
};

// in-place forward DFT of `nl` interleaved lines of length N: element i of line l at x[i * nl + l].
// tw[j] = exp(-2 pi i j / N).  y: scratch of the same size (only used when N is not a power of two);
// returns the buffer that holds the result.
VB_DEV float2* recon_dft_lines(float2* x, float2* y, const float2* tw, int N, int nl) {
    const int nt = (int)blockDim.x, tid = (int)threadIdx.x;
    // a thread keeps ONE lane l and strides over rows: no division inside the loops (the lane count is
    // not a power of two in general; the butterfly index splits by shifts)
    const int rpi = max(1, nt / nl);                     // rows per iteration
    const int l = tid % nl, r0 = tid / nl;
    const bool act = r0 < rpi;
    if ((N & (N - 1)) == 0) {
        int logN = 0;
        while ((1 << logN) < N) ++logN;
        if (act && logN > 0)
            for (int i = r0; i < N; i += rpi) {                  // bit reversal (swap once per pair)
                const int j = (int)(__brev((unsigned)i) >> (32 - logN));
                if (i < j) { const float2 t = x[i * nl + l]; x[i * nl + l] = x[j * nl + l]; x[j * nl + l] = t; }
            }
        __syncthreads();
        int lh = 0;                                              // log2(half)
        for (int len = 2; len <= N; len <<= 1, ++lh) {
            const int half = len >> 1, step = N >> (lh + 1);
            if (act)
                for (int b = r0; b < (N >> 1); b += rpi) {
                    const int grp = b >> lh, k = b & (half - 1);
                    const int i0 = ((grp << (lh + 1)) + k) * nl + l, i1 = i0 + half * nl;
                    const float2 w = tw[k * step], a = x[i0], c = x[i1];
                    const float xr = c.x * w.x - c.y * w.y, xi = c.x * w.y + c.y * w.x;
                    x[i0] = make_float2(a.x + xr, a.y + xi);
                    x[i1] = make_float2(a.x - xr, a.y - xi);
                }
            __syncthreads();
        }
        return x;
    }
    if (act)
        for (int u = r0; u < N; u += rpi) {
            float sr = 0.0f, si = 0.0f;
            int ph = 0;                                          // (u * i) mod N
            for (int i = 0; i < N; ++i) {
                const float2 w = tw[ph], a = x[i * nl + l];
                sr += a.x * w.x - a.y * w.y;
                si += a.x * w.y + a.y * w.x;
                ph += u;
                if (ph >= N) ph -= N;
            }
            y[u * nl + l] = make_float2(sr, si);
        }
    __syncthreads();
    return y;
}

// grid = (tiles, 1, subjects); dynamic shared memory: ((1 + need_y) * Nmax * nl + Nmax) float2
VB_KERNEL void VB_LAUNCH_BOUNDS(VB_RECON_THREADS, 2) recon_pass_kernel(ReconArgs a, int nmax) {
    VB_DYN_SMEM(smem);
    const int si = (int)blockIdx.z;
    const VbSubject S = a.subj[si];
    const int n0 = S.n0, n1 = S.n1, n2 = S.n2, inner = n1 * n2;
    const int N = a.pass == 0 ? n0 : n1;
    // pass 1: tile = inner positions [q0, q0 + nl).  pass 2: tile = LU frequency rows x KS slices.
    const int KS = min(n2, a.nl), LU = max(1, a.nl / KS), ktiles = (n2 + KS - 1) / KS;
    int nl, q0 = 0, u0 = 0, k0 = 0;
    if (a.pass == 0) {
        q0 = (int)blockIdx.x * a.nl;
        if (q0 >= inner) return;
        nl = min(a.nl, inner - q0);
    } else {
        const int ut = (int)blockIdx.x / ktiles, kt = (int)blockIdx.x - ut * ktiles;
        u0 = ut * LU; k0 = kt * KS;
        if (u0 >= n0) return;
        nl = min(LU, n0 - u0) * KS;                      // (lanes of a partial last slice tile are masked below)
    }
    float2* x = (float2*)smem;
    float2* y = x + (a.need_y ? (size_t)nmax * a.nl : 0);
    float2* tw = y + (size_t)nmax * a.nl;
    const int nt = (int)blockDim.x, tid = (int)threadIdx.x;
    for (int j = tid; j < N; j += nt) {
        double s, c;
        vb_sincospi(-2.0 * (double)j / (double)N, &s, &c);
        tw[j] = make_float2((float)c, (float)s);
    }
    const float2* src = a.src + S.off;
    const int s0 = n0 >> 1, s1 = n1 >> 1;
    // a thread keeps one lane of the tile and strides over rows (no division per element)
    const int rpi = max(1, nt / nl), l = tid % nl, r0 = tid / nl;
    const bool act = r0 < rpi;
    const int du = l / KS, k = k0 + (l - du * KS);
    // load with the input fftshift of THIS axis: a[i] = K[(i - s) mod N]
    if (act)
        for (int i = r0; i < N; i += rpi) {
            float2 v = make_float2(0.0f, 0.0f);
            if (a.pass == 0) {
                int r = i - s0; if (r < 0) r += n0;
                v = src[(size_t)r * inner + q0 + l];
            } else {
                int cc = i - s1; if (cc < 0) cc += n1;
                if (k < n2) v = src[((size_t)(u0 + du) * n1 + cc) * n2 + k];
            }
            x[i * nl + l] = v;
        }
    __syncthreads();
    const float2* res = recon_dft_lines(x, y, tw, N, nl);
    float2* dst = a.dst + S.off;
    if (act)
        for (int u = r0; u < N; u += rpi) {
            if (a.pass == 0) {
                dst[(size_t)u * inner + q0 + l] = res[u * nl + l];            // temp[u][col][k]
            } else if (k < n2) {
                // F[line][u] -> out[c][rr][k], c = (u + s1) mod n1, rr = n0 - 1 - ((line + s0) mod n0)
                int c = u + s1; if (c >= n1) c -= n1;
                int t = u0 + du + s0; if (t >= n0) t -= n0;
                dst[((size_t)c * n0 + (n0 - 1 - t)) * n2 + k] = res[u * nl + l];
            }
        }
}
