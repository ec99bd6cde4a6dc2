// k-space reconstruction of Vent_Analysis.process_RAW (Vent_Analysis.py:536-540):
//   for every slice k:  img[:, :, k] = fftshift(fft2(fftshift(K[:, :, k])))
//   raw_HPvent = transpose(img, (1, 0, 2))[:, ::-1, :]
// i.e. with F = fft2 of the shifted slice, s0 = n0 // 2, s1 = n1 // 2:
//   out[c][rr][k] = F[(n0 - 1 - rr - s0) mod n0][(c - s1) mod n1].
// (CPU restatement: oracle/recon_oracle.py.)  Two passes, one per axis; all three index
// permutations (input shift, output shift, transpose + flip) are folded into the loads of pass 1 /
// pass 2 and the store of pass 2, so the data cross HBM four times in all (32 bytes per voxel).
// Slices are the fastest axis of the volume: a CTA takes one line (a column in pass 1, a frequency
// row in pass 2) for a tile of up to VB_RECON_NL slices, so every global access is a contiguous
// run of slices, and runs the 1-D transforms of the tile's slices side by side in shared memory --
// radix-2 for power-of-two lengths, a direct DFT with a shared twiddle table otherwise (any
// length; DICOM matrices like 208 are not powers of two).
#pragma once
#include "vb_common.cuh"

#define VB_RECON_THREADS 256
#define VB_RECON_NL 16            /* slices per tile */

struct ReconArgs {
    const float2* src;            // pass 1: k-space [n0][n1][n2]; pass 2: the axis-0 transform
    float2* dst;                  // pass 1: temp [n0][n1][n2]; pass 2: image [n1][n0][n2]
    const VbSubject* subj;
    int pass;                     // 0: along rows (axis 0); 1: along columns (axis 1) + output permutation
};

// in-place forward DFT of `nl` interleaved lines of length N: element i of line l at x[i * nl + l].
// tw[j] = exp(-2 pi i j / N).  y: scratch of the same size (only used when N is not a power of two);
// returns the buffer that holds the result.
VB_DEV float2* recon_dft_lines(float2* x, float2* y, const float2* tw, int N, int nl) {
    const int nt = (int)blockDim.x, tid = (int)threadIdx.x;
    if ((N & (N - 1)) == 0) {
        int logN = 0;
        while ((1 << logN) < N) ++logN;
        for (int e = tid; e < N * nl; e += nt) {                 // bit reversal (swap once per pair)
            const int i = e / nl, l = e - i * nl;
            const int j = (int)(__brev((unsigned)i) >> (32 - logN));
            if (logN > 0 && i < j) { const float2 t = x[i * nl + l]; x[i * nl + l] = x[j * nl + l]; x[j * nl + l] = t; }
        }
        __syncthreads();
        for (int len = 2; len <= N; len <<= 1) {
            const int half = len >> 1, step = N / len;
            for (int e = tid; e < (N >> 1) * nl; e += nt) {
                const int b = e / nl, l = e - b * nl;
                const int grp = b / half, k = b - grp * half;
                const int i0 = (grp * len + k) * nl + l, i1 = i0 + half * nl;
                const float2 w = tw[k * step], a = x[i0], c = x[i1];
                const float xr = c.x * w.x - c.y * w.y, xi = c.x * w.y + c.y * w.x;
                x[i0] = make_float2(a.x + xr, a.y + xi);
                x[i1] = make_float2(a.x - xr, a.y - xi);
            }
            __syncthreads();
        }
        return x;
    }
    for (int e = tid; e < N * nl; e += nt) {
        const int u = e / nl, l = e - u * nl;
        float sr = 0.0f, si = 0.0f;
        int ph = 0;                                              // (u * i) mod N
        for (int i = 0; i < N; ++i) {
            const float2 w = tw[ph], a = x[i * nl + l];
            sr += a.x * w.x - a.y * w.y;
            si += a.x * w.y + a.y * w.x;
            ph += u;
            if (ph >= N) ph -= N;
        }
        y[e] = make_float2(sr, si);
    }
    __syncthreads();
    return y;
}

// grid = (lines, slice tiles, subjects); dynamic shared memory: (2 * Nmax * VB_RECON_NL + Nmax) float2
VB_KERNEL void VB_LAUNCH_BOUNDS(VB_RECON_THREADS, 2) recon_pass_kernel(ReconArgs a, int nmax) {
    VB_DYN_SMEM(smem);
    const int si = (int)blockIdx.z;
    const VbSubject S = a.subj[si];
    const int n0 = S.n0, n1 = S.n1, n2 = S.n2;
    const int N = a.pass == 0 ? n0 : n1, lines = a.pass == 0 ? n1 : n0;
    const int line = (int)blockIdx.x, k0 = (int)blockIdx.y * VB_RECON_NL;
    if (line >= lines || k0 >= n2) return;
    const int nl = min(VB_RECON_NL, n2 - k0);
    float2* x = (float2*)smem;
    float2* y = x + (size_t)nmax * VB_RECON_NL;
    float2* tw = y + (size_t)nmax * VB_RECON_NL;
    const int nt = (int)blockDim.x, tid = (int)threadIdx.x;
    for (int j = tid; j < N; j += nt) {
        double s, c;
        vb_sincospi(-2.0 * (double)j / (double)N, &s, &c);
        tw[j] = make_float2((float)c, (float)s);
    }
    const float2* src = a.src + S.off;
    const int s0 = n0 >> 1, s1 = n1 >> 1;
    // load with the input fftshift of THIS axis: a[i] = K[(i - s) mod N]
    for (int e = tid; e < N * nl; e += nt) {
        const int i = e / nl, l = e - i * nl;
        size_t g;
        if (a.pass == 0) { int r = i - s0; if (r < 0) r += n0; g = ((size_t)r * n1 + line) * n2 + k0 + l; }
        else { int cc = i - s1; if (cc < 0) cc += n1; g = ((size_t)line * n1 + cc) * n2 + k0 + l; }
        x[i * nl + l] = src[g];
    }
    __syncthreads();
    const float2* res = recon_dft_lines(x, y, tw, N, nl);
    float2* dst = a.dst + S.off;
    for (int e = tid; e < N * nl; e += nt) {
        const int u = e / nl, l = e - u * nl;
        size_t g;
        if (a.pass == 0) {
            g = ((size_t)u * n1 + line) * n2 + k0 + l;                        // temp[u][column][k]
        } else {
            // F[line][u] -> out[c][rr][k], c = (u + s1) mod n1, rr = n0 - 1 - ((line + s0) mod n0)
            int c = u + s1; if (c >= n1) c -= n1;
            int t = line + s0; if (t >= n0) t -= n0;
            g = ((size_t)c * n0 + (n0 - 1 - t)) * n2 + k0 + l;
        }
        dst[g] = res[e];
    }
}
