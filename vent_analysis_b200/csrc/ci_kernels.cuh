// Cluster index: exact integer sphere-growing search.
//
// Replaces CI.calculate_CI / CI.calculate_CV (CI.py:87-145) and the percentile of
// Vent_Analysis.calculate_CI (Vent_Analysis.py:268-270).  CPU restatement: oracle/ci_oracle.c.
//
// The reference evaluates the sphere in FLAT, row-fastest index space (px2vec on unchecked
// coordinates, CI.py:65-68,:93): f = i + j*n0 + k*n0*n1, neighbour f' = f + off[row] with
// off = dx + dy*n0 + dz*n0*n1, a hit iff 0 <= f' < V and flat voxel f' is a defect.  For shell
// boundary t (row count ii_t) it stops at the first t with 2 * |unique hits among rows < ii_t|
// < ii_t.  np.intersect1d de-duplicates the hits but the denominator counts table rows, so the
// host keeps only the first occurrence of every flat offset ("unique rows") and remembers, per
// unique row, which boundaries end right after it.
//
// Kernel: the defect map is bit-packed in flat order (28 KB for 128x128x14) and staged in
// shared memory; one warp grows the sphere of one defect voxel, 32 unique rows per step
// (coalesced 8-byte table reads, one shared-memory bit test per lane, ballot + popc prefix),
// checks the boundaries that fall inside the step and exits at the first stop.
#pragma once
#include "vb_common.cuh"

#define VB_CI_THREADS 256
#define VB_CI_SMEM_WORDS_MAX 53248      /* 208 KB of bitset in shared memory; larger -> global */

struct CiTableDev {
    const int2* rows;   // [n_unique]  x = flat offset, y = (first boundary << 16) | boundary count
    const int* ii;      // [n_bounds]  row count at each boundary
    const int* up;      // [ceil(n_unique / 32)]  largest ii of a boundary ending inside the 32-row step (0: none)
    const double* radii; // [n_bounds]  shell radius in min-voxel units (CI.py:105)
    int n_unique;
    int n_bounds;
};

struct CiArgs {
    const uint8_t* defect;        // [V_total] 0/1, C order
    const VbSubject* subj;
    int n;
    const CiTableDev* tables;     // device array
    uint32_t* bits;               // flat bitsets, subject i at word_off[i]
    const long long* word_off;    // [n+1]
    int* list;                    // defect flat indices, subject i at S.off
    int* count;                   // [n] defect counts
    uint16_t* shell;              // [V_total] out
    vb_subject_stats* stats;
};

// flat (row-fastest) bitset + defect list.  grid = (chunks, subjects); one thread per word.
VB_KERNEL void ci_pack_kernel(CiArgs a) {
    const int si = (int)blockIdx.y;
    const VbSubject S = a.subj[si];
    const int n0 = S.n0, n1 = S.n1, n2 = S.n2;
    const int words = (S.V + 31) >> 5;
    const uint8_t* d = a.defect + S.off;
    uint32_t* bits = a.bits + a.word_off[si];
    int* list = a.list + S.off;
    if (blockIdx.x == 0 && threadIdx.x == 0) a.stats[si].ci_error = 0;
    const int ncols = n1 * n2;
    if ((n0 & 31) == 0 && (ncols & 3) == 0 && (S.off & 3) == 0) {
        // Fast path: a flat word is 32 consecutive rows of one (col, slice).  A thread takes four
        // neighbouring (col, slice) positions and walks the 32 rows of one slab with 4-byte
        // loads, so a warp reads 128 contiguous bytes per row (the volume is read exactly once,
        // coalesced) instead of 32 single bytes one row stride apart per word.
        const int G = ncols >> 2, slabs = n0 >> 5, W0 = n0 >> 5;
        for (int t = (int)(blockIdx.x * blockDim.x + threadIdx.x); t < G * slabs; t += (int)(gridDim.x * blockDim.x)) {
            const int g = t % G, sl = t / G;
            const uint8_t* p = d + (size_t)(sl * 32) * ncols + 4 * g;
            uint32_t acc[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int b0 = 0; b0 < 32; b0 += 8) {
                uint32_t v[8];
#pragma unroll
                for (int b = 0; b < 8; ++b) v[b] = *(const uint32_t*)(p + (size_t)(b0 + b) * ncols);
                // byte e of t: bit b = "row b0 + b of position e is set" (eight rows per word)
                uint32_t t = 0u;
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const uint32_t w = v[b];
                    const uint32_t nz = ((((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w) & 0x80808080u) >> 7;   // 0/1 per byte
                    t |= nz << b;
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[e] |= ((t >> (8 * e)) & 0xFFu) << b0;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int q = 4 * g + e, j = q / n2, k = q - j * n2;
                const int w = sl + W0 * (j + n1 * k);
                bits[w] = acc[e];
                if (acc[e]) {
                    int pos = atomicAdd(&a.count[si], __popc(acc[e]));
                    uint32_t m = acc[e];
                    while (m) {
                        const int b = __ffs((int)m) - 1;
                        m &= m - 1;
                        list[pos++] = (w << 5) + b;
                    }
                }
            }
        }
        return;
    }
    for (int w = (int)(blockIdx.x * blockDim.x + threadIdx.x); w < words; w += (int)(gridDim.x * blockDim.x)) {
        uint32_t word = 0;
        int f = w << 5;
        int i = f % n0, jk = f / n0, j = jk % n1, k = jk / n1;
        for (int b = 0; b < 32 && f < S.V; ++b, ++f) {
            if (d[((size_t)i * n1 + j) * n2 + k]) word |= 1u << b;
            if (++i == n0) { i = 0; if (++j == n1) { j = 0; ++k; } }
        }
        bits[w] = word;
        if (word) {
            int pos = atomicAdd(&a.count[si], __popc(word));
            uint32_t m = word;
            while (m) {
                const int b = __ffs((int)m) - 1;
                m &= m - 1;
                list[pos++] = (w << 5) + b;
            }
        }
    }
}

// One warp per defect voxel.  grid = (ctas_per_subject, subjects).
template <bool SMEM_BITS>
VB_KERNEL void VB_LAUNCH_BOUNDS(VB_CI_THREADS, 1) ci_grow_kernel(CiArgs a) {
    VB_DYN_SMEM(smem);
    const int si = (int)blockIdx.y;
    const VbSubject S = a.subj[si];
    const int D = a.count[si];
    const int nw = (int)(blockDim.x >> 5), wid = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
    const int words = (S.V + 31) >> 5;
    const uint32_t* gbits = a.bits + a.word_off[si];
    const uint32_t* bits = gbits;
    if (SMEM_BITS) {
        uint32_t* sb = (uint32_t*)smem;
        if ((int)blockIdx.x * nw < D)               // CTAs without work skip the staging
            for (int w = (int)threadIdx.x; w < words; w += (int)blockDim.x) sb[w] = gbits[w];
        bits = sb;
        __syncthreads();
    }
    const CiTableDev T = a.tables[S.table];
    const int* list = a.list + S.off;
    const int n0 = S.n0, n1 = S.n1, n2 = S.n2;
    const unsigned V = (unsigned)S.V;
    for (int di = (int)blockIdx.x * nw + wid; di < D; di += (int)gridDim.x * nw) {
        const int f = list[di];
        int cum = 0, found = -1;
        for (int base = 0; base < T.n_unique; base += 32) {
            const int u = base + lane;
            int2 row = make_int2(0, 0);
            bool hit = false;
            if (u < T.n_unique) {
                row = T.rows[u];
                const unsigned g = (unsigned)(f + row.x);       // negative wraps to >= V
                if (g < V) hit = (bits[g >> 5] >> (g & 31u)) & 1u;
            }
            const unsigned hm = __ballot_sync(VB_FULL, hit);
            // A boundary t of this step can only stop the search if 2 * hits < ii[t], and the hits
            // up to any row of the step are >= cum: with up = the largest ii[t] of the step, nothing
            // can stop while 2 * cum >= up (most steps inside a cluster) -- warp uniform.
            const int upv = T.up[base >> 5];
            if (2 * cum < upv) {
                const int pre = cum + __popc(hm & (0xffffffffu >> (31 - lane)));   // inclusive prefix
                int cand = -1;
                const int nb = row.y & 0xFFFF;
                if (nb) {                                            // boundaries that end after row u
                    const int t0 = (int)((unsigned)row.y >> 16);
                    for (int t = t0; t < t0 + nb; ++t)
                        if (2 * pre < T.ii[t]) { cand = t; break; }  // C < 0.5  (CI.py:97)
                }
                const unsigned cm = __ballot_sync(VB_FULL, cand >= 0);
                if (cm) {
                    found = __shfl_sync(VB_FULL, cand, __ffs((int)cm) - 1);
                    break;
                }
            }
            cum += __popc(hm);
        }
        if (lane == 0) {
            const int i = f % n0, jk = f / n0, j = jk % n1, k = jk / n1;
            const size_t c = ((size_t)i * n1 + j) * n2 + k;
            if (found < 0) {
                a.stats[si].ci_error = 1;                        // reference: ValueError (CI.py:101-103)
                a.shell[S.off + c] = (uint16_t)0xFFFE;
            } else {
                a.shell[S.off + c] = (uint16_t)found;
            }
        }
    }
}

// sort(CIarray[defect>0])[int(0.95*D)] on the shell ordinal (the radius is strictly increasing in
// it): counting sort in shared memory, one CTA per subject.
VB_KERNEL void VB_LAUNCH_BOUNDS(VB_CI_THREADS, 1) ci_percentile_kernel(CiArgs a) {
    VB_DYN_SMEM(smem);
    unsigned* hist = (unsigned*)smem;
    const int nt = (int)blockDim.x, tid = (int)threadIdx.x;
    for (int si = (int)blockIdx.x; si < a.n; si += (int)gridDim.x) {
        const VbSubject S = a.subj[si];
        const CiTableDev T = a.tables[S.table];
        const int D = a.count[si];
        const int* list = a.list + S.off;
        const int n0 = S.n0, n1 = S.n1, n2 = S.n2;
        __syncthreads();
        for (int i = tid; i < T.n_bounds; i += nt) hist[i] = 0;
        __syncthreads();
        for (int di = tid; di < D; di += nt) {
            const int f = list[di];
            const int i = f % n0, jk = f / n0, j = jk % n1, k = jk / n1;
            const unsigned t = a.shell[S.off + ((size_t)i * n1 + j) * n2 + k];
            if (t < (unsigned)T.n_bounds) atomicAdd(&hist[t], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            vb_subject_stats* st = a.stats + si;
            st->ci_n = D;
            int t95 = -1;
            if (D > 0 && !st->ci_error) {
                const long long k = (long long)(0.95 * (double)D);     // int(0.95*len(CVlist))
                long long cum = 0;
                for (int t = 0; t < T.n_bounds; ++t) {
                    cum += hist[t];
                    if (cum > k) { t95 = t; break; }
                }
            }
            st->ci_t95 = t95;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Vent_Analysis.build4DdataArray (Vent_Analysis.py:292-313): the six result volumes of a subject
// interleaved as float32 [row][col][slice][6]: proton[0] HPvent[1] mask[2] N4HPvent[3]
// defectArray[4] CIarray[5], CIarray = float32(radius[shell] * min(vox)) on defect voxels, else 0
// (the float64 product the reference stores, then its float32 cast).  A channel whose input is
// null stays 0, as when the reference's attribute is missing.  One voxel per thread and step:
// reads are coalesced per channel, a warp writes 768 contiguous bytes (three 8-byte stores each).
struct Pack6Args {
    const float* proton;
    const float* image;
    const uint8_t* mask;
    const float* n4;
    const uint8_t* defect;
    const uint16_t* shell;
    const VbSubject* subj;
    const CiTableDev* tables;     // may be null when shell is null
    const double* vox_min;        // [n]
    float* out;                   // [6 * V_total]
};
VB_KERNEL void pack6_kernel(Pack6Args a) {
    const int si = (int)blockIdx.y;
    const VbSubject S = a.subj[si];
    const double vmin = a.vox_min[si];
    const double* radii = nullptr;
    int nb = 0;
    if (a.shell && S.table >= 0) { radii = a.tables[S.table].radii; nb = a.tables[S.table].n_bounds; }
    for (int v = (int)(blockIdx.x * blockDim.x + threadIdx.x); v < S.V; v += (int)(gridDim.x * blockDim.x)) {
        const size_t g = (size_t)S.off + (size_t)v;
        const float c0 = a.proton ? a.proton[g] : 0.0f;
        const float c1 = a.image ? a.image[g] : 0.0f;
        const float c2 = a.mask ? (float)a.mask[g] : 0.0f;
        const float c3 = a.n4 ? a.n4[g] : 0.0f;
        const float c4 = a.defect ? (float)a.defect[g] : 0.0f;
        float c5 = 0.0f;
        if (radii) {
            const int t = (int)a.shell[g];
            if (t < nb) c5 = (float)__dmul_rn(radii[t], vmin);
        }
        float2* o = (float2*)(a.out + 6 * g);
        o[0] = make_float2(c0, c1); o[1] = make_float2(c2, c3); o[2] = make_float2(c4, c5);
    }
}
