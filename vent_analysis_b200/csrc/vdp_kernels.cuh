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
#ifndef VB_SNR_THREADS
/* threads per CTA of snr_kernel (<= VB_VDP_THREADS): 512 = two CTAs per SM, so that one CTA's mask
   pass / barriers overlap the other's image pass (measured 5 % faster than one CTA of 1024) */
#define VB_SNR_THREADS (VB_VDP_THREADS >= 1024 ? 512 : VB_VDP_THREADS)
#endif
#define VB_SNR_MAX_AXIS 3072   // n0 + n1 + n2 must not exceed this (projection flags in smem)

// ---------------------------------------------------------------------------------------------
// calculate_SNR: one CTA per subject.  Signal = mask > 0.  Noise = voxels outside
// {rows with mask, plus row 0} x {cols [first masked col > 0, last masked col)} x
// {slices with mask, plus slice 0}, minus the first / last `fov` rows.  Population std.
// 4 consecutive voxels per thread and step (16-byte image loads when the subject is aligned)
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

// four mask bytes as one word (0 beyond the volume); bit 7 of every non-zero byte
VB_DEV uint32_t vb_load4b_word(const uint8_t* p, int v, int V, bool aligned) {
    if (aligned && v + 3 < V) return *(const uint32_t*)(p + v);
    uint32_t w = 0u;
#pragma unroll
    for (int e = 0; e < 4; ++e)
        if (v + e < V) w |= (uint32_t)p[v + e] << (8 * e);
    return w;
}
VB_DEV uint32_t vb_nonzero_bytes(uint32_t w) {
    return ((((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w) & 0x80808080u);
}

// Row streaming through a ring of shared-memory stages filled by bulk asynchronous copies (TMA,
// cp.async.bulk + mbarrier), warp specialised: lane 0 of the PRODUCER warp (the first warp that owns
// no column group) keeps VB_SNR_STAGES chunks of `R` rows in flight; the consumer warps wait on a
// stage's "full" barrier, reduce its rows and hand the stage back through its "free" barrier (one
// arrival per warp) -- no block-wide barrier inside a pass.  `chunk` counts the chunks this CTA
// has streamed so far (it selects stage and barrier phase, so it runs on across passes and
// subjects).  BODY(row0, rows, simg, smsk) runs on the consumer warps.
#ifndef VB_SNR_STAGES
#define VB_SNR_STAGES 3      /* 3 stages of three C2 rows measured 0.50 ms per 1776 subjects; 4 x 2 rows 0.55, 6 x 1 row 0.61; direct loads 0.55 */
#endif
struct SnrRing {
    unsigned char* base;          // VB_SNR_STAGES stages of `stage_bytes`
    vb_mbar_t* full;
    vb_mbar_t* free_;
    int stage_bytes, R, ncw;      // consumer warps
    unsigned chunk;
};
template <class F>
VB_DEV void snr_stream(SnrRing& g, const float* img, const uint8_t* msk, int n0, int ncols, F body) {
    g.R = max(1, min(n0, g.stage_bytes / ((img ? 5 : 1) * ncols)));     // rows per stage: mask-only passes take five times more
    const int nch = (n0 + g.R - 1) / g.R, w = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
    const unsigned c0 = g.chunk;
    if (w == g.ncw) {                                     // producer
        if (lane == 0)
            for (int c = 0; c < nch; ++c) {
                const unsigned cg = c0 + (unsigned)c, st = cg % VB_SNR_STAGES;
                vb_free_wait(&g.free_[st], cg / VB_SNR_STAGES, g.ncw);
                const int row0 = c * g.R, rows = min(g.R, n0 - row0);
                unsigned char* dst = g.base + (size_t)st * g.stage_bytes;
                const unsigned mb = (unsigned)(rows * ncols), ib = img ? 4u * mb : 0u;
                vb_stage_begin(&g.full[st], mb + ib);
                if (img) vb_bulk_g2s(dst, img + (size_t)row0 * ncols, ib, &g.full[st]);
                vb_bulk_g2s(dst + (img ? 4 * (size_t)g.R * ncols : 0), msk + (size_t)row0 * ncols, mb, &g.full[st]);
                vb_stage_end(&g.full[st]);
            }
    } else if (w < g.ncw) {                               // consumers
        for (int c = 0; c < nch; ++c) {
            const unsigned cg = c0 + (unsigned)c, st = cg % VB_SNR_STAGES;
            vb_stage_wait(&g.full[st], cg / VB_SNR_STAGES);
            const unsigned char* sp = g.base + (size_t)st * g.stage_bytes;
            body(c * g.R, min(g.R, n0 - c * g.R), (const float*)sp, (const uint8_t*)(sp + (img ? 4 * (size_t)g.R * ncols : 0)));
            __syncwarp();
            if (lane == 0) vb_free_arrive(&g.free_[st]);
        }
    }
    g.chunk = c0 + (unsigned)nch;
    __syncthreads();
}

VB_KERNEL void VB_LAUNCH_BOUNDS(VB_VDP_THREADS, 1)
snr_kernel(const float* VB_RESTRICT image, const uint8_t* VB_RESTRICT mask, const VbSubject* subj, int n,
           int fov, vb_subject_stats* stats, int ring_bytes) {
    VB_DYN_SMEM(ring_smem);
    VB_SHARED int flags[VB_SNR_MAX_AXIS];
    VB_SHARED double red[64];
    VB_SHARED int s_cmin, s_cmax;
    VB_SHARED vb_mbar_t s_full[VB_SNR_STAGES], s_free[VB_SNR_STAGES];
    const int nt = (int)blockDim.x, tid = (int)threadIdx.x;
    SnrRing ring;
    ring.base = ring_smem; ring.full = s_full; ring.free_ = s_free; ring.chunk = 0u; ring.stage_bytes = 0; ring.R = 1; ring.ncw = 0;
    int ring_ncw = -1;                                   // consumer-warp count the "free" barriers were set up for
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
        // bulk path: whole rows as 16-byte multiples at a 16-byte aligned offset, and a stage of at
        // least one row must fit the ring
        const int stage_bytes = (ring_bytes / VB_SNR_STAGES) & ~15;
        const int stage_rows = fast ? stage_bytes / (5 * ncols) : 0;
        const int ncw = fast ? (G * RP + 31) >> 5 : 0;            // warps that own column groups; the next one produces
        const bool bulk = stage_rows >= 1 && (S.off & 15) == 0 && (ncols & 15) == 0 && ncw < (nt >> 5);
        if (bulk) {
            ring.stage_bytes = stage_bytes;
            ring.ncw = ncw;
            if (ring_ncw != ncw) {                                // (block uniform) first use / another geometry
                __syncthreads();
                if (tid == 0) {
                    for (int k = 0; k < VB_SNR_STAGES; ++k) { vb_mbar_init(&s_full[k], 1); vb_mbar_init(&s_free[k], ncw); }
                    vb_mbar_fence_init();
                }
                ring.chunk = 0u;
                ring_ncw = ncw;
                __syncthreads();
            }
        }
        if (bulk) {
            uint32_t colacc = 0;
            snr_stream(ring, (const float*)nullptr, msk, n0, ncols, [&](int row0, int rows, const float*, const uint8_t* sm) {
                if (mine)
                    for (int i = ph; i < rows; i += RP) {
                        const uint32_t wv = *(const uint32_t*)(sm + (size_t)i * ncols + 4 * g);
                        colacc |= wv;
                        if (wv) atomicOr((unsigned*)&rowf[row0 + i], 1u);
                    }
            });
            if (mine) {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if ((colacc >> (8 * e)) & 0xFFu) {
                        const int rem = 4 * g + e, j = rem / n2, k = rem - j * n2;
                        atomicOr((unsigned*)&colf[j], 1u); atomicOr((unsigned*)&slcf[k], 1u);
                    }
            }
        } else if (fast) {
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
        if (bulk) {
            unsigned cs = 0;                       // bit e: (col, slice) of my e-th position is inside the box
            if (mine) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int rem = 4 * g + e, j = rem / n2, k = rem - j * n2;
                    if ((j >= cmin && j < cmax && cmin >= 0) && slcf[k]) cs |= 1u << e;
                }
            }
            snr_stream(ring, img, msk, n0, ncols, [&](int row0, int rows, const float* sx, const uint8_t* sm) {
                if (mine)
                    for (int ii = ph; ii < rows; ii += RP) {
                        const int i = row0 + ii;
                        const float4 xv = *(const float4*)(sx + (size_t)ii * ncols + 4 * g);
                        const uint32_t wv = *(const uint32_t*)(sm + (size_t)ii * ncols + 4 * g);
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
            });
        } else if (fast) {
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
// VDP statistics, one CTA per subject: M = #mask and the two anchors of Vent_Analysis.py:245-255,
// both taken from the SORTED list of the mask voxels exactly as the reference does:
//   mean = np.mean(sorted(N4[mask>0]))   numpy's float32 PAIRWISE sum of the sorted values (leaves
//          of <= 128 elements with 8 strided accumulators, split at n/2 rounded down to a multiple
//          of 8), divided by M in float64 and rounded to float32 -- bit for bit;
//   p99  = sorted(...)[int(M*.99)].
// The kernel therefore sorts the mask voxels (order-preserving key of the float bits):
//   1. ONE pass over HBM: the mask is read once, the image only where the mask is set (16-byte
//      loads, software pipelined); the keys are appended to array A (warp-aggregated append) --
//      shared memory, overflowing into a per-CTA scratch in global memory when the mask does not
//      fit -- key range and count in registers;
//   2. bucket sort: histogram over two levels of range-normalised buckets in shared memory (about
//      8-16 keys per bucket whatever the distribution), exclusive scan,
//      scatter A -> B (B: per-CTA scratch in global memory, L2 resident) and a vectorised copy of
//      the bucketed keys back into A (a global-memory read per comparison would expose the L2
//      latency in step 3: measured 4x slower);
//   3. rank inside the bucket by counting (ties by position), sorted keys to B, copied back to A;
//   4. numpy's pairwise summation tree evaluated in parallel on the sorted array (8 lanes per leaf,
//      one per accumulator; the leaf sums combined bottom-up along numpy's recursion tree).
// Also zeroes the defect / LB counters.
#define VB_VDP_LOG_COARSE 11        /* level-1 buckets */
#define VB_VDP_LOAD 8               /* keys per level-2 bucket aimed at (8..16 after rounding) */
// level-2 bucket of a key (d = key - lowest key): cinfo[c] = first fine bucket | log2(parts) << 24
VB_DEV int vdp_fine_bucket(const unsigned* cinfo, unsigned d, int sh1, unsigned m1) {
    const unsigned ci = cinfo[d >> sh1];
    return (int)((ci & 0xFFFFFFu) + ((d & m1) >> (sh1 - (int)(ci >> 24))));
}

// block-wide exclusive prefix sum of h[0..nb) in place (each thread owns a contiguous run)
VB_DEV void vdp_block_exscan(unsigned* h, int nb, unsigned* wsum) {
    const int nt = (int)blockDim.x, tid = (int)threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int per = (nb + nt - 1) / nt;
    const int lo = min(nb, tid * per), hi = min(nb, lo + per);
    unsigned s = 0;
    for (int b = lo; b < hi; ++b) s += h[b];
    unsigned inc = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned t = __shfl_up_sync(VB_FULL, inc, d);
        if (lane >= d) inc += t;
    }
    __syncthreads();                                   // wsum may still be read from a previous call
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    unsigned run = inc - s;
    for (int i = 0; i < w; ++i) run += wsum[i];
    for (int b = lo; b < hi; ++b) { const unsigned t = h[b]; h[b] = run; run += t; }
    __syncthreads();
}

// depth of numpy's pairwise recursion for n elements (the right child is the larger one)
VB_DEV int vdp_pw_depth(int n) {
    int d = 0;
    while (n > 128) { int n2 = n >> 1; n2 -= n2 & 7; n -= n2; ++d; }
    return d;
}

// np.add.reduce of the float32 array a[0..M) given as sorted keys (numpy/_core/src/umath/
// loops_utils.h.src, *_pairwise_sum: PW_BLOCKSIZE 128, unroll 8), evaluated by the whole block.
// hv / hk: heap of 2^(D+1) partial sums / "has a value" flags, D = vdp_pw_depth(M).
VB_DEV float vdp_pairwise_sum(const unsigned* keys, int M, int D, float* hv, unsigned char* hk) {
    const int nt = (int)blockDim.x, tid = (int)threadIdx.x;
    const int nslots = 1 << D;
    for (int h = tid; h < 2 * nslots; h += nt) hk[h] = 0;
    __syncthreads();
    const int g = tid >> 3, j = tid & 7, ng = nt >> 3;
    for (int base = 0; base < nslots; base += ng) {            // trip count is block uniform
        const int slot = base + g;
        int start = 0, len = 0, h = 1, d = 0;
        bool owner = false;
        if (slot < nslots) {
            len = M;
            while (len > 128) {                                // walk numpy's recursion along the bits of `slot`
                int n2 = len >> 1;
                n2 -= n2 & 7;
                if ((slot >> (D - 1 - d)) & 1) { start += n2; len -= n2; h = 2 * h + 1; }
                else { len = n2; h = 2 * h; }
                ++d;
            }
            owner = (slot & ((1 << (D - d)) - 1)) == 0;        // one slot per leaf
        }
        float r = 0.0f;
        const int n8 = len - (len & 7);
        if (owner && len >= 8) {
            r = vb_key_float(keys[start + j]);
            for (int i = 8; i < n8; i += 8) r = __fadd_rn(r, vb_key_float(keys[start + i + j]));
        }
        // ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7))
        r = __fadd_rn(r, __shfl_xor_sync(VB_FULL, r, 1));
        r = __fadd_rn(r, __shfl_xor_sync(VB_FULL, r, 2));
        r = __fadd_rn(r, __shfl_xor_sync(VB_FULL, r, 4));
        if (owner && j == 0) {
            if (len >= 8) {
                for (int i = n8; i < len; ++i) r = __fadd_rn(r, vb_key_float(keys[start + i]));
            } else {
                r = 0.0f;
                for (int i = 0; i < len; ++i) r = __fadd_rn(r, vb_key_float(keys[start + i]));
            }
            hv[h] = r;
            hk[h] = 1;
        }
    }
    __syncthreads();
    for (int d = D - 1; d >= 0; --d) {
        for (int h = (1 << d) + tid; h < (2 << d); h += nt)
            if (!hk[h] && hk[2 * h] && hk[2 * h + 1]) { hv[h] = __fadd_rn(hv[2 * h], hv[2 * h + 1]); hk[h] = 1; }
        __syncthreads();
    }
    return hv[1];
}

// global -> shared copy of n keys by the whole block, 16 bytes per load, four loads in flight;
// ends with a block barrier
VB_DEV void vdp_copy_keys(unsigned* dst, const unsigned* src, int n) {
    const int nt = (int)blockDim.x, tid = (int)threadIdx.x, n4 = n >> 2;
    for (int i = tid; i < n4; i += 4 * nt) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (i + u * nt < n4) v[u] = *(const uint4*)(src + 4 * (size_t)(i + u * nt));
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (i + u * nt < n4) *(uint4*)(dst + 4 * (size_t)(i + u * nt)) = v[u];
    }
    for (int i = 4 * n4 + tid; i < n; i += nt) dst[i] = src[i];
    __syncthreads();
}

#define VB_VDP_HEAP_DEPTH 10        /* pairwise heaps up to this depth live in shared memory */
VB_KERNEL void VB_LAUNCH_BOUNDS(VB_VDP_THREADS, 1)
vdp_stats_kernel(const float* n4, const uint8_t* VB_RESTRICT mask, const VbSubject* subj, int n,
                 vb_subject_stats* stats, int key_cap, int dyn_words, unsigned* ws_a, unsigned* ws_b, long long ws_stride) {
    VB_DYN_SMEM(smem);
    // dynamic shared memory: [2^11 coarse buckets][2^11 + key_cap / VB_VDP_LOAD fine buckets][key_cap keys (array A)]
    unsigned* cinfo = (unsigned*)smem;
    VB_SHARED double red[64];
    VB_SHARED unsigned wsum[32];
    VB_SHARED float pw_val[2 << VB_VDP_HEAP_DEPTH];
    VB_SHARED unsigned char pw_kind[2 << VB_VDP_HEAP_DEPTH];
    VB_SHARED int s_nkeys;
    const int nt = (int)blockDim.x, tid = (int)threadIdx.x, lane = tid & 31;
    unsigned* As = cinfo + (2 << VB_VDP_LOG_COARSE) + key_cap / VB_VDP_LOAD;
    unsigned* Ag = ws_a + (size_t)blockIdx.x * (size_t)ws_stride;
    unsigned* B = ws_b + (size_t)blockIdx.x * (size_t)ws_stride;
    for (int si = (int)blockIdx.x; si < n; si += (int)gridDim.x) {
        const VbSubject S = subj[si];
        const float* x = n4 + S.off;
        const uint8_t* msk = mask + S.off;
        const bool aligned = (S.off & 3) == 0;
        __syncthreads();
        if (tid == 0) s_nkeys = 0;
        __syncthreads();
        int cnt = 0;
        unsigned kmin = 0xffffffffu, kmax = 0u;
        // software pipelined: the mask words of the next step are requested together with the
        // image loads of the current one (one DRAM latency per step instead of two)
        uint32_t mw[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) mw[u] = vb_load4b_word(msk, 4 * (u * nt + tid), S.V, aligned);
        for (int base = 0; base < S.V; base += 16 * nt) {            // trip count is block uniform
            Vb4f f[4];
            uint32_t mn[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (mw[u]) f[u] = vb_load4f(x, base + 4 * (u * nt + tid), S.V, aligned);
#pragma unroll
            for (int u = 0; u < 4; ++u) mn[u] = vb_load4b_word(msk, base + 16 * nt + 4 * (u * nt + tid), S.V, aligned);
            int c = 0;
#pragma unroll
            for (int u = 0; u < 4; ++u) c += __popc(vb_nonzero_bytes(mw[u]));
            if (__any_sync(VB_FULL, c)) {
                // warp-aggregated append: exclusive prefix of the lanes' counts, one atomic per warp
                int inc = c;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int t = __shfl_up_sync(VB_FULL, inc, d);
                    if (lane >= d) inc += t;
                }
                int wbase = 0;
                if (lane == 31) wbase = atomicAdd(&s_nkeys, inc);
                wbase = __shfl_sync(VB_FULL, wbase, 31);
                int pos = wbase + inc - c;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (!mw[u]) continue;
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if ((mw[u] >> (8 * e)) & 0xFFu) {
                            const unsigned key = vb_float_key(f[u].v[e]);
                            ++cnt;
                            kmin = min(kmin, key); kmax = max(kmax, key);
                            if (pos < key_cap) As[pos] = key; else Ag[pos] = key;
                            ++pos;
                        }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) mw[u] = mn[u];
        }
        const int M = (int)vb_block_reduce((long long)cnt, VbAddLL(), 0LL, (long long*)red);
        // (signed reductions of key ^ 0x80000000 keep the unsigned order)
        const unsigned lo0 = (unsigned)vb_block_reduce((int)(kmin ^ 0x80000000u), VbMinI(), 0x7fffffff, (int*)red) ^ 0x80000000u;
        const unsigned hi0 = (unsigned)vb_block_reduce((int)(kmax ^ 0x80000000u), VbMaxI(), (int)0x80000000, (int*)red) ^ 0x80000000u;
        float p99 = 0.0f, mean = 0.0f;
        if (M > 0) {
            const unsigned range = hi0 - lo0;
            const bool in_smem = M <= key_cap;
            unsigned* A = As;
            if (!in_smem) {                                          // the mask does not fit: all keys in global memory
                for (int i = tid; i < key_cap; i += nt) Ag[i] = As[i];
                A = Ag;
                __syncthreads();
            }
            if (range != 0u) {                                       // (all keys equal: A is sorted as it is)
                // Two-level buckets.  Float keys are logarithmically spaced, so one zero or tiny voxel
                // in the mask stretches [lo0, hi0] over dozens of binades and equal-width buckets
                // would put the bulk of the keys into a handful of them (measured: 600 keys per
                // bucket, 9x the instructions).  Level 1: <= 2^11 coarse buckets over the range;
                // level 2: a coarse bucket with c keys is cut into pow2ceil(c / load) equal parts.
                unsigned* fh = cinfo + (1 << VB_VDP_LOG_COARSE);     // fine buckets
                const int fine_cap = (in_smem ? (1 << VB_VDP_LOG_COARSE) + key_cap / VB_VDP_LOAD : dyn_words - (1 << VB_VDP_LOG_COARSE));
                const int sh1 = max(0, (32 - __clz((int)range)) - VB_VDP_LOG_COARSE);
                const int nbc = (int)(range >> sh1) + 1;             // <= 2^11
                const unsigned m1 = (1u << sh1) - 1u;
                for (int i = tid; i < nbc; i += nt) cinfo[i] = 0;
                __syncthreads();
                for (int i = tid; i < M; i += nt) atomicAdd(&cinfo[(A[i] - lo0) >> sh1], 1u);
                __syncthreads();
                int lt = 0;                                          // log2 of the load aimed at per fine bucket
                while ((1 << lt) < VB_VDP_LOAD) ++lt;
                while (((M >> lt) << 1) + nbc > fine_cap) ++lt;
                for (int c = tid; c < nbc; c += nt) {
                    const unsigned want = cinfo[c] >> lt;            // parts = pow2ceil(want), at most 2^sh1
                    int lg = (want > 1u) ? 32 - __clz((int)(want - 1u)) : 0;
                    lg = min(lg, sh1);
                    fh[c] = 1u << lg;
                    cinfo[c] = (unsigned)lg << 24;
                }
                __syncthreads();
                if (tid == 0) s_nkeys = (int)fh[nbc - 1];
                vdp_block_exscan(fh, nbc, wsum);                     // first fine bucket of every coarse one
                const int nf = (int)fh[nbc - 1] + s_nkeys;           // fine buckets in all
                for (int c = tid; c < nbc; c += nt) cinfo[c] |= fh[c];
                __syncthreads();
                for (int i = tid; i < nf; i += nt) fh[i] = 0;
                __syncthreads();
                for (int i = tid; i < M; i += nt) atomicAdd(&fh[vdp_fine_bucket(cinfo, A[i] - lo0, sh1, m1)], 1u);
                __syncthreads();
                vdp_block_exscan(fh, nf, wsum);                      // bucket starts
                for (int i = tid; i < M; i += nt) {
                    const unsigned key = A[i];
                    B[atomicAdd(&fh[vdp_fine_bucket(cinfo, key - lo0, sh1, m1)], 1u)] = key;   // afterwards fh[b] = end of bucket b
                }
                __syncthreads();
                if (in_smem) vdp_copy_keys(A, B, M);                 // bucketed keys next to the SM
                const unsigned* src = in_smem ? A : B;
                unsigned* dst = in_smem ? B : A;
                for (int p = tid; p < M; p += nt) {
                    const unsigned key = src[p];
                    const unsigned d = key - lo0;
                    const unsigned ci = cinfo[d >> sh1];
                    const int b = (int)((ci & 0xFFFFFFu) + ((d & m1) >> (sh1 - (int)(ci >> 24))));
                    const int o = b ? (int)fh[b - 1] : 0, e = (int)fh[b];
                    int rank = p - o;                                // a bucket one key wide: all its keys are equal
                    if (sh1 != (int)(ci >> 24)) {
                        rank = 0;
                        for (int q = o; q < e; ++q) {
                            const unsigned y = src[q];
                            rank += (int)((y < key) || (y == key && q < p));
                        }
                    }
                    dst[o + rank] = key;
                }
                __syncthreads();
                if (in_smem) vdp_copy_keys(A, B, M);                 // sorted keys back
            }
            long long k = (long long)((double)M * .99);              // int(len(signal_list)*.99)
            if (k >= M) k = M - 1;
            p99 = vb_key_float(A[k]);
            const int D = vdp_pw_depth(M);
            float total;
            if (D <= VB_VDP_HEAP_DEPTH) total = vdp_pairwise_sum(A, M, D, pw_val, pw_kind);
            else total = vdp_pairwise_sum(A, M, D, (float*)B, (unsigned char*)(B + ((size_t)2 << D)));
            mean = (float)((double)total / (double)M);               // np.mean: float32 sum / intp count in float64
        }
        if (tid == 0) {
            vb_subject_stats* st = stats + si;
            st->n_mask = M;
            st->mean = mean;
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
// The float32 quotient tests of Vent_Analysis.py:246-256 as comparisons with per-subject thresholds:
// for d > 0 the rounded quotient RN(x / d) is non-decreasing in x, so {x : RN(x / d) <= c} is
// {x <= T} with T found by a few exact divisions next to c * d (same for <).  NaN fails every
// comparison, as in numpy.  Degenerate divisors (<= 0, non-finite, tiny) keep IEEE divisions.
VB_DEV float vb_next_up_pos(float x) { return __uint_as_float(__float_as_uint(x) + 1u); }
VB_DEV float vb_next_down_pos(float x) { return __uint_as_float(__float_as_uint(x) - 1u); }
// largest x with RN(x / d) <= c (LE) or < c (!LE), for normal c, d > 0; NaN if not found
template <bool LE>
VB_DEV float vdp_quotient_threshold(float c, float d) {
    float x = __fmul_rn(c, d);
    int guard = 0;
    for (; guard < 64; ++guard) {                        // first x that fails
        const float q = __fdiv_rn(x, d);
        if (!(LE ? (q <= c) : (q < c))) break;
        x = vb_next_up_pos(x);
    }
    for (; guard < 128; ++guard) {                       // last x that passes
        const float q = __fdiv_rn(x, d);
        if (LE ? (q <= c) : (q < c)) return x;
        x = vb_next_down_pos(x);
    }
    return __uint_as_float(0x7fc00000u);
}
struct VdpThr { float d, t1, t2, t3, t4, t5; bool fast; };   // x <= d: below thresh; x <= t_i: LB code <= i
// threads 0..5 find the thresholds (s_thr: 8 floats of shared memory); all threads must call
VB_DEV VdpThr vdp_thresholds(float* s_thr, float thresh, float mean, float p99) {
    const int tid = (int)threadIdx.x;
    const bool okdiv = mean >= 1e-30f && mean <= 1e30f && p99 >= 1e-30f && p99 <= 1e30f && thresh >= 1e-3f && thresh <= 1e3f;
    if (okdiv && tid < 6) {
        const float cs[5] = {0.16f, 0.34f, 0.52f, 0.7f, 0.88f};
        s_thr[tid] = (tid == 0) ? vdp_quotient_threshold<false>(thresh, mean) : vdp_quotient_threshold<true>(cs[tid - 1], p99);
    }
    __syncthreads();
    VdpThr T;
    T.d = s_thr[0]; T.t1 = s_thr[1]; T.t2 = s_thr[2]; T.t3 = s_thr[3]; T.t4 = s_thr[4]; T.t5 = s_thr[5];
    T.fast = okdiv && T.d == T.d && T.t1 == T.t1 && T.t2 == T.t2 && T.t3 == T.t3 && T.t4 == T.t4 && T.t5 == T.t5;
    return T;
}
VB_DEV void vdp_classify_one(const VdpThr& T, float xe, float thresh, float mean, float p99, bool& below, int& cd) {
    if (T.fast) {
        below = xe <= T.d;
        cd = 6 - (int)(xe <= T.t1) - (int)(xe <= T.t2) - (int)(xe <= T.t3) - (int)(xe <= T.t4) - (int)(xe <= T.t5);
        cd = (xe == xe) ? cd : 0;                          // NaN compares false everywhere -> 0, as in numpy
    } else {
        below = __fdiv_rn(xe, mean) < thresh;
        const float b = __fdiv_rn(xe, p99);
        cd = (b <= 0.16f) ? 1 : (b <= 0.34f) ? 2 : (b <= 0.52f) ? 3 : (b <= 0.7f) ? 4 : (b <= 0.88f) ? 5
           : (b > 0.88f) ? 6 : 0;
    }
}
#ifndef VB_CLS_THREADS
#define VB_CLS_THREADS 512
#endif
// byte-wise version: any shape / alignment
VB_DEV void vdp_classify_bytes(unsigned char* smem, int* red, float* s_thr, const float* VB_RESTRICT n4, const uint8_t* VB_RESTRICT mask,
                               const VbSubject& S, int si, float thresh, int R, uint8_t* lb, uint8_t* defect,
                               uint8_t* border, vb_subject_stats* stats) {
    const int n0 = S.n0, n1 = S.n1, n2 = S.n2, ncols = n1 * n2;
    const int i0 = (int)blockIdx.x * R;
    const int nt = (int)blockDim.x, tid = (int)threadIdx.x;
    const int ra = max(i0 - 2, 0), rb = min(i0 + R + 2, n0);          // rows of d0 held
    const int ma = max(i0 - 1, 0), mb = min(i0 + R + 1, n0);          // rows of the median held
    uint8_t* d0 = smem;                                               // [(R + 4) * ncols]
    uint8_t* md = smem + (size_t)(R + 4) * ncols;                     // [(R + 2) * ncols]
    const float mean = stats[si].mean, p99 = stats[si].p99;
    // the same exact per-subject thresholds as the word version (identical results by construction)
    const VdpThr T = vdp_thresholds(s_thr, thresh, mean, p99);
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
                    bool below;
                    int cd;
                    vdp_classify_one(T, f.v[e], thresh, mean, p99, below, cd);
                    d0[v + e - v0] = (uint8_t)(below && in);
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

// ---------------------------------------------------------------------------------------------
// Four-voxels-per-instruction version of the same tile (volumes whose row length is a multiple
// of 4 at a 4-voxel aligned offset).
//  * The float32 quotient tests are replaced by comparisons with per-subject thresholds: for
//    d > 0 the rounded quotient RN(x / d) is non-decreasing in x, so {x : RN(x / d) <= c} is
//    {x <= T} with T found by a few exact divisions next to c * d (same for <).  NaN fails every
//    comparison, as in numpy.  Degenerate divisors (<= 0, non-finite, tiny) keep IEEE divisions.
//  * d0 / median tiles are bytes handled as 32-bit words (four (col, slice) positions at once):
//    the 3x3 count is three word additions, "count >= 5" one add/shift/and, the border two XORs.
//    Column neighbours are n2 bytes away: two aligned words and a funnel shift; the per-byte
//    "has a left / right neighbour" masks are constants of the thread.
VB_DEV uint32_t vb_funnel_r(uint32_t lo, uint32_t hi, int shift) {
#if defined(VB_EMULATE)
    return (uint32_t)((((uint64_t)hi << 32) | (uint64_t)lo) >> (shift & 31));
#else
    return __funnelshift_r(lo, hi, shift);
#endif
}
struct VdpNbr { int wl, sl, wr, sr; uint32_t maskL, maskR; };
VB_DEV VdpNbr vdp_nbr(int g, int n1, int n2) {
    VdpNbr k;
    k.maskL = k.maskR = 0u;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int j = (4 * g + e) / n2;
        if (j > 0) k.maskL |= 0xFFu << (8 * e);
        if (j < n1 - 1) k.maskR |= 0xFFu << (8 * e);
    }
    const int aL = 4 * g - n2, aR = 4 * g + n2;
    k.wl = aL >> 2; k.sl = (aL & 3) * 8;                 // arithmetic shift: floor for negative aL
    k.wr = aR >> 2; k.sr = (aR & 3) * 8;
    return k;
}

VB_DEV void vdp_classify_words(unsigned char* smem, int* red, float* s_thr, int P, const float* VB_RESTRICT n4,
                               const uint8_t* VB_RESTRICT mask, const VbSubject& S, int si, float thresh, int R,
                               uint8_t* lb, uint8_t* defect, uint8_t* border, vb_subject_stats* stats) {
    const int n0 = S.n0, n1 = S.n1, n2 = S.n2, ncols = n1 * n2, G = ncols >> 2;
    const int i0 = (int)blockIdx.x * R;
    const int nt = (int)blockDim.x, tid = (int)threadIdx.x;
    const int ra = max(i0 - 2, 0), rb = min(i0 + R + 2, n0);          // rows of d0 held
    const int ma = max(i0 - 1, 0), mb = min(i0 + R + 1, n0);          // rows of the median held
    const int ib = min(i0 + R, n0);
    uint32_t* d0w = (uint32_t*)smem + P;                              // [(R + 4) * G], P words of padding around
    uint32_t* mdw = d0w + (size_t)(R + 4) * G + 2 * P;                // [(R + 2) * G]
    const float mean = stats[si].mean, p99 = stats[si].p99;
    const VdpThr T = vdp_thresholds(s_thr, thresh, mean, p99);   // thresholds, else exact divisions
    const float* x = n4 + S.off + (size_t)ra * ncols;
    const uint8_t* msk = mask + S.off + (size_t)ra * ncols;
    uint8_t* lbo = lb + S.off + (size_t)ra * ncols;
    const int ng = (rb - ra) * G;
    const int own0 = (i0 - ra) * G, own1 = (ib - ra) * G;             // groups this CTA owns
    int c12 = 0;
    uint32_t any_d0 = 0u;
    for (int gb = tid; gb < ng; gb += 4 * nt) {                       // four independent 20-byte loads in flight
        // d0 and the code are zero outside the mask: the image is only read (and classified)
        // where the mask word is set
        float4 xv[4];
        uint32_t mw[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int gi = gb + u * nt;
            mw[u] = (gi < ng) ? *(const uint32_t*)(msk + 4 * (size_t)gi) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int gi = gb + u * nt;
            if (mw[u]) xv[u] = *(const float4*)(x + 4 * (size_t)gi);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int gi = gb + u * nt;
            if (gi >= ng) break;
            uint32_t dw = 0u, cw = 0u, w12 = 0u;
            if (mw[u]) {
            const float xs[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float xe = xs[e];
                const bool in = ((mw[u] >> (8 * e)) & 0xFFu) != 0u;
                int cd;
                bool below;
                vdp_classify_one(T, xe, thresh, mean, p99, below, cd);
                cd = in ? cd : 0;
                dw |= (uint32_t)(below && in) << (8 * e);
                cw |= (uint32_t)cd << (8 * e);
                w12 |= (uint32_t)(cd == 1 || cd == 2) << e;
            }
            }
            d0w[gi] = dw;
            any_d0 |= dw;
            if (gi >= own0 && gi < own1) {
                *(uint32_t*)(lbo + 4 * (size_t)gi) = cw;
                c12 += __popc(w12);
            }
        }
    }
    int nd = 0;
    if (!__syncthreads_or(any_d0 != 0u)) {
        // nothing below the threshold in rows [ra, rb): median and border of the owned rows are zero
        for (int g = tid; g < G; g += nt)
            for (int i = i0; i < ib; ++i) {
                *(uint32_t*)(defect + S.off + (size_t)i * ncols + 4 * g) = 0u;
                if (border) *(uint32_t*)(border + S.off + (size_t)i * ncols + 4 * g) = 0u;
            }
        c12 = vb_block_reduce(c12, VbAddI(), 0, red);
        if (tid == 0 && c12) atomicAdd(&stats[si].n_lb12, c12);
        return;
    }
    // median of rows [ma, mb): a thread owns four (col, slice) positions and walks down the rows
    // with a sliding window of the three-column sums
    for (int g = tid; g < G; g += nt) {
        const VdpNbr k = vdp_nbr(g, n1, n2);
        uint32_t s_prev = 0u, s_cur;
        if (ma > 0) {                                                 // row ma - 1 is held whenever it exists
            const uint32_t* p = d0w + (size_t)(ma - 1 - ra) * G;
            s_prev = p[g] + (vb_funnel_r(p[k.wl], p[k.wl + 1], k.sl) & k.maskL) + (vb_funnel_r(p[k.wr], p[k.wr + 1], k.sr) & k.maskR);
        }
        {
            const uint32_t* p = d0w + (size_t)(ma - ra) * G;
            s_cur = p[g] + (vb_funnel_r(p[k.wl], p[k.wl + 1], k.sl) & k.maskL) + (vb_funnel_r(p[k.wr], p[k.wr + 1], k.sr) & k.maskR);
        }
        for (int i = ma; i < mb; ++i) {
            uint32_t s_next = 0u;
            if (i + 1 < n0) {                                         // row i + 1 <= mb is held whenever it exists
                const uint32_t* p = d0w + (size_t)(i + 1 - ra) * G;
                s_next = p[g] + (vb_funnel_r(p[k.wl], p[k.wl + 1], k.sl) & k.maskL) + (vb_funnel_r(p[k.wr], p[k.wr + 1], k.sr) & k.maskR);
            }
            // per byte: count (<= 9) >= 5  <=>  bit 7 of count + 123
            const uint32_t d = ((s_prev + s_cur + s_next + 0x7B7B7B7Bu) >> 7) & 0x01010101u;
            mdw[(size_t)(i - ma) * G + g] = d;
            if (i >= i0 && i < ib) {
                *(uint32_t*)(defect + S.off + (size_t)i * ncols + 4 * g) = d;
                nd += __popc(d);
            }
            s_prev = s_cur; s_cur = s_next;
        }
    }
    __syncthreads();
    if (border) {
        for (int g = tid; g < G; g += nt) {
            const VdpNbr k = vdp_nbr(g, n1, n2);
            for (int i = i0; i < ib; ++i) {
                const uint32_t* a = mdw + (size_t)(i - ma) * G;
                const uint32_t c = a[g];
                const uint32_t up = (i == 0) ? c : a[g - G], dn = (i == n0 - 1) ? c : a[g + G];
                const uint32_t dl = (vb_funnel_r(a[k.wl], a[k.wl + 1], k.sl) & k.maskL) | (c & ~k.maskL);
                const uint32_t dr = (vb_funnel_r(a[k.wr], a[k.wr + 1], k.sr) & k.maskR) | (c & ~k.maskR);
                *(uint32_t*)(border + S.off + (size_t)i * ncols + 4 * g) = (dn ^ up) | (dr ^ dl);
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

// grid = (row tiles, subjects); dynamic shared memory: (2R + 6) * max row bytes + 16 * P.
// A subject is handled by the word kernel when its rows are 4-voxel aligned, else by the byte
// kernel (the host launches whichever the batch needs; each kernel skips the other's subjects).
VB_DEV bool vdp_word_eligible(const VbSubject& S, int P) {
    return ((S.n1 * S.n2) & 3) == 0 && (S.off & 3) == 0 && (S.n2 + 3) / 4 + 1 <= P;
}
VB_KERNEL void VB_LAUNCH_BOUNDS(VB_CLS_THREADS, 2)
vdp_classify4_kernel(const float* VB_RESTRICT n4, const uint8_t* VB_RESTRICT mask, const VbSubject* subj, float thresh,
                     int R, int P, uint8_t* lb, uint8_t* defect, uint8_t* border, vb_subject_stats* stats) {
    VB_DYN_SMEM(smem);
    VB_SHARED int red[64];
    VB_SHARED float s_thr[8];
    const int si = (int)blockIdx.y;
    const VbSubject S = subj[si];
    if ((int)blockIdx.x * R >= S.n0 || !vdp_word_eligible(S, P)) return;
    vdp_classify_words(smem, red, s_thr, P, n4, mask, S, si, thresh, R, lb, defect, border, stats);
}
VB_KERNEL void VB_LAUNCH_BOUNDS(VB_CLS_THREADS, 2)
vdp_classify_kernel(const float* VB_RESTRICT n4, const uint8_t* VB_RESTRICT mask, const VbSubject* subj, float thresh,
                    int R, int P, uint8_t* lb, uint8_t* defect, uint8_t* border, vb_subject_stats* stats) {
    VB_DYN_SMEM(smem);
    VB_SHARED int red[64];
    VB_SHARED float s_thr[8];
    const int si = (int)blockIdx.y;
    const VbSubject S = subj[si];
    if ((int)blockIdx.x * R >= S.n0 || vdp_word_eligible(S, P)) return;
    vdp_classify_bytes(smem, red, s_thr, n4, mask, S, si, thresh, R, lb, defect, border, stats);
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
