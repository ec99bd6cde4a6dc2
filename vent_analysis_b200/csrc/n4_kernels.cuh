// N4 bias-field correction, one persistent CTA per subject.
//
// Replaces Vent_Analysis.N4_bias_correction (Vent_Analysis.py:316-334), i.e.
// sitk.N4BiasFieldCorrectionImageFilter() with all defaults (ITK 5.3
// itkN4BiasFieldCorrectionImageFilter.hxx; restated on the CPU in oracle/n4_oracle.c).
//
// Design (DESIGN.md §4 "N4"):
//   * A CTA (VB_N4_THREADS = 896 threads, 72 registers) pulls subjects from a global work counter and
//     runs the WHOLE filter for one subject -- 4 levels x <=50 iterations x {histogram, Wiener
//     sharpening, B-spline fit, update, convergence} -- with block / named barriers only; no host
//     round trip, subjects converge independently.
//   * Only mask voxels take part in the iterations, so the volume is COMPACTED once per subject into
//     arrays of length M in C ORDER (row-major over the mask voxels: the order of the reference's
//     convergence loop): the log image and the convergence sequence (global memory, L2 resident),
//     U = logI - bias of the committed lattice -- for the benchmark volumes (128x128x14, ~36k mask
//     voxels = 143 KB) in SHARED MEMORY for the whole filter, larger masks in global memory, same
//     code -- and a per-voxel code (row << 16 | active-column index).  The columns (col, slice) that
//     carry mask voxels are grouped in blocks of 32 (one warp); the voxels of one row inside a block
//     are contiguous in the compact arrays (rank base rbb[row][block]), so a warp that walks the
//     rows of its block reads one conflict-free stream.  Walk items (block x row segment) are drawn
//     largest-first from a shared counter.
//   * Only the update pass evaluates the bias (old and new field) and writes the new U; histogram
//     and fit read U.  It is a FLAT loop over the compact arrays: row and column from the per-voxel
//     code, the collapsed coefficients of both fields from per-column tables G[a][column] (rebuilt
//     from the control lattice after every fit) -- no row walk, every lane busy.  (Tables that do
//     not fit, or volumes too large for the code packing: the column walk with a sliding window.)
//   * Parzen histogram: a flat loop over the compact U array, fixed-point (count-conserving)
//     shared-memory reductions into one copy per lane => exactly associative, deterministic.
//   * Sharpening: 4 packed 512-point complex FFTs in double in shared memory, two radix-2 stages
//     per pass, twiddles computed once per CTA.
//   * B-spline fit: the Lee-Wolberg-Shin scatter is separable (t = tr*tc*ts,
//     sum t^2 = Sr*Sc*Ss), so delta = sum w^3 z / sum w^2 is three small contractions:
//     rows (registers, sliding 4-tap window along the column walk) -> columns (64-bit fixed-point
//     shared-memory adds built from two native 32-bit atomics with carry forwarding) -> slices.
//     omega depends only on the mask and the lattice size: once per level.
//   * Convergence: ITK's CalculateConvergenceMeasurement is a SEQUENTIAL float32 Welford
//     recurrence whose rounding (the running mean stalls once its update drops below half an
//     ulp) decides when a level stops.  It is reproduced here in parallel: threads run their
//     consecutive steps exactly from guessed start states, the float bit-pattern increments are
//     prefix-summed, and this repeats until no increment changes -- the trajectory then IS the
//     sequential one.  The first chunk predicts its starts from the exact-arithmetic running
//     mean.  The chain is skipped when a rigorous bound already decides "continue"; otherwise it
//     runs on a team of warps while the other warps compute the next iteration's histogram /
//     sharpening / fit speculatively (committed only if the chain says "continue").
//   * Shared-memory tables, U and the atomics of the hot loops go through 32-bit SHARED addresses
//     (vb_saddr: LDS / STS / ATOMS / RED.shared).  Pointers into global memory are opaque local
//     copies: at the register cap the compiler otherwise re-derives them from the kernel parameters
//     at each use -- but an opaque pointer into shared memory is a GENERIC pointer (64-bit address
//     arithmetic, generic LD / ATOM), which is what the shared addresses avoid.
#pragma once
#include <float.h>
#include "vb_common.cuh"

#ifndef VB_N4_THREADS
/* 28 warps x 72 registers measured 1.1 % faster than 32 x 64 (and than 24 x 80, 30 x 64, 26 x 72):
   the kernel is bound by instruction issue, a few more registers per thread trim the spills */
#define VB_N4_THREADS 896
#endif
#define VB_N4_WARPS (VB_N4_THREADS / 32)
#define VB_BSPLINE_EPS 1e-3f
#ifndef VB_N4_EPT
#define VB_N4_EPT 8                               /* convergence chain: elements per thread   */
#endif
#ifndef VB_N4_CHAIN_WARPS
#define VB_N4_CHAIN_WARPS 12                      /* warps of the chain team while it overlaps (10 / 14 of 28 measured slower) */
#endif
#define VB_N4_SERIAL 64                           /* first elements: one thread, sequentially */
#ifndef VB_N4_WARP_END
#define VB_N4_WARP_END 64                         /* ... then one warp up to here (multiple of 8; 64 = none), then the team */
#endif
#ifndef VB_N4_SIGMA_PREDICT_ALL
#define VB_N4_SIGMA_PREDICT_ALL 0                 /* 1: every chunk predicts its sigma starts */
#endif
#ifndef VB_N4_PREDICT_END
#define VB_N4_PREDICT_END 4096                    /* team chunks that start below this predict their starts (full length) */
#endif
#ifndef VB_N4_STREAM
#define VB_N4_STREAM 0                            /* per-iteration streams (log image, voxel codes, chain sequence) bypass L1: the small walk tables stay there */
#endif
#if VB_N4_STREAM
#define VB_N4_LDS(p) vb_ld_stream(p)
#define VB_N4_STS(p, v) vb_st_stream(p, v)
#else
#define VB_N4_LDS(p) (*(p))
#define VB_N4_STS(p, v) (*(p) = (v))
#endif
#ifndef VB_N4_HIST4
#define VB_N4_HIST4 0                              /* histogram: 16-byte LDS of four voxels (faster for one subject alone, not for a full batch) */
#endif
#ifndef VB_N4_UB
#define VB_N4_UB 6                                 /* flat update pass: voxels per trip (loads issued together) */
#endif
#ifndef VB_N4_MERGED_RED
#define VB_N4_MERGED_RED 1                        /* n4_update: its four block reductions behind one pair of barriers */
#endif
#ifndef VB_N4_G2
#define VB_N4_G2 1                                /* two coefficient tables (swap) instead of table + copy when both fit */
#endif
#define VB_N4_RED 128                             /* doubles of block-reduction scratch */
#define VB_N4_HCOPIES 32                          /* histogram copies: one per lane (stride nbins | 1) */

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
    float* ws_logI;          // [grid][Vmax]      compact log image (L2 resident; read once per iteration)
    float* ws_u;             // [grid][Vmax]      compact U = logI - bias when it does not fit smem
    uint32_t* ws_colmask;    // [grid][cmw_max]   per column: row bitmask words
    uint32_t* ws_rbb;        // [grid][rbb_max]   C-order rank of the first mask voxel of (row, block)
    uint32_t* ws_wbase;      // [grid][wb_max]    compact offset of every (block, 32-row word) (+1)
    uint32_t* ws_acol;       // [grid][ncols_max] active (mask-carrying) columns, ascending
    uint32_t* ws_order;      // [grid][wb_max]    work items, largest first
    float* ws_axis;          // [grid][axis_stride]  (4-byte words)
    float* ws_ec;            // [grid][Vmax]      convergence sequence exp(old - new) in C order
    uint32_t* ws_code;       // [grid][Vmax]      per mask voxel (C order): row << 16 | index of its column among the active ones
    int* work_counter;
    float* debug;            // may be null: [n][debug_stride]
    long long Vmax, cmw_max, rbb_max, wb_max, ncols_max;
    int axis_stride, debug_stride;
    int Lmax;                // max lattice entries (after the last refinement)
    int fftN;                // padded histogram size
    int union_bytes;         // size of the shared-memory union region
    int logI_cap;            // floats of compact log image that fit shared memory (0 = none)
    int rowtab_rows;         // rows whose row-axis tables fit shared memory (0 = keep them in global)
    // cluster mode (K > 1): a thread-block cluster of K CTAs works on ONE subject -- small batches, single
    // subjects (the reference's GUI use).  The compact arrays then live in global memory (L2) and are
    // shared by the CTAs; per-CTA partial results meet in the scratch below.
    int K;                   // CTAs per subject; 1: one persistent CTA per subject pulls subjects from the counter
    unsigned long long* cl_hist;   // [clusters][4][K][n_bins]    partial histograms (fixed point); per team (all threads / phase team) two sets used in turn
    unsigned* cl_q;          // [clusters][4][K][2 * q_max]  partial Q accumulators of the fit (low, high words); sets as cl_hist
    double* cl_red;          // [clusters][2][K][4]          partial (min, max, s1, s2) of the update / range passes
    unsigned* cl_bar;        // [clusters][4]             arrival counters of the cross-CTA barriers (zeroed before the launch)
    int* cl_bcast;           // [clusters][16]            scalars of the setup, written by CTA 0 of the cluster
    int q_max;
    int cl_chain_warps;      // warps of the chain team in cluster mode (the phase team has 1 / K of the work there); 0: VB_N4_CHAIN_WARPS
    int no_overlap;          // 1: run the convergence chain and the next iteration's phases one after the other
                             // (same results; A/B timing of the overlap)
};

// debug record layout (floats): [0]=bmin [1]=bmax [2..4]=ncp [5]=n_trace, then
// hist[nbins], emap[nbins], lattice[Lmax], trace[...]
#define VB_DBG_HDR 24   /* [8..15]: phase time in microseconds (setup, hist, sharpen, fit, update, chain, finalize, other) */

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

// Per-axis tables, 16 words per sample: fit3[4] | fit2[4] | evw[4] | bfit | bev | lo | hi
struct N4Axis {
    const float4* fit3;   // B^3 / sum(B^2)   (delta numerator weights)
    const float4* fit2;   // B^2              (omega weights)
    const float4* evw;    // B                (field evaluation)
    const int* bfit;      // first supporting control point, fit parameterisation
    const int* bev;       // first supporting control point, evaluation parameterisation
    const int* lo;        // [ncp] first sample whose fit support contains control point b
    const int* hi;        // [ncp] one past the last such sample
};
VB_DEV N4Axis n4_axis(float* base, int n) {
    N4Axis a;
    a.fit3 = (const float4*)base;
    a.fit2 = (const float4*)(base + 4 * (size_t)n);
    a.evw = (const float4*)(base + 8 * (size_t)n);
    a.bfit = (const int*)(base + 12 * (size_t)n);
    a.bev = (const int*)(base + 13 * (size_t)n);
    a.lo = (const int*)(base + 14 * (size_t)n);
    a.hi = (const int*)(base + 15 * (size_t)n);
    return a;
}

// BSplineScatteredDataPointSetToImageFilter / BSplineControlPointImageFilter parametrisation
// of one axis (oracle: fit_axis_table / eval_axis_table in n4_oracle.c).
VB_DEV void n4_build_axis(float* base, int slot, int n, int ncp) {
    float4* fit3 = (float4*)base;
    float4* fit2 = (float4*)(base + 4 * (size_t)slot);
    float4* evw = (float4*)(base + 8 * (size_t)slot);
    int* bfit = (int*)(base + 12 * (size_t)slot);
    int* bev = (int*)(base + 13 * (size_t)slot);
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

// support ranges of the control points along one axis (bfit is non-decreasing)
VB_DEV void n4_build_support(float* base, int slot, int n, int ncp) {
    const int* bfit = (const int*)(base + 12 * (size_t)slot);
    int* lo = (int*)(base + 14 * (size_t)slot);
    int* hi = (int*)(base + 15 * (size_t)slot);
    for (int b = (int)threadIdx.x; b < ncp; b += (int)blockDim.x) {
        int l = 0;
        while (l < n && bfit[l] < b - 3) ++l;
        int h = l;
        while (h < n && bfit[h] <= b) ++h;
        lo[b] = l; hi[b] = h;
    }
}

// CL = false (one CTA per subject): the cluster accessors below are the literals 1 / 0, so the cluster-mode code
// and state of every function fold away at instantiation.
template <bool CL>
struct N4CtxT {
    // subject
    int n0, n1, n2, ncols, W, nblk;   // nblk: blocks of 32 ACTIVE columns
    long long V;
    const float* img;
    const uint8_t* msk;
    float* out;
    float* logI;          // compact log image (global memory, L2)
    float* U;             // compact U = logI - bias(L) of the committed lattice (shared or global memory)
    bool u_smem;          // U lives in shared memory
    uint32_t* colmask;    // [W][ncols]
    uint32_t* rbb;        // [n0][nblk] (+1)
    uint32_t* wbase;      // [nblk][W] (+1)
    uint32_t* acol;       // [nact]
    uint32_t* iorder;     // [nblk * nseg] work items sorted by voxel count, largest first
    int* ctr;             // shared-memory work counters, one per team barrier id
    // cluster mode (see N4Args): K CTAs per subject, this CTA's rank; walk items j -> j * istep + ioff
    int K, crank, istep, ioff, q_max;
    VB_DEV_MEMBER int k() const { return CL ? K : 1; }
    VB_DEV_MEMBER int rank() const { return CL ? crank : 0; }
    VB_DEV_MEMBER int step() const { return CL ? istep : 1; }
    VB_DEV_MEMBER int off() const { return CL ? ioff : 0; }
    VB_DEV_MEMBER void deal(int s, int o) { if (CL) { istep = s; ioff = o; } }
    unsigned long long* cl_hist;
    unsigned* cl_q;
    double* cl_red;
    unsigned* cl_bar;
    int* cl_bcast;
    unsigned gen[3];      // cross-CTA barriers passed so far, per team barrier id
    // which of the two sets of exchange slots is used next.  Histogram and fit run on all threads or on the
    // phase team only: the threads of the chain team then miss exchanges, so each of the two teams has sets
    // (and a count) of its own -- [0]: all threads, [1]: the phase team
    int par_hist0, par_hist1, par_q0, par_q1, par_red;
    int nact, nseg, wps;  // active columns; row segments per block; 32-row words per segment
    int M;                // mask voxels (length of the compact arrays)
    int union_bytes;      // size of the shared-memory union region
    float* axis_base;
    float* ec;            // [M] exp(old - new) in C order (chain input)
    uint32_t* code;       // [M] row << 16 | active-column index of every mask voxel (null: volume too large for the packing)
    // row-axis tables (shared memory when they fit, else inside axis_base)
    const float4* r_evw;
    const float4* r_fit3;
    const int* r_bev;
    const int* r_bfit;
    float* rowtab;        // shared-memory home of the four arrays above (null: global)
    // shared memory
    float* un_f;          // union region
    double* un_d;
    float* L;             // lattice [ncp0][ncp1][ncp2]
    float* Lold;          // lattice before the last update
    float* Om;            // omega lattice of the current level
    float* Phi;           // lattice increment of the last fit (committed by n4_commit)
    float* G;             // [ncp0][gstride] column coefficients of c.L per active column (null: evaluate on the fly)
    float* G2;            // second table behind G when both fit: the update pass builds the new lattice's table there and swaps
    float* gspace;        // spare shared memory behind the compact log image
    int gcap, gstride;    // floats available there; row stride of G (active columns rounded up to 32)
    float* E;             // E(u|v) table
    float* H;             // histogram
    double* red;          // VB_N4_RED doubles reduction scratch
    double* tw;           // FFT twiddles cos | sin, fftN doubles (constant for the launch)
    int ncp[3];
    int slot[3];          // per-axis table length: max(samples, final control points)
    int nbins, fftN, hshift;
    long long t_chain;    // debug: clock ticks spent in n4_convergence
    int n_sweeps;         // debug: Jacobi sweeps of the convergence chain
    long long t_sub[4];   // debug: chain time in prefix / chunk setup / mu sweeps / sigma sweeps
};

template <class CTX> VB_DEV N4Axis n4_ax(const CTX& c, int a) {
    size_t off = 0;
    for (int i = 0; i < a; ++i) off += 16 * (size_t)c.slot[i];
    return n4_axis(c.axis_base + off, c.slot[a]);
}

// A team is a warp-aligned group of threads of the CTA with its own named barrier: the whole
// block (barrier 0), or -- while the convergence chain of iteration i overlaps the speculative
// histogram / sharpening / fit of iteration i+1 -- the "chain" team and the "phase" team.
struct N4Team {
    int tid, nt, bar;     // index inside the team, team size (multiple of 32), barrier id
};
VB_DEV void n4_sync(const N4Team& T) {
    if (T.bar == 0) __syncthreads(); else vb_bar_sync(T.bar, T.nt);
}
VB_DEV int n4_sync_or(const N4Team& T, int pred) {
    return (T.bar == 0) ? __syncthreads_or(pred) : vb_bar_or(T.bar, T.nt, pred);
}
// Barrier over team T of EVERY CTA of the cluster (cluster mode; K == 1: the team barrier).  The CTAs
// of a cluster are co-scheduled, so thread 0 of the team may spin on a counter in global memory:
// arrive (release), wait until all K teams have arrived for this generation (acquire).  Everything
// the teams wrote to global memory before the barrier is visible to all of them after it -- through
// L2: such data must be read with vb_ld_cg (a line may sit in this SM's L1 from an earlier pass).
template <class CTX> VB_DEV void n4_csync(CTX& c, const N4Team& T) {
    n4_sync(T);
    if (c.k() == 1) return;
    const unsigned target = (++c.gen[T.bar]) * (unsigned)c.k();
    if (T.tid == 0) {
        __threadfence();
        atomicAdd(c.cl_bar + T.bar, 1u);
        while (vb_ld_acquire(c.cl_bar + T.bar) < target) { }
        __threadfence();
    }
    n4_sync(T);
}

// Loads of arrays that other CTAs of the cluster write between barriers (U, the convergence sequence):
// through L2 in cluster mode; one CTA per subject: the plain load.
template <class V, class CTX> VB_DEV V n4_ld_x(const CTX& c, const V* p) { return (c.k() > 1) ? vb_ld_cg(p) : VB_N4_LDS(p); }

// ---------------------------------------------------------------------------------------------
// block-wide exclusive scan of one int per thread; every thread gets its prefix and the total
VB_DEV int n4_block_scan(int v, int& total, int* scratch /* >= 33 ints */) {
    const int lane = vb_lane(), w = vb_warp(), nw = (int)(blockDim.x >> 5);
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(VB_FULL, inc, d);
        if (lane >= d) inc += t;
    }
    __syncthreads();
    if (lane == 31) scratch[w] = inc;
    __syncthreads();
    int wt = (lane < nw) ? scratch[lane] : 0;            // nw <= 32
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(VB_FULL, wt, d);
        if (lane >= d) wt += t;
    }
    total = __shfl_sync(VB_FULL, wt, 31);
    const int woff = (w > 0) ? __shfl_sync(VB_FULL, wt, (w > 0) ? w - 1 : 0) : 0;
    return woff + inc - v;
}

// in-place exclusive scan of arr[0..n) (global), arr[n] = total; returns the total
VB_DEV int n4_array_scan(uint32_t* arr, int n, int* scratch) {
    const int nt = (int)blockDim.x, tid = (int)threadIdx.x;
    const int per = (n + nt - 1) / nt;
    const int lo = min(n, tid * per), hi = min(n, lo + per);
    int s = 0;
    for (int i = lo; i < hi; ++i) s += (int)arr[i];
    int total;
    int run = n4_block_scan(s, total, scratch);
    for (int i = lo; i < hi; ++i) {
        const int t = (int)arr[i];
        arr[i] = (uint32_t)run;
        run += t;
    }
    if (tid == 0) arr[n] = (uint32_t)total;
    __syncthreads();
    return total;
}

// ---------------------------------------------------------------------------------------------
// column context of one lane: evaluation weights of its (col, slice) and a 4-tap sliding window
// of collapsed column coefficients G[a] = sum_{b,g} wc[b] ws[g] L[a][bc+b][bs+g]
struct N4Col {
    int bc, bs;
    float wc0, wc1, wc2, wc3;
    float4 ws;
};
template <class CTX> VB_DEV N4Col n4_col(const CTX& c, const N4Axis& ac, const N4Axis& as, int q) {
    N4Col k;
    const int cc = q / c.n2, ss = q - cc * c.n2;
    k.bc = ac.bev[cc]; k.bs = as.bev[ss];
    const float4 wc = ac.evw[cc];
    k.wc0 = wc.x; k.wc1 = wc.y; k.wc2 = wc.z; k.wc3 = wc.w;
    k.ws = as.evw[ss];
    return k;
}
template <class CTX> VB_DEV float n4_coef(const float* L, const CTX& c, const N4Col& k, int a) {
    const float* lp = L + ((size_t)a * c.ncp[1] + k.bc) * c.ncp[2] + k.bs;
    const int st = c.ncp[2];
    float g = 0.0f;
    g += k.wc0 * (k.ws.x * lp[0] + k.ws.y * lp[1] + k.ws.z * lp[2] + k.ws.w * lp[3]);
    lp += st;
    g += k.wc1 * (k.ws.x * lp[0] + k.ws.y * lp[1] + k.ws.z * lp[2] + k.ws.w * lp[3]);
    lp += st;
    g += k.wc2 * (k.ws.x * lp[0] + k.ws.y * lp[1] + k.ws.z * lp[2] + k.ws.w * lp[3]);
    lp += st;
    g += k.wc3 * (k.ws.x * lp[0] + k.ws.y * lp[1] + k.ws.z * lp[2] + k.ws.w * lp[3]);
    return g;
}
struct N4Win {
    float g0, g1, g2, g3;
    int cur;
};
// `G` (optional): this lane's column of a precomputed coefficient table, G[a * gs]
template <class CTX> VB_DEV float n4_coef_t(const float* L, const CTX& c, const N4Col& k, const float* G, int gs, int a) {
    return G ? G[(size_t)a * gs] : n4_coef(L, c, k, a);
}
template <class CTX> VB_DEV void n4_win_advance(N4Win& w, const float* L, const CTX& c, const N4Col& k, int base, const float* G = nullptr, int gs = 0) {
    if (base == w.cur) return;
    if (w.cur < 0 || base - w.cur >= 4 || base < w.cur) {
        w.g0 = n4_coef_t(L, c, k, G, gs, base); w.g1 = n4_coef_t(L, c, k, G, gs, base + 1);
        w.g2 = n4_coef_t(L, c, k, G, gs, base + 2); w.g3 = n4_coef_t(L, c, k, G, gs, base + 3);
        w.cur = base;
    } else {
        while (w.cur < base) {
            w.g0 = w.g1; w.g1 = w.g2; w.g2 = w.g3; w.g3 = n4_coef_t(L, c, k, G, gs, w.cur + 4);
            ++w.cur;
        }
    }
}
// the same window over a precomputed coefficient column G[a * gs] (no lattice, no column weights)
VB_DEV void n4_win_advance_tab(N4Win& w, int base, const float* G, int gs) {
    if (base == w.cur) return;
    if (w.cur < 0 || base - w.cur >= 4 || base < w.cur) {
        w.g0 = G[(size_t)base * gs]; w.g1 = G[(size_t)(base + 1) * gs];
        w.g2 = G[(size_t)(base + 2) * gs]; w.g3 = G[(size_t)(base + 3) * gs];
        w.cur = base;
    } else {
        while (w.cur < base) {
            w.g0 = w.g1; w.g1 = w.g2; w.g2 = w.g3; w.g3 = G[(size_t)(w.cur + 4) * gs];
            ++w.cur;
        }
    }
}
VB_DEV float n4_win_eval(const N4Win& w, const float4 rw) {
    return rw.x * w.g0 + rw.y * w.g1 + rw.z * w.g2 + rw.w * w.g3;
}

// x / d from a correctly rounded reciprocal rd = RN(1/d): one Newton correction of the quotient.
// Gives the IEEE-rounded quotient (checked against 2e7 random float32 divisions in the ranges
// used here); 3 instructions instead of a full division per voxel.
VB_DEV float n4_div(float x, float d, float rd) {
    const float q = __fmul_rn(x, rd);
    return fmaf(fmaf(-q, d, x), rd, q);
}

// column-coefficient table of lattice L for every active column: dst[a][k]
template <class CTX> VB_DEV void n4_build_G(const CTX& c, const float* L, float* dst) {
    const N4Axis ac = n4_ax(c, 1), as = n4_ax(c, 2);
    const int total = c.ncp[0] * c.nact;
    for (int e = (int)threadIdx.x; e < total; e += (int)blockDim.x) {
        const int a = e / c.nact, k = e - a * c.nact;
        const N4Col col = n4_col(c, ac, as, (int)c.acol[k]);
        dst[(size_t)a * c.gstride + k] = n4_coef(L, c, col, a);
    }
    __syncthreads();
}

// exp(x) for the small arguments of the convergence measure: 1 + (x + x^2/2 + ...), whose final
// addition is (almost always) the correctly rounded float, like glibc's expf in the oracle.
VB_DEV float n4_exp_small(float x) {
    if (fabsf(x) < 0.03125f) {
        const float x2 = x * x;
        float p = 1.0f / 120.0f;
        p = fmaf(p, x, 1.0f / 24.0f);
        p = fmaf(p, x, 1.0f / 6.0f);
        p = fmaf(p, x, 0.5f);
        return 1.0f + fmaf(p, x2, x);
    }
    return expf(x);
}

// Work item of a warp: 32 consecutive ACTIVE columns (columns that carry mask voxels) x one
// segment of rows.  Splitting the rows gives the 16 warps enough items to balance.
struct N4Item {
    int blk, seg, wi0, wi1, q, k;        // k: index of this lane's column among the active columns
    bool valid, empty;
};
template <class CTX> VB_DEV N4Item n4_item(const CTX& c, int it) {
    N4Item I;
    I.blk = it % c.nblk; I.seg = it / c.nblk;
    I.wi0 = I.seg * c.wps; I.wi1 = min(c.W, I.wi0 + c.wps);
    const int k = I.blk * 32 + vb_lane();
    I.valid = k < c.nact;
    I.k = I.valid ? k : c.nact - 1;
    I.q = (int)c.acol[I.k];
    I.empty = c.wbase[(size_t)I.blk * c.W + I.wi1] == c.wbase[(size_t)I.blk * c.W + I.wi0];
    return I;
}

// Dynamic scheduler: the warps of a team draw the next (largest remaining) item from a shared
// counter.  The counter is reset by n4_items_begin (followed by a team barrier in the caller).
template <class CTX> VB_DEV void n4_items_begin(const CTX& c, const N4Team& T) {
    if (T.tid == 0) c.ctr[T.bar] = 0;
}
template <class CTX> VB_DEV int n4_items_next(const CTX& c, const N4Team& T) {
    int i = 0;
    if ((T.tid & 31) == 0) i = atomicAdd(&c.ctr[T.bar], 1);
    return __shfl_sync(VB_FULL, i, 0);
}
#define VB_N4_FOR_ITEMS(c, T, it) \
    for (int i_ = n4_items_next(c, T) * (c).step() + (c).off(), it = 0; i_ < (c).nblk * (c).nseg && ((it = (int)(c).iorder[i_]), true); \
         i_ = n4_items_next(c, T) * (c).step() + (c).off())

// What the row walks need from the context, as opaque local copies: otherwise the compiler
// re-derives every pointer from the kernel parameters (blockIdx * stride + base, or shared-memory
// base + offsets: up to 5 instructions) at each use inside the loops.
struct N4Walk {
    const uint32_t* colmask;
    const uint32_t* rbb;
    int ncols, nblk, n0;
};
template <class CTX> VB_DEV N4Walk n4_walk(const CTX& c) {
    N4Walk w;
    w.colmask = c.colmask; w.rbb = c.rbb; w.ncols = c.ncols; w.nblk = c.nblk; w.n0 = c.n0;
    VB_OPAQUE_PTR(w.colmask);
    VB_OPAQUE_PTR(w.rbb);
    return w;
}

// Walk the mask voxels of one item row by row.  The compact arrays are in C ORDER (row-major over
// the mask voxels of the volume, the order of the reference's convergence loop): the voxels of one
// row inside a block of 32 consecutive active columns are contiguous there, starting at the rank
// rbb[row][block].  BODY sees: r (row), bit (this lane's voxel is in the mask), idx (compact index =
// C-order rank of this lane's voxel when bit), bal (lane mask of the row).  Rows without any mask
// voxel in the block are skipped; the per-(row, block) bases of a 32-row word are fetched once by
// the lanes and broadcast by shuffle.
#define VB_N4_FOR_ROWS(c, I, ...)                                                            \
    {                                                                                        \
        const unsigned lt_ = vb_lanemask_lt();                                                     \
        for (int wi_ = (I).wi0; wi_ < (I).wi1; ++wi_) {                                      \
            const uint32_t mw_ = (I).valid ? (c).colmask[(size_t)wi_ * (c).ncols + (I).q] : 0u; \
            uint32_t rem_ = __reduce_or_sync(VB_FULL, mw_);                                  \
            if (!rem_) continue;                                                             \
            const int rr_ = wi_ * 32 + vb_lane();                                            \
            const unsigned rbw_ = (rr_ < (c).n0) ? (c).rbb[(size_t)rr_ * (c).nblk + (I).blk] : 0u; \
            while (rem_) {                                                                   \
                const int b_ = __popc((rem_ - 1u) & ~rem_);                                  \
                rem_ &= rem_ - 1u;                                                           \
                const int r = wi_ * 32 + b_;                                                 \
                const bool bit = (mw_ >> b_) & 1u;                                           \
                const unsigned bal = __ballot_sync(VB_FULL, bit);                            \
                const unsigned idx = __shfl_sync(VB_FULL, rbw_, b_) + (unsigned)__popc(bal & lt_); \
                __VA_ARGS__                                                                  \
            }                                                                                \
        }                                                                                    \
    }

// ---------------------------------------------------------------------------------------------
// setup: column bitmasks, compaction tables, compact log image, range and count of the mask
template <class CTX> VB_DEV void n4_setup(CTX& c, int label, float* smem_logI, int smem_cap, float& bmin, float& bmax, int& M) {
    const int nt = (int)blockDim.x, tid = (int)threadIdx.x, lane = vb_lane(), w = vb_warp(), nw = nt >> 5;
    // column bitmasks (coalesced over columns) and the list of active columns
    uint32_t* flag = c.rbb;                               // scratch: [ncols + 1]
    for (int q = tid; q < c.ncols; q += nt) {
        uint32_t any = 0;
        for (int wi = 0; wi < c.W; ++wi) {
            uint32_t bits = 0;
            const int rend = min(32, c.n0 - wi * 32);
            for (int b0 = 0; b0 < rend; b0 += 8) {
                uint8_t mv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    mv[j] = (b0 + j < rend) ? vb_ld_stream(c.msk + (size_t)(wi * 32 + b0 + j) * c.ncols + q) : (uint8_t)0;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (b0 + j < rend && (int)mv[j] == label) bits |= 1u << (b0 + j);
            }
            c.colmask[(size_t)wi * c.ncols + q] = bits;
            any |= bits;
        }
        flag[q] = any ? 1u : 0u;
    }
    __syncthreads();
    c.nact = n4_array_scan(flag, c.ncols, (int*)c.red);
    for (int q = tid; q < c.ncols; q += nt)
        if (flag[q + 1] != flag[q]) c.acol[flag[q]] = (uint32_t)q;
    c.nblk = (c.nact + 31) >> 5;
    // row segments: enough items to keep every warp busy (about three items per warp)
    c.nseg = 1;
    while (c.nseg < min(c.W, 8) && c.nblk * c.nseg < 3 * nw) c.nseg <<= 1;
    c.wps = (c.W + c.nseg - 1) / c.nseg;
    c.nseg = (c.W + c.wps - 1) / c.wps;
    __syncthreads();                                      // flag (= rbb) is rewritten below
    // per (row, block) and per (block, word) voxel counts
    for (int blk = w; blk < c.nblk; blk += nw) {
        const int k = blk * 32 + lane;
        const bool valid = k < c.nact;
        const int q = valid ? (int)c.acol[k] : 0;
        for (int wi = 0; wi < c.W; ++wi) {
            const uint32_t mw = valid ? c.colmask[(size_t)wi * c.ncols + q] : 0u;
            const int rend = min(32, c.n0 - wi * 32);
            int tot = 0, mine = 0;
            for (int b = 0; b < rend; ++b) {
                const int cnt = __popc(__ballot_sync(VB_FULL, (mw >> b) & 1u));
                if (lane == b) mine = cnt;
                tot += cnt;
            }
            if (lane < rend) c.rbb[(size_t)(wi * 32 + lane) * c.nblk + blk] = (uint32_t)mine;
            if (lane == 0) c.wbase[(size_t)blk * c.W + wi] = (uint32_t)tot;
        }
    }
    __syncthreads();
    M = n4_array_scan(c.rbb, c.n0 * c.nblk, (int*)c.red);
    n4_array_scan(c.wbase, c.nblk * c.W, (int*)c.red);
    // histogram fixed point: items are handed out dynamically, so a histogram copy may receive
    // any number of voxels up to M
    {
        int bits = 0;
        while ((1LL << bits) <= (long long)M) ++bits;
        c.hshift = min(24, 31 - bits);
    }
    // work items, largest first (rank sort; ties by index), for the dynamic scheduler
    {
        const int ni = c.nblk * c.nseg;
        for (int a = tid; a < ni; a += nt) {
            const int blk = a % c.nblk, wi0 = (a / c.nblk) * c.wps, wi1 = min(c.W, wi0 + c.wps);
            const unsigned ca = c.wbase[(size_t)blk * c.W + wi1] - c.wbase[(size_t)blk * c.W + wi0];
            int rank = 0;
            for (int b = 0; b < ni; ++b) {
                const int bb = b % c.nblk, w0 = (b / c.nblk) * c.wps, w1 = min(c.W, w0 + c.wps);
                const unsigned cb = c.wbase[(size_t)bb * c.W + w1] - c.wbase[(size_t)bb * c.W + w0];
                rank += (cb > ca) || (cb == ca && b < a);
            }
            c.iorder[rank] = (uint32_t)a;
        }
        __syncthreads();
    }
    c.u_smem = M <= smem_cap;
    if (c.u_smem) c.U = smem_logI;                     // the compact U image fits shared memory

    float tmin = FLT_MAX, tmax = -FLT_MAX;
    const unsigned lt = (1u << lane) - 1u;
    for (int it = w; it < c.nblk * c.nseg; it += nw) {
        const N4Item I = n4_item(c, it);
        if (I.empty) continue;
        for (int wi = I.wi0; wi < I.wi1; ++wi) {
            const uint32_t mw = I.valid ? c.colmask[(size_t)wi * c.ncols + I.q] : 0u;
            const uint32_t wm = __reduce_or_sync(VB_FULL, mw);
            if (!wm) continue;
            const int rr = wi * 32 + lane;
            const unsigned rbw = (rr < c.n0) ? c.rbb[(size_t)rr * c.nblk + I.blk] : 0u;   // C-order rank of the first voxel of (row, block)
#pragma unroll 1
            for (int b0 = 0; b0 < 32; b0 += 8) {
                if (((wm >> b0) & 0xFFu) == 0) continue;
                float x[8];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    x[j] = ((mw >> (b0 + j)) & 1u) ? vb_ld_stream(c.img + (size_t)(wi * 32 + b0 + j) * c.ncols + I.q) : 0.0f;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const bool bit = (mw >> (b0 + j)) & 1u;
                    const unsigned bal = __ballot_sync(VB_FULL, bit);
                    const unsigned base = __shfl_sync(VB_FULL, rbw, b0 + j);
                    if (bit) {
                        const float l = (x[j] > 0.0f) ? logf(x[j]) : x[j];
                        const unsigned at = base + (unsigned)__popc(bal & lt);
                        c.logI[at] = l;
                        c.U[at] = l;                     // zero lattice: U = log image
                        if (c.code) c.code[at] = ((uint32_t)(wi * 32 + b0 + j) << 16) | (uint32_t)I.k;
                        tmin = fminf(tmin, l);
                        tmax = fmaxf(tmax, l);
                    }
                }
            }
        }
    }
    bmin = vb_block_reduce(tmin, VbMinF(), FLT_MAX, (float*)c.red);
    bmax = vb_block_reduce(tmax, VbMaxF(), -FLT_MAX, (float*)c.red);
}

// cluster mode: combine this CTA's (min, max, s1, s2) with those of the other CTAs, in rank order (every
// CTA ends up with the same four numbers, bit for bit)
template <class CTX> VB_DEV void n4_cluster_reduce4(CTX& c, const N4Team& T, float& mn, float& mx, double& s1, double& s2) {
    if (c.k() == 1) return;
    // (two sets of slots used in turn: a set is rewritten two exchanges later, i.e. after another cluster
    // barrier that every CTA reaches only when it has read this one -- no second barrier per exchange)
    double* slots = c.cl_red + (size_t)c.par_red * 4 * c.k();
    c.par_red ^= 1;
    if (T.tid == 0) {
        double* mine = slots + 4 * c.rank();
        mine[0] = (double)mn; mine[1] = (double)mx; mine[2] = s1; mine[3] = s2;
    }
    n4_csync(c, T);
    float a = FLT_MAX, b = -FLT_MAX;
    double p = 0.0, q = 0.0;
    for (int r = 0; r < c.k(); ++r) {
        a = fminf(a, (float)vb_ld_cg(slots + 4 * r));
        b = fmaxf(b, (float)vb_ld_cg(slots + 4 * r + 1));
        p += vb_ld_cg(slots + 4 * r + 2);
        q += vb_ld_cg(slots + 4 * r + 3);
    }
    mn = a; mx = b; s1 = p; s2 = q;
}

// range of U = logI - bias(L) over the mask (after a lattice refinement)
template <class CTX> VB_DEV void n4_range(CTX& c, float& bmin, float& bmax) {
    const int nt = (int)blockDim.x, lane = vb_lane(), w = vb_warp(), nw = nt >> 5;
    const N4Axis ac = n4_ax(c, 1), as = n4_ax(c, 2);
    float tmin = FLT_MAX, tmax = -FLT_MAX;
    for (int it = w + nw * c.rank(); it < c.nblk * c.nseg; it += nw * c.k()) {      // (cluster mode: the items are dealt over the CTAs)
        const N4Item I = n4_item(c, it);
        if (I.empty) continue;
        const N4Col col = n4_col(c, ac, as, I.q);
        const float* Gl = c.G ? c.G + I.k : nullptr;
        N4Win win; win.cur = -1; win.g0 = win.g1 = win.g2 = win.g3 = 0.0f;
        VB_N4_FOR_ROWS(c, I, {
            n4_win_advance(win, c.L, c, col, c.r_bev[r], Gl, c.gstride);
            if (bit) {
                const float u = c.logI[idx] - n4_win_eval(win, c.r_evw[r]);
                c.U[idx] = u;
                tmin = fminf(tmin, u);
                tmax = fmaxf(tmax, u);
            }
        })
    }
    bmin = vb_block_reduce(tmin, VbMinF(), FLT_MAX, (float*)c.red);
    bmax = vb_block_reduce(tmax, VbMaxF(), -FLT_MAX, (float*)c.red);
    if (c.k() > 1) {
        N4Team full;
        full.tid = (int)threadIdx.x; full.nt = nt; full.bar = 0;
        double z1 = 0.0, z2 = 0.0;
        n4_cluster_reduce4(c, full, bmin, bmax, z1, z2);
    }
}

// one voxel of the Parzen histogram: its unit weight split between two neighbouring bins of this lane's copy
VB_DEV void n4_hist_add(vb_saddr sh, float u, float bmin, float slope, float rslope, int nbins, uint32_t one, float fone) {
    const float cidx = n4_div(u - bmin, slope, rslope);
    const float fl = floorf(cidx);
    const int bi = (int)fl;
    const float off = cidx - fl;
    if (off == 0.0f) {
        if (bi >= 0 && bi < nbins) vb_reds_add_u32(sh + 4u * (unsigned)bi, one);
    } else if (bi >= 0 && bi < nbins - 1) {
        const uint32_t o = (uint32_t)__float2uint_rn(off * fone);
        vb_reds_add_u32(sh + 4u * (unsigned)bi, one - o);
        vb_reds_add_u32(sh + 4u * (unsigned)bi + 4u, o);
    }
}

// ---------------------------------------------------------------------------------------------
// P1: triangular Parzen histogram of U over the mask (SharpenImage, first half)
template <class CTX> VB_DEV void n4_histogram(CTX& c, const N4Team& T, float bmin, float bmax) {
    const int nt = T.nt, tid = T.tid;
    // One copy per LANE with an odd stride: lanes of a warp never share an address, and lanes that
    // hit the same bin fall into 32 different banks (the U values cluster in a few bins, so copies
    // per warp serialised most of the atomics: 21 -> 11 us per iteration flat, -> with lane copies)
    const int hstride = c.nbins | 1;
    uint32_t* hc = (uint32_t*)c.un_f;                    // [VB_N4_HCOPIES][hstride]
    for (int i = tid; i < hstride * VB_N4_HCOPIES; i += nt) hc[i] = 0u;
    n4_sync(T);
    const float slope = (bmax - bmin) / (float)(c.nbins - 1);
    const float rslope = __frcp_rn(slope);
    const uint32_t one = 1u << c.hshift;
    const float fone = (float)one;
    uint32_t* h = hc + (size_t)(tid & (VB_N4_HCOPIES - 1)) * hstride;
    const vb_saddr sh = vb_saddr_of(h);                 // this lane's copy, as a 32-bit shared address (RED.shared)
    // U is stored per voxel, so the histogram needs no geometry: a flat, fully populated loop over
    // the compact array (the fixed-point sums do not depend on the order)
    const float* U = c.U;
    VB_OPAQUE_PTR(U);
    const int nbins = c.nbins;
    // cluster mode: this CTA's slice of the compact array (U is in global memory then; read through L2)
    const int M0 = (int)((long long)c.M * c.rank() / c.k()), M = (int)((long long)c.M * (c.rank() + 1) / c.k());
    if (VB_N4_HIST4 && c.u_smem) {                       // U in shared memory: four voxels per 16-byte LDS
        const vb_saddr sU = vb_saddr_of(c.U);
        const int M4 = M & ~3;
#pragma unroll 2
        for (int i = 4 * tid; i < M4; i += 4 * nt) {
            const float4 v = vb_lds_f32x4(sU + 4u * (unsigned)i);
            n4_hist_add(sh, v.x, bmin, slope, rslope, nbins, one, fone);
            n4_hist_add(sh, v.y, bmin, slope, rslope, nbins, one, fone);
            n4_hist_add(sh, v.z, bmin, slope, rslope, nbins, one, fone);
            n4_hist_add(sh, v.w, bmin, slope, rslope, nbins, one, fone);
        }
        for (int i = M4 + tid; i < M; i += nt)
            n4_hist_add(sh, vb_lds_f32(sU + 4u * (unsigned)i), bmin, slope, rslope, nbins, one, fone);
    } else {
#pragma unroll 4
        for (int i = M0 + tid; i < M; i += nt) n4_hist_add(sh, n4_ld_x(c, U + i), bmin, slope, rslope, nbins, one, fone);
    }
    n4_sync(T);
    if (c.k() > 1) {                                       // partial sums of the CTAs meet in global memory (two sets used in turn, see n4_cluster_reduce4)
        const int ts = (T.bar != 0) ? 1 : 0;
        unsigned long long* part = c.cl_hist + (size_t)(2 * ts + (ts ? c.par_hist1 : c.par_hist0)) * c.k() * c.nbins;
        if (ts) c.par_hist1 ^= 1; else c.par_hist0 ^= 1;
        for (int bin = tid; bin < c.nbins; bin += nt) {
            unsigned long long s = 0;
#pragma unroll
            for (int k = 0; k < VB_N4_HCOPIES; ++k) s += hc[(size_t)k * hstride + bin];
            part[(size_t)c.rank() * c.nbins + bin] = s;
        }
        n4_csync(c, T);
        for (int bin = tid; bin < c.nbins; bin += nt) {
            unsigned long long s = 0;
            for (int r = 0; r < c.k(); ++r) s += vb_ld_cg(part + (size_t)r * c.nbins + bin);
            c.H[bin] = (float)((double)s / (double)one);
        }
        n4_sync(T);
        return;
    }
    for (int bin = tid; bin < c.nbins; bin += nt) {
        unsigned long long s = 0;
#pragma unroll
        for (int k = 0; k < VB_N4_HCOPIES; ++k) s += hc[(size_t)k * hstride + bin];
        c.H[bin] = (float)((double)s / (double)one);
    }
    n4_sync(T);
}

// (Running the whole sharpening step on a sub-team of four or eight warps behind its own named barrier
// was measured slower: 14.6 -> 16.3 us per iteration, the pointwise passes lose more than the barriers gain.)
// (Out-of-place transforms with three radix-2 stages per pass -- bit reversal folded into the first pass,
// 13 instead of 21 team barriers per sharpening step -- measured 14.2 -> 14.1 us: the step is not bound by its
// barriers; the 16 doubles a thread then holds cost the other phases registers: 74.4 -> 78.0 ms per 1776 subjects.)
// in-place radix-2 FFT on shared (re, im); tw = exp(-2 pi i k / N); sign<0 forward, >0 backward
VB_DEV void n4_fft(const N4Team& T, double* re, double* im, int N, int logN, int sign, const double* twr, const double* twi) {
    const int nt = T.nt, tid = T.tid;
    for (int i = tid; i < N; i += nt) {
        const int j = (int)(__brev((unsigned)i) >> (32 - logN));
        if (i < j) {
            double t = re[i]; re[i] = re[j]; re[j] = t;
            t = im[i]; im[i] = im[j]; im[j] = t;
        }
    }
    n4_sync(T);
    // Two radix-2 stages (len, 2 len) per pass: a thread owns the four elements
    // {i, i + len/2, i + len, i + len + len/2} of a block of 2 len and does both stages on them
    // (the same butterflies, twiddles and operation order as the plain radix-2 loop, so the
    // results are bit-identical) -- half the passes over shared memory and half the barriers.
    // A warp covers 128 consecutive elements per pass, so while the next pass works on aligned
    // blocks of <= 128 elements it only needs the warp to be in step, not the team.
    int len = 2;
    for (; 2 * len <= N; len <<= 2) {
        const int half = len >> 1, step1 = N / len, step2 = N / (2 * len);
        for (int q = tid; q < (N >> 2); q += nt) {
            const int grp = q / half, k = q - grp * half;
            const int i0 = grp * 2 * len + k, i1 = i0 + half, i2 = i0 + len, i3 = i2 + half;
            const double w1r = twr[k * step1], w1i = (sign < 0) ? twi[k * step1] : -twi[k * step1];
            // stage len: (i0, i1) and (i2, i3)
            double ar = re[i0], ai = im[i0], br = re[i1], bi = im[i1];
            double xr = br * w1r - bi * w1i, xi = br * w1i + bi * w1r;
            br = ar - xr; bi = ai - xi; ar += xr; ai += xi;
            double cr = re[i2], ci = im[i2], dr = re[i3], di = im[i3];
            xr = dr * w1r - di * w1i; xi = dr * w1i + di * w1r;
            dr = cr - xr; di = ci - xi; cr += xr; ci += xi;
            // stage 2 len: (i0, i2) with twiddle k, (i1, i3) with twiddle k + half
            const double w2r = twr[k * step2], w2i = (sign < 0) ? twi[k * step2] : -twi[k * step2];
            const double w3r = twr[(k + half) * step2], w3i = (sign < 0) ? twi[(k + half) * step2] : -twi[(k + half) * step2];
            xr = cr * w2r - ci * w2i; xi = cr * w2i + ci * w2r;
            re[i2] = ar - xr; im[i2] = ai - xi; re[i0] = ar + xr; im[i0] = ai + xi;
            xr = dr * w3r - di * w3i; xi = dr * w3i + di * w3r;
            re[i3] = br - xr; im[i3] = bi - xi; re[i1] = br + xr; im[i1] = bi + xi;
        }
        if (8 * len <= 128) __syncwarp(); else n4_sync(T);
    }
    if (len <= N) {                                          // odd number of stages: the last one alone
        const int half = len >> 1, step = N / len;
        for (int b = tid; b < (N >> 1); b += nt) {
            const int grp = b / half, k = b - grp * half;
            const int i0 = grp * len + k, i1 = i0 + half;
            const double wr = twr[k * step], wi = (sign < 0) ? twi[k * step] : -twi[k * step];
            const double xr = re[i1] * wr - im[i1] * wi, xi = re[i1] * wi + im[i1] * wr;
            re[i1] = re[i0] - xr; im[i1] = im[i0] - xi;
            re[i0] += xr; im[i0] += xi;
        }
        n4_sync(T);
    }
}

// SharpenImage, second half: Wiener-deconvolve the histogram with the Gaussian bias model and
// build E(u|v).  Two real sequences share one complex FFT (4 FFTs instead of ITK's 7).
// The twiddles depend only on the (launch-constant) padded histogram size: c.tw, computed once per CTA.
template <class CTX> VB_DEV void n4_twiddles(CTX& c) {
    const int N = c.fftN;
    for (int n = (int)threadIdx.x; n < (N >> 1); n += (int)blockDim.x) {
        double sn, co;
        vb_sincospi(-2.0 * (double)n / (double)N, &sn, &co);
        c.tw[n] = co; c.tw[(N >> 1) + n] = sn;
    }
    __syncthreads();
}
template <class CTX> VB_DEV void n4_sharpen(CTX& c, const N4Team& T, float bmin, float bmax, float fwhm, float wiener) {
    const int nt = T.nt, tid = T.tid;
    const int N = c.fftN;
    int logN = 0;
    while ((1 << logN) < N) ++logN;
    const int hoff = (int)(0.5 * (double)(N - c.nbins));
    double* Ar = c.un_d;          double* Ai = Ar + N;
    double* Br = Ai + N;          double* Bi = Br + N;
    double* Fr = Bi + N;          double* Fi = Fr + N;     // Ff
    const double* twr = c.tw;     const double* twi = twr + (N >> 1);
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
    }
    n4_sync(T);
    n4_fft(T, Ar, Ai, N, logN, -1, twr, twi);
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
    n4_sync(T);
    n4_fft(T, Br, Bi, N, logN, +1, twr, twi);
    for (int n = tid; n < N; n += nt) {
        const double u = Br[n] > 0.0 ? Br[n] : 0.0;
        const float x = bmin + ((float)n - (float)hoff) * slope;
        Ar[n] = (double)x * u;
        Ai[n] = u;
    }
    n4_sync(T);
    n4_fft(T, Ar, Ai, N, logN, -1, twr, twi);
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
    n4_sync(T);
    n4_fft(T, Br, Bi, N, logN, +1, twr, twi);
    for (int n = tid; n < c.nbins; n += nt) {
        const double den = Bi[n + hoff];
        c.E[n] = (den != 0.0) ? (float)(Br[n + hoff] / den) : 0.0f;
    }
    n4_sync(T);
}

// ---------------------------------------------------------------------------------------------
// Separable B-spline contraction over the mask.
//   OMEGA = true : value 1, weights B^2            -> c.Om
//   OMEGA = false: value U - E(U), weights B^3/S   -> phi = delta/omega; Lold = L; L += phi
// Rows: a lane reduces its column (one row segment) to row coefficients through a sliding 4-tap
// window; a coefficient that leaves the window is spread over the 4 column control points into
// Q[a][b][s] with 64-bit FIXED-POINT shared-memory atomics (exactly associative => the result
// does not depend on the order in which warps arrive).  Slices -> lattice afterwards.
#define VB_N4_QSCALE 1099511627776.0f         /* 2^40: sums stay below 2^63 for |sum| < 8e6 */
// 64-bit add from two native 32-bit shared-memory atomics: every wrap of the low word is seen
// by exactly one adder (the one whose old value + addend overflowed), which forwards the carry.
// Once all adders are done the pair holds the exact 64-bit sum, whatever the interleaving.
// Q is addressed through its 32-bit shared address (ATOMS / RED.shared, 32-bit address arithmetic).
VB_DEV void n4_add64(vb_saddr q, long long x) {
    const unsigned xlo = (unsigned)(unsigned long long)x;
    const unsigned old = vb_atoms_add_u32(q, xlo);
    const unsigned carry = (unsigned)(old + xlo < old);
    vb_reds_add_u32(q + 4u, (unsigned)((unsigned long long)x >> 32) + carry);
}
VB_DEV void n4_qflush(vb_saddr Q, int a, int ncp1, int n2, int bc, int s, const float4 wc, float v) {
    if (v == 0.0f) return;
    const vb_saddr qp = Q + 8u * (unsigned)((a * ncp1 + bc) * n2 + s);
    const unsigned st = 8u * (unsigned)n2;
    n4_add64(qp, __float2ll_rn(wc.x * v * VB_N4_QSCALE));
    n4_add64(qp + st, __float2ll_rn(wc.y * v * VB_N4_QSCALE));
    n4_add64(qp + 2u * st, __float2ll_rn(wc.z * v * VB_N4_QSCALE));
    n4_add64(qp + 3u * st, __float2ll_rn(wc.w * v * VB_N4_QSCALE));
}

// SM: U and the row tables live in shared memory (LDS through 32-bit shared addresses in the walk).
template <bool OMEGA, bool SM, class CTX>
VB_DEV void n4_contract_t(CTX& c, const N4Team& T, float bmin, float bmax) {
    const int nt = T.nt, tid = T.tid;
    const int ncp0 = c.ncp[0], ncp1 = c.ncp[1], ncp2 = c.ncp[2];
    const int n2 = c.n2;
    const N4Axis ar = n4_ax(c, 0), ac = n4_ax(c, 1), as = n4_ax(c, 2);
    const float slope = (bmax - bmin) / (float)(c.nbins - 1);
    const float rslope = __frcp_rn(slope);
    const int nbm1 = c.nbins - 1;
    // (private copies of Q per warp index were measured slower: 31.9 -> 34.9 us per iteration)
    unsigned* Q = (unsigned*)c.un_d;                     // [ncp0][ncp1][n2] x {low, high} words
    const N4Walk wk = n4_walk(c);
    const float* U = c.U;
    const float4* rfit3 = c.r_fit3;
    const int* rbfit = c.r_bfit;
    VB_OPAQUE_PTR(Q); VB_OPAQUE_PTR(U); VB_OPAQUE_PTR(rfit3); VB_OPAQUE_PTR(rbfit);
    const vb_saddr sQ = vb_saddr_of(Q), sE = vb_saddr_of(c.E);
    const vb_saddr sU = SM ? vb_saddr_of(c.U) : 0, sW = SM ? vb_saddr_of(c.r_fit3) : 0, sB = SM ? vb_saddr_of(c.r_bfit) : 0;
    const int Qn = ncp0 * ncp1 * n2;
    for (int e = tid; e < 2 * Qn; e += nt) Q[e] = 0u;
    // cluster mode: the fit's items are dealt over the CTAs (item j of this CTA = j * K + rank); omega is
    // computed by every CTA for itself (once per level)
    c.deal(OMEGA ? 1 : c.k(), OMEGA ? 0 : c.rank());
    n4_items_begin(c, T);
    n4_sync(T);

    VB_N4_FOR_ITEMS(c, T, it) {
        const N4Item I = n4_item(c, it);
        if (I.empty) continue;
        const int cc = I.q / n2, ss = I.q - cc * n2;
        const int bc = ac.bfit[cc];
        const float4 wc = OMEGA ? ac.fit2[cc] : ac.fit3[cc];
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int cur = -1;                                    // window = coefficients cur .. cur+3
        VB_N4_FOR_ROWS(wk, I, {
            const int base = SM ? vb_lds_s32(sB + 4u * (unsigned)r) : rbfit[r];
            if (cur < 0) cur = base;
            while (cur < base) {                         // warp uniform
                if (I.valid) n4_qflush(sQ, cur, ncp1, n2, bc, ss, wc, a0);
                a0 = a1; a1 = a2; a2 = a3; a3 = 0.f;
                ++cur;
            }
            if (bit) {
                float val;
                float4 wt;
                if (OMEGA) {
                    val = 1.0f;
                    wt = ar.fit2[r];
                } else {
                    const float u = SM ? vb_lds_f32(sU + 4u * idx) : n4_ld_x(c, U + idx);
                    const float cidx = n4_div(u - bmin, slope, rslope);
                    const float fl = floorf(cidx);
                    const int bi = (int)fl;
                    float s;
                    if (bi >= 0 && bi < nbm1) {
                        const float e0 = vb_lds_f32(sE + 4u * (unsigned)bi), e1 = vb_lds_f32(sE + 4u * (unsigned)bi + 4u);
                        s = e0 + (e1 - e0) * (cidx - fl);
                    } else {
                        s = vb_lds_f32(sE + 4u * (unsigned)nbm1);
                    }
                    val = u - s;
                    wt = SM ? vb_lds_f32x4(sW + 16u * (unsigned)r) : rfit3[r];
                }
                a0 += wt.x * val; a1 += wt.y * val; a2 += wt.z * val; a3 += wt.w * val;
            }
        })
        if (I.valid && cur >= 0) {
            n4_qflush(sQ, cur, ncp1, n2, bc, ss, wc, a0);
            n4_qflush(sQ, cur + 1, ncp1, n2, bc, ss, wc, a1);
            n4_qflush(sQ, cur + 2, ncp1, n2, bc, ss, wc, a2);
            n4_qflush(sQ, cur + 3, ncp1, n2, bc, ss, wc, a3);
        }
    }
    n4_sync(T);
    c.deal(1, 0);
    if (!OMEGA && c.k() > 1) {                             // partial sums of the CTAs meet in global memory (integers: exact)
        const int ts = (T.bar != 0) ? 1 : 0;
        unsigned* part = c.cl_q + (size_t)(2 * ts + (ts ? c.par_q1 : c.par_q0)) * c.k() * 2 * c.q_max;     // (two sets used in turn, see n4_cluster_reduce4)
        if (ts) c.par_q1 ^= 1; else c.par_q0 ^= 1;
        unsigned* mine = part + (size_t)c.rank() * 2 * c.q_max;
        for (int e = tid; e < 2 * Qn; e += nt) mine[e] = Q[e];
        n4_csync(c, T);
        for (int e = tid; e < Qn; e += nt) {
            unsigned long long sum = 0;
            for (int r = 0; r < c.k(); ++r) {
                const unsigned* qr = part + (size_t)r * 2 * c.q_max + 2 * e;
                sum += ((unsigned long long)vb_ld_cg(qr + 1) << 32) | (unsigned long long)vb_ld_cg(qr);
            }
            Q[2 * e] = (unsigned)sum; Q[2 * e + 1] = (unsigned)(sum >> 32);
        }
        n4_sync(T);
    }

    // slices -> lattice
    const int Ln = ncp0 * ncp1 * ncp2;
    for (int e = tid; e < Ln; e += nt) {
        const int g = e % ncp2, ab = e / ncp2;
        float sum = 0.0f;
        const int shi = as.hi[g];
        for (int s = as.lo[g]; s < shi; ++s) {
            const int k = g - as.bfit[s];
            const float4 wt = OMEGA ? as.fit2[s] : as.fit3[s];
            const float wk = (k == 0) ? wt.x : (k == 1) ? wt.y : (k == 2) ? wt.z : wt.w;
            const unsigned* qc = Q + 2 * ((size_t)ab * n2 + s);
            const long long qi = (long long)(((unsigned long long)qc[1] << 32) | (unsigned long long)qc[0]);
            const float qv = (float)((double)qi * (1.0 / (double)VB_N4_QSCALE));
            sum += qv * wk;
        }
        if (OMEGA) {
            c.Om[e] = sum;
        } else {
            const float om = c.Om[e];
            // itk::Math::NotAlmostEquals(omega, 0): |omega| > 0.1 * FLT_EPSILON
            const float phi = (fabsf(om) > 0.1f * FLT_EPSILON) ? sum / om : 0.0f;
            c.Phi[e] = phi;                              // committed by n4_commit (the fit may be speculative)
        }
    }
    n4_sync(T);
}

template <bool OMEGA, class CTX>
VB_DEV void n4_contract(CTX& c, const N4Team& T, float bmin, float bmax) {
    if (!OMEGA && c.u_smem && c.rowtab) n4_contract_t<OMEGA, true>(c, T, bmin, bmax);
    else n4_contract_t<OMEGA, false>(c, T, bmin, bmax);
}

// ---------------------------------------------------------------------------------------------
// CalculateConvergenceMeasurement: sequential float Welford over e[0..M) (C order), reproduced
// in parallel.  One step of the reference (n4_oracle.c: n4_convergence) is
//     N += 1;  if (N > 1) sigma = f32(sigma + f64(f32(d*d)) * (N-1)/N),  d = f32(pixel - mu)
//     mu = f32(f64(mu) * (1 - 1/N) + f64(f32(pixel / N)))
// Both updates are "state + small term, rounded to float".  With 1/N = ch + cl split into two
// floats the terms are formed in float arithmetic to ~1e-14 (absolute), so the rounded state is
// the reference's except when the exact sum lies within that distance of a rounding boundary.
struct N4Step {
    float b;        // f32(pixel / N)
    float ch, cl;   // 1/N = ch + cl
};
VB_DEV N4Step n4_step_consts(float pixel, int n) {
    N4Step s;
    const float nf = (float)n;
    s.ch = __frcp_rn(nf);
    s.cl = __fmul_rn(fmaf(-s.ch, nf, 1.0f), s.ch);
    s.b = n4_div(pixel, nf, s.ch);                      // == RN(pixel / N), the reference's float division
    return s;
}
VB_DEV float n4_mu_next(float mu, const N4Step& s) {
    const float t = fmaf(-mu, s.cl, fmaf(-mu, s.ch, s.b));      // pixel/N - mu/N
    return __fadd_rn(mu, t);
}
VB_DEV float n4_sigma_next(float sigma, float mu_prev, float pixel, const N4Step& s) {
    const float d = __fsub_rn(pixel, mu_prev);
    const float dd = __fmul_rn(d, d);
    const float th = fmaf(-dd, s.ch, dd);                       // dd * (N-1)/N, head
    const float tl = fmaf(-dd, s.cl, fmaf(-dd, s.ch, __fsub_rn(dd, th))); // ... and tail
    const float r = __fadd_rn(sigma, th);                       // TwoSum: r + err == sigma + th exactly
    const float bb = __fsub_rn(r, sigma);
    const float err = __fadd_rn(__fsub_rn(sigma, __fsub_rn(r, bb)), __fsub_rn(th, bb));
    return __fadd_rn(r, __fadd_rn(err, tl));
}

// Exclusive block prefix of v with ONE barrier, which also ORs `flag` over the block (in/out).
// `tot` is a per-sweep buffer (double buffered by the caller).
VB_DEV int n4_chain_scan(const N4Team& T, int v, int& total, int* tot, int& flag) {
    const int lane = (T.tid & 31), w = (T.tid >> 5), nw = T.nt >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(VB_FULL, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) tot[w] = inc;
    flag = n4_sync_or(T, flag);
    const int tv = (lane < nw) ? tot[lane] : 0;            // nw <= 32: two warp reductions
    total = (int)__reduce_add_sync(VB_FULL, (unsigned)tv);
    const int woff = (int)__reduce_add_sync(VB_FULL, (unsigned)((lane < w) ? tv : 0));
    return woff + inc - v;
}
// the same for the 32 lanes of one warp: no shared memory, no barrier
VB_DEV int n4_warp_scan(int v, int& total, int& flag) {
    const int lane = vb_lane();
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(VB_FULL, inc, d);
        if (lane >= d) inc += t;
    }
    total = __shfl_sync(VB_FULL, inc, 31);
    flag = __any_sync(VB_FULL, flag);
    return inc - v;
}

// exclusive prefix of a float over the team (WARP = false: one barrier) or the warp; only used
// for start guesses, so the summation order does not matter
template <bool WARP>
VB_DEV float n4_guess_scan(const N4Team& T, float v, float* tot) {
    const int lane = (T.tid & 31), w = (T.tid >> 5), nw = T.nt >> 5;
    float inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const float t = __shfl_up_sync(VB_FULL, inc, d);
        if (lane >= d) inc += t;
    }
    float woff = 0.0f;
    if (!WARP) {
        if (lane == 31) tot[w] = inc;
        n4_sync(T);
        float tv = (lane < nw) ? tot[lane] : 0.0f;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const float t = __shfl_up_sync(VB_FULL, tv, d);
            if (lane >= d) tv += t;
        }
        woff = __shfl_sync(VB_FULL, tv, (w > 0) ? w - 1 : 0);
        if (w == 0) woff = 0.0f;
    }
    return woff + inc - v;
}

// sigma + (th + tl) rounded to float (TwoSum keeps the part of th that the first add drops)
VB_DEV float n4_sigma_add(float sigma, float th, float tl) {
    const float r = __fadd_rn(sigma, th);
    const float bb = __fsub_rn(r, sigma);
    const float err = __fadd_rn(__fsub_rn(sigma, __fsub_rn(r, bb)), __fsub_rn(th, bb));
    return __fadd_rn(r, __fadd_rn(err, tl));
}
VB_DEV int n4_binade(int bits) { return bits >> 23; }

// One chunk [start, start + len) of the chain: every participating thread runs its <= VB_N4_EPT
// consecutive steps exactly from a guessed start state; the starts are re-derived from the
// prefix sum of the threads' bit-pattern increments until none moves.  WARP = false: the
// threads of team T, one barrier per sweep.  WARP = true: the 32 lanes of ONE warp, no barrier
// at all (used for the first elements, where a chunk may not be longer than the number of
// elements before it and a block-wide sweep would cost a barrier for a handful of steps).
// (mu_b, sg_b): float bit patterns of the state at the chunk start (in) / end (out), uniform.
// PREDICT: the starts of the first mu sweep are guessed from the exact-arithmetic running mean
// mu(lo) = mu0 + sum_{start <= i < lo} (px_i - mu0) / lo (one extra float scan) instead of mu0.
// In the first few thousand elements the mean still moves by many ulps per element, and with
// this guess a chunk there may be much longer than the number of elements before it.
template <bool WARP, bool PREDICT, class CTX>
VB_DEV void n4_chain_chunk(CTX& c, const N4Team& T, const float* epx, int start, int len, int& mu_b, int& sg_b,
                           int (*s_tot)[32], int& buf, float* s_guess) {
    const int nt = WARP ? 32 : T.nt, tid = WARP ? (T.tid & 31) : T.tid;
    const int ept = (len + nt - 1) / nt;                      // <= VB_N4_EPT
    const int lo = start + min(len, tid * ept);
    const int cnt = start + min(len, tid * ept + ept) - lo;   // steps of this thread
    const int mu0 = mu_b, sg0 = sg_b;
    float px[VB_N4_EPT];
    N4Step st[VB_N4_EPT];
    if (cnt == VB_N4_EPT && VB_N4_EPT == 8) {                  // lo is a multiple of 8 here: 16-byte loads
        const float4 p0 = n4_ld_x(c, (const float4*)(epx + lo)), p1 = n4_ld_x(c, (const float4*)(epx + lo + 4));   // (cluster mode: other CTAs wrote parts of it)
        px[0] = p0.x; px[1] = p0.y; px[2] = p0.z; px[3] = p0.w; px[4] = p1.x; px[5] = p1.y; px[6] = p1.z; px[7] = p1.w;
    } else {
#pragma unroll
        for (int i = 0; i < VB_N4_EPT; ++i) px[i] = (i < cnt) ? n4_ld_x(c, epx + lo + i) : 1.0f;
    }
#pragma unroll
    for (int i = 0; i < VB_N4_EPT; ++i) st[i] = n4_step_consts(px[i], lo + i + 1);
    // mu: one scan per sweep.  The starts of a sweep are mu0 + prefix(increments of the previous
    // sweep) (all mu0 for the first one: previous increments 0), so when NO thread's increment
    // differs from its previous one the starts reproduce themselves: thread 0 starts from the exact
    // state, hence its increment is exact, hence thread 1's start, ... -- the trajectory is the
    // sequential one and the sweep that found this is the last (a stalled mean needs one sweep).
    int km = mu0, total = 0, changed = 1, inc_prev = 0;
    bool first = PREDICT;                                     // guessed starts are not a prefix of increments
    if (PREDICT) {
        const float m0 = __int_as_float(mu0);
        float ds = 0.0f;
#pragma unroll
        for (int i = 0; i < VB_N4_EPT; ++i)
            if (i < cnt) ds += px[i] - m0;
        const float ex = n4_guess_scan<WARP>(T, ds, s_guess);
        if (tid > 0 && cnt > 0) km = __float_as_int(m0 + ex / (float)lo);
    }
    for (;;) {
        float mu = __int_as_float(km);
#pragma unroll
        for (int i = 0; i < VB_N4_EPT; ++i)
            if (i < cnt) mu = n4_mu_next(mu, st[i]);
        const int inc = __float_as_int(mu) - km;
        changed = (inc != inc_prev) || first;
        first = false;
        const int off = WARP ? n4_warp_scan(inc, total, changed) : n4_chain_scan(T, inc, total, s_tot[buf], changed);
        buf ^= 1;
        ++c.n_sweeps;
        if (!changed) break;
        inc_prev = inc;
        km = mu0 + off;
    }
    const int mu_end = mu0 + total;
    // sigma, along the exact mu trajectory.  The terms dd*(N-1)/N do not depend on sigma:
    // form them once (head in px[], tail in st[].b).  A step then only depends on sigma
    // through the binade it rounds in, so a thread re-runs its steps only when a new start
    // lies in another binade than the one it ran from (or its span crosses a binade).
    {
        float mu = __int_as_float(km);
#pragma unroll
        for (int i = 0; i < VB_N4_EPT; ++i)
            if (i < cnt) {
                const float d = __fsub_rn(px[i], mu);
                const float dd = __fmul_rn(d, d);
                const float th = fmaf(-dd, st[i].ch, dd);
                const float tl = fmaf(-dd, st[i].cl, fmaf(-dd, st[i].ch, __fsub_rn(dd, th)));
                mu = n4_mu_next(mu, st[i]);
                px[i] = th; st[i].b = tl;
            }
    }
    // first guess: sigma grows about linearly with N inside the chunk; PREDICT: sigma at the
    // thread's start is sg0 + the (float) prefix of the terms, which are known by now
    int ks = __float_as_int(__int_as_float(sg0) * ((float)lo / (float)start));
    if (PREDICT || VB_N4_SIGMA_PREDICT_ALL) {
        float ts = 0.0f;
#pragma unroll
        for (int i = 0; i < VB_N4_EPT; ++i)
            if (i < cnt) ts += px[i];
        ks = __float_as_int(__int_as_float(sg0) + n4_guess_scan<WARP>(T, ts, s_guess));
    }
    if (tid == 0 || cnt == 0) ks = sg0;
    // Same stopping rule as for mu: once no thread's increment differs from the one it had in the
    // previous sweep, the starts (sg0 + prefix of those increments) reproduce themselves.
    int inc = 0;
    bool rerun = true, sfirst = true;
    for (;;) {
        changed = sfirst ? 1 : 0;
        if (rerun) {
            float sg = __int_as_float(ks);
#pragma unroll
            for (int i = 0; i < VB_N4_EPT; ++i)
                if (i < cnt) sg = n4_sigma_add(sg, px[i], st[i].b);
            const int inc_new = __float_as_int(sg) - ks;
            if (inc_new != inc) changed = 1;
            inc = inc_new;
        }
        const int off = WARP ? n4_warp_scan(inc, total, changed) : n4_chain_scan(T, inc, total, s_tot[buf], changed);
        buf ^= 1;
        c.n_sweeps += 1000;
        if (!changed) break;
        sfirst = false;
        const int k = sg0 + off;
        // the increments computed from `ks` also hold from `k` when start and end stay in
        // the same single binade (every add then rounds to the same grid)
        const bool same = (k == ks) || (n4_binade(k) == n4_binade(ks) && n4_binade(k + inc) == n4_binade(k) &&
                                        n4_binade(ks + inc) == n4_binade(ks));
        rerun = !same;
        ks = k;
    }
    mu_b = mu_end;
    sg_b = sg0 + total;
}

template <class CTX> VB_DEV float n4_convergence(CTX& c, const N4Team& T, int M) {
    const int tid = T.tid;
    VB_SHARED int s_state[2];                 // float bit patterns of (mu, sigma) handed between the phases
    VB_SHARED int s_tot[2][32];               // warp totals, double buffered
    int buf = 0;
    const float* epx = c.ec;                  // pixel, C order
    const int S0 = min(M, VB_N4_SERIAL);
    VB_SHARED float s_px[VB_N4_SERIAL];
    long long tq = vb_clock();
    if (tid < S0) s_px[tid] = n4_ld_x(c, epx + tid);
    n4_sync(T);
    // phase 1: the first elements, one thread, sequentially.  Phase 2 (optional, VB_N4_WARP_END >
    // VB_N4_SERIAL): up to VB_N4_WARP_END on one warp, chunks of <= 32 * VB_N4_EPT elements, no barrier.
    VB_SHARED float s_guess[32];
    const int wend = min(M, max(VB_N4_WARP_END, VB_N4_SERIAL));
    if (tid < 32) {
        int mu_b = 0, sg_b = 0;
        if (tid == 0) {
            float mu = 0.0f, sigma = 0.0f;
            for (int j = 0; j < S0; ++j) {
                const float px = s_px[j];
                const N4Step s = n4_step_consts(px, j + 1);
                if (j > 0) sigma = n4_sigma_next(sigma, mu, px, s);
                mu = (j == 0) ? s.b : n4_mu_next(mu, s);
            }
            mu_b = __float_as_int(mu); sg_b = __float_as_int(sigma);
        }
        mu_b = __shfl_sync(VB_FULL, mu_b, 0); sg_b = __shfl_sync(VB_FULL, sg_b, 0);
        { const long long tn = vb_clock(); c.t_sub[0] += tn - tq; tq = tn; }
        for (int start = S0; start < wend;) {
            const int len = min(min(start, 32 * VB_N4_EPT), wend - start);
            n4_chain_chunk<true, false>(c, T, epx, start, len, mu_b, sg_b, s_tot, buf, s_guess);
            start += len;
        }
        if (tid == 0) { s_state[0] = mu_b; s_state[1] = sg_b; }
        { const long long tn = vb_clock(); c.t_sub[1] += tn - tq; tq = tn; }
    }
    n4_sync(T);
    int mu_b = s_state[0], sg_b = s_state[1];
    buf = 0;
    // phase 3: the team, chunks of <= nt * VB_N4_EPT elements, one barrier per sweep.  While the
    // mean still moves (start < VB_N4_PREDICT_END) the starts are predicted and a chunk has the
    // full length; after that a chunk is never longer than the number of elements before it (the
    // start perturbations then contract) and the flat guess (a stalling mean) is the better one.
    // (requesting the elements of chunk k + 1 before chunk k sweeps was measured slower: the eight
    // values do not survive the sweeps in registers, 75.3 -> 77.2 ms per 1776 subjects; a
    // prefetch.global.L1 of them changed nothing, loads that do not allocate in L1 cost 2 %)
    for (int start = max(S0, wend); start < M;) {
        if (start < VB_N4_PREDICT_END) {
            const int len = min(T.nt * VB_N4_EPT, M - start);
            n4_chain_chunk<false, true>(c, T, epx, start, len, mu_b, sg_b, s_tot, buf, s_guess);
            start += len;
        } else {
            const int len = min(min(start, T.nt * VB_N4_EPT), M - start);
            n4_chain_chunk<false, false>(c, T, epx, start, len, mu_b, sg_b, s_tot, buf, s_guess);
            start += len;
        }
    }
    { const long long tn = vb_clock(); c.t_sub[2] += tn - tq; tq = tn; }
    const float mu = __int_as_float(mu_b), sigma = __int_as_float(sg_b);
    n4_sync(T);                               // s_state / s_tot may be rewritten by the next call
    const float sd = (float)sqrt((double)sigma / ((double)(float)M - 1.0));
    return __fdiv_rn(sd, mu);
}

// ---------------------------------------------------------------------------------------------
// P3: evaluate the old and the new bias on the mask: e = exp(old - new) (C order, for the
// convergence measure) and the range of the new U = logI - bias
// The sequential float Welford (n4_convergence) is only run when its outcome can matter.  Its
// running mean lags the exact one by at most u*(j+1)/4 after j steps (u = 2^-23, one rounding of
// at most u/2 per step, damped by (j-1)/j), so by Cauchy-Schwarz its value is at least
//     (sqrt(S) - u*M^1.5/(4*sqrt(3))) / sqrt(M-1) / mean  ~  cv_exact - u*M/(4*sqrt(3))
// (S = exact sum of squared deviations).  If that bound, with 2 % to spare for the float
// rounding of sigma and mean, still exceeds the threshold the level certainly continues and the
// exact double-precision cv is returned instead (`exact_needed` forces the chain: last allowed
// iteration of a level, debug capture).
// Returns true when the sequential chain has to be run on c.ec; otherwise `conv` holds the
// (exact-arithmetic) value and the level certainly continues.
// The update pass proper.  TAB (the usual case: both fields come from per-column coefficient tables):
// a FLAT loop over the compact arrays -- thread i takes mask voxel i, i + nt, ... (C order), whose row
// and active-column index come from the per-voxel code written by n4_setup; the four collapsed
// coefficients of each field are read from the tables (lanes of a warp sit on consecutive columns of
// at most a few rows: conflict-free or broadcast).  No row walk, no ballots, every lane busy: 2.3x
// fewer warp instructions than walking the columns with a sliding window (lane utilisation of the
// walk is 73 % on the benchmark masks, and its row bookkeeping costs more than the 8 extra loads).
// !TAB (tables do not fit, or the volume is too large for the code packing): the column walk.
// s1 / s2: sums of (e - 1) and (e - 1)^2 for the skip bound of n4_update; at most 64 terms are added
// in float before they move to the double accumulators (relative error of a partial <= 64 * 2^-24,
// accounted for in the bound).
// SM: U and the row tables live in shared memory too (the benchmark volumes): every table access of
// the loop is an LDS / STS through a 32-bit shared address.
template <bool SM, class CTX>
VB_DEV void n4_update_flat(CTX& c, const N4Team& T, const float* Gn, float& tmin, float& tmax, double& s1, double& s2) {
    const float* logI = c.logI;
    float* ec = c.ec;
    const uint32_t* code = c.code;
    VB_OPAQUE_PTR(logI); VB_OPAQUE_PTR(ec); VB_OPAQUE_PTR(code);
    float* U = c.U;
    const float4* revw = c.r_evw;
    const int* rbev = c.r_bev;
    if (!SM) { VB_OPAQUE_PTR(U); VB_OPAQUE_PTR(revw); VB_OPAQUE_PTR(rbev); }
    const vb_saddr sU = SM ? vb_saddr_of(c.U) : 0, sW = SM ? vb_saddr_of(c.r_evw) : 0, sB = SM ? vb_saddr_of(c.r_bev) : 0;
    const vb_saddr sGo = vb_saddr_of(c.G), sGn = vb_saddr_of(Gn);
    // (cluster mode: this CTA's slice [M0, M) of the compact arrays)
    const int M0 = (int)((long long)c.M * c.rank() / c.k()), M = (int)((long long)c.M * (c.rank() + 1) / c.k()), nt = T.nt;
    const unsigned gs4 = 4u * (unsigned)c.gstride;
    // VB_N4_UB voxels per trip: their codes and log-image values (L2) are requested together, before
    // the first of them is evaluated -- the stores of the loop body would otherwise keep the compiler
    // from moving the next loads up, and a thread would wait one L2 round trip per voxel
    for (int i0 = M0 + T.tid; i0 < M; i0 += 64 * nt) {
        float f1 = 0.0f, f2 = 0.0f;
        const int iend = min(M, i0 + 64 * nt);
        for (int ib = i0; ib < iend; ib += VB_N4_UB * nt) {
            uint32_t cd[VB_N4_UB];
            float li[VB_N4_UB];
#pragma unroll
            for (int u = 0; u < VB_N4_UB; ++u) {
                const int i = ib + u * nt;
                cd[u] = (i < iend) ? VB_N4_LDS(code + i) : 0u;
                li[u] = (i < iend) ? VB_N4_LDS(logI + i) : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < VB_N4_UB; ++u) {
                const int i = ib + u * nt;
                if (i < iend) {
                    const unsigned r = cd[u] >> 16, k = cd[u] & 0xffffu;
                    const int base = SM ? vb_lds_s32(sB + 4u * r) : rbev[r];
                    const float4 rw = SM ? vb_lds_f32x4(sW + 16u * r) : revw[r];
                    const unsigned at = (unsigned)base * gs4 + 4u * k;
                    const vb_saddr ao = sGo + at, an = sGn + at;
                    N4Win wo, wn;
                    wo.g0 = vb_lds_f32(ao); wo.g1 = vb_lds_f32(ao + gs4); wo.g2 = vb_lds_f32(ao + 2u * gs4); wo.g3 = vb_lds_f32(ao + 3u * gs4);
                    wn.g0 = vb_lds_f32(an); wn.g1 = vb_lds_f32(an + gs4); wn.g2 = vb_lds_f32(an + 2u * gs4); wn.g3 = vb_lds_f32(an + 3u * gs4);
                    const float bo = n4_win_eval(wo, rw), bn = n4_win_eval(wn, rw);
                    const float un = li[u] - bn;
                    if (SM) vb_sts_f32(sU + 4u * (unsigned)i, un); else U[i] = un;   // read by the next histogram / fit
                    tmin = fminf(tmin, un);
                    tmax = fmaxf(tmax, un);
                    const float ev = n4_exp_small(bo - bn);
                    VB_N4_STS(ec + i, ev);                               // chain input (C order = compact order)
                    const float em = ev - 1.0f;                          // exact (Sterbenz) for 0.5 <= ev <= 2
                    f1 += em; f2 = fmaf(em, em, f2);
                }
            }
        }
        s1 += (double)f1; s2 += (double)f2;
    }
}

template <class CTX> VB_DEV void n4_update_walk(CTX& c, const N4Team& T, float& tmin, float& tmax, double& s1, double& s2) {
    const float* logI = c.logI;
    float* ec = c.ec;
    float* U = c.U;
    const float4* revw = c.r_evw;
    const int* rbev = c.r_bev;
    VB_OPAQUE_PTR(logI); VB_OPAQUE_PTR(ec); VB_OPAQUE_PTR(U); VB_OPAQUE_PTR(revw); VB_OPAQUE_PTR(rbev);
    const N4Axis ac = n4_ax(c, 1), as = n4_ax(c, 2);
    const N4Walk wk = n4_walk(c);
    VB_N4_FOR_ITEMS(c, T, it) {
        const N4Item I = n4_item(c, it);
        if (I.empty) continue;
        const N4Col col = n4_col(c, ac, as, I.q);
        N4Win wo, wn;
        wo.cur = wn.cur = -1;
        wo.g0 = wo.g1 = wo.g2 = wo.g3 = wn.g0 = wn.g1 = wn.g2 = wn.g3 = 0.0f;
        float f1 = 0.0f, f2 = 0.0f;
        int nf = 0;
        VB_N4_FOR_ROWS(wk, I, {
            const int base = rbev[r];
            float li = 0.0f;
            if (bit) li = logI[idx];
            n4_win_advance(wo, c.Lold, c, col, base);
            n4_win_advance(wn, c.L, c, col, base);
            if (bit) {
                const float4 rw = revw[r];
                const float bo = n4_win_eval(wo, rw), bn = n4_win_eval(wn, rw);
                const float un = li - bn;
                U[idx] = un;
                tmin = fminf(tmin, un);
                tmax = fmaxf(tmax, un);
                const float ev = n4_exp_small(bo - bn);
                ec[idx] = ev;
                const float em = ev - 1.0f;
                f1 += em; f2 = fmaf(em, em, f2);
            }
            if (++nf == 64) { s1 += (double)f1; s2 += (double)f2; f1 = f2 = 0.0f; nf = 0; }     // warp uniform
        })
        s1 += (double)f1; s2 += (double)f2;
    }
}

template <class CTX> VB_DEV bool n4_update(CTX& c, int M, float& bmin, float& bmax, float thr, bool exact_needed, float& conv) {
    const int nt = (int)blockDim.x;
    float tmin = FLT_MAX, tmax = -FLT_MAX;
    double s1 = 0.0, s2 = 0.0;
    float* Gn = c.G ? (c.G2 ? c.G2 : c.un_f) : nullptr;  // table of the new lattice (spare table, else the union region: free here)
    N4Team T;
    T.tid = (int)threadIdx.x; T.nt = nt; T.bar = 0;
    c.deal(c.k(), c.rank());                     // (cluster mode: the column walk's items are dealt over the CTAs)
    n4_items_begin(c, T);
    if (Gn) n4_build_G(c, c.L, Gn); else __syncthreads();
    if (Gn) {
        if (c.u_smem && c.rowtab) n4_update_flat<true>(c, T, Gn, tmin, tmax, s1, s2);
        else n4_update_flat<false>(c, T, Gn, tmin, tmax, s1, s2);
    } else {
        n4_update_walk(c, T, tmin, tmax, s1, s2);
    }
#if !VB_N4_MERGED_RED
    bmin = vb_block_reduce(tmin, VbMinF(), FLT_MAX, (float*)c.red);
    bmax = vb_block_reduce(tmax, VbMaxF(), -FLT_MAX, (float*)c.red);
    s1 = vb_block_reduce(s1, VbAddD(), 0.0, c.red);
    s2 = vb_block_reduce(s2, VbAddD(), 0.0, c.red);
#else
    {   // the four block reductions of vb_block_reduce (same trees, same results) behind ONE pair of barriers
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            tmin = fminf(tmin, __shfl_xor_sync(VB_FULL, tmin, m));
            tmax = fmaxf(tmax, __shfl_xor_sync(VB_FULL, tmax, m));
            s1 += __shfl_xor_sync(VB_FULL, s1, m);
            s2 += __shfl_xor_sync(VB_FULL, s2, m);
        }
        const int nw = nt >> 5, lane = vb_lane(), w = vb_warp();
        float* redf = (float*)(c.red + 64);
        __syncthreads();
        if (lane == 0) { c.red[w] = s1; c.red[32 + w] = s2; redf[w] = tmin; redf[32 + w] = tmax; }
        __syncthreads();
        s1 = (lane < nw) ? c.red[lane] : 0.0;
        s2 = (lane < nw) ? c.red[32 + lane] : 0.0;
        tmin = (lane < nw) ? redf[lane] : FLT_MAX;
        tmax = (lane < nw) ? redf[32 + lane] : -FLT_MAX;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            tmin = fminf(tmin, __shfl_xor_sync(VB_FULL, tmin, m));
            tmax = fmaxf(tmax, __shfl_xor_sync(VB_FULL, tmax, m));
            s1 += __shfl_xor_sync(VB_FULL, s1, m);
            s2 += __shfl_xor_sync(VB_FULL, s2, m);
        }
        bmin = tmin; bmax = tmax;
    }
#endif
    c.deal(1, 0);
    n4_cluster_reduce4(c, T, bmin, bmax, s1, s2);        // (cluster mode; also orders the U / sequence writes of all CTAs)
    __syncthreads();
    if (c.G2) {                                          // the new lattice's table becomes the current one
        c.G2 = c.G; c.G = Gn;
    } else if (Gn) {
        for (int e = (int)threadIdx.x; e < c.ncp[0] * c.gstride; e += (int)blockDim.x) c.G[e] = Gn[e];
        __syncthreads();
    }
    if (!exact_needed) {
        const double N = (double)M;
        // the float partials of s1 / s2 carry a relative error <= 64 * 2^-24 each; |s1| <= sqrt(N s2), so
        // (N - 1) var is known to within 3 * 4e-6 * s2: the bound uses the low end
        double var = (s2 * (1.0 - 1.2e-5) - s1 * s1 / N) / (N - 1.0);
        if (var < 0.0) var = 0.0;
        const double mean = 1.0 + s1 / N;
        const double lag = 1.1920929e-7 * N / (4.0 * 1.7320508) * 1.05;
        const double lower = (sqrt(var) - lag) / mean * 0.98;
        if (lower > (double)thr) { conv = (float)(sqrt(var) / mean); return false; }
    }
    return true;
}

// L <- L + Phi, keeping the previous lattice for the next update pass
template <class CTX> VB_DEV void n4_commit(CTX& c) {
    const int Ln = c.ncp[0] * c.ncp[1] * c.ncp[2];
    for (int e = (int)threadIdx.x; e < Ln; e += (int)blockDim.x) {
        const float l = c.L[e];
        c.Lold[e] = l;
        c.L[e] = l + c.Phi[e];
    }
    __syncthreads();
}

// final dense pass: bias at EVERY voxel, out = in / exp(bias)   (N4 GenerateData tail)
template <class CTX> VB_DEV void n4_finalize(CTX& c, float* logbias) {
    const int nt = (int)blockDim.x, lane = vb_lane(), w = vb_warp(), nw = nt >> 5;
    const int ncols = c.ncols;
    const N4Axis ac = n4_ax(c, 1), as = n4_ax(c, 2);
    const int nblk_all = (ncols + 31) >> 5;              // every column, not only the active ones
    for (int blk = w + nw * c.rank(); blk < nblk_all; blk += nw * c.k()) {     // (cluster mode: column blocks dealt over the CTAs)
        const int q = blk * 32 + lane;
        if (q >= ncols) continue;
        const N4Col col = n4_col(c, ac, as, q);
        N4Win win; win.cur = -1; win.g0 = win.g1 = win.g2 = win.g3 = 0.0f;
        for (int r0 = 0; r0 < c.n0; r0 += 8) {
            float x[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                x[j] = (r0 + j < c.n0) ? vb_ld_stream(c.img + (size_t)(r0 + j) * ncols + q) : 0.0f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int r = r0 + j;
                if (r < c.n0) {
                    n4_win_advance(win, c.L, c, col, c.r_bev[r]);
                    const float bias = n4_win_eval(win, c.r_evw[r]);
                    const size_t v = (size_t)r * ncols + q;
                    vb_st_stream(c.out + v, __fdiv_rn(x[j], expf(bias)));
                    if (logbias) vb_st_stream(logbias + v, bias);
                }
            }
        }
    }
    __syncthreads();
}

// BSplineControlPointImageFilter::RefineControlPointLattice (every axis n -> 2n-3)
template <class CTX> VB_DEV void n4_refine(CTX& c) {
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

template <class CTX> VB_DEV void n4_build_tables(CTX& c) {
    const int n[3] = {c.n0, c.n1, c.n2};
    size_t off = 0;
    for (int a = 0; a < 3; ++a) {
        n4_build_axis(c.axis_base + off, c.slot[a], n[a], c.ncp[a]);
        off += 16 * (size_t)c.slot[a];
    }
    __syncthreads();
    off = 0;
    for (int a = 0; a < 3; ++a) {
        n4_build_support(c.axis_base + off, c.slot[a], n[a], c.ncp[a]);
        off += 16 * (size_t)c.slot[a];
    }
    const N4Axis ar = n4_ax(c, 0);
    if (c.rowtab) {
        float4* evw = (float4*)c.rowtab;
        float4* fit3 = evw + c.n0;
        int* bev = (int*)(fit3 + c.n0);
        int* bfit = bev + c.n0;
        for (int r = (int)threadIdx.x; r < c.n0; r += (int)blockDim.x) {
            evw[r] = ar.evw[r]; fit3[r] = ar.fit3[r]; bev[r] = ar.bev[r]; bfit[r] = ar.bfit[r];
        }
        c.r_evw = evw; c.r_fit3 = fit3; c.r_bev = bev; c.r_bfit = bfit;
    } else {
        c.r_evw = ar.evw; c.r_fit3 = ar.fit3; c.r_bev = ar.bev; c.r_bfit = ar.bfit;
    }
    __syncthreads();
}

// CL = false: one persistent CTA per subject (c.k() is the literal 1: the cluster-mode code and state
// fold away); CL = true: a cluster of a.K CTAs per subject.
template <bool CL>
VB_DEV void n4_subject_body(const N4Args& a) {
    VB_DYN_SMEM(smem);
    VB_SHARED int s_next;
    VB_SHARED float s_conv;
    VB_SHARED int s_ctr[4];
    const int tid = (int)threadIdx.x;
    N4Team full, TC, TP;
    full.tid = tid; full.nt = (int)blockDim.x; full.bar = 0;
    const int chain_warps = (CL && a.cl_chain_warps > 0) ? a.cl_chain_warps : VB_N4_CHAIN_WARPS;
    TC.tid = tid; TC.nt = min(chain_warps * 32, (int)blockDim.x - 32); TC.bar = 1;
    TP.tid = tid - TC.nt; TP.nt = (int)blockDim.x - TC.nt; TP.bar = 2;
    N4CtxT<CL> c;
    c.un_f = (float*)smem;
    c.un_d = (double*)smem;
    unsigned char* p = smem + a.union_bytes;
    const size_t Lb = sizeof(float) * (size_t)((a.Lmax + 3) & ~3);
    c.L = (float*)p;    p += Lb;
    c.Lold = (float*)p; p += Lb;
    c.Om = (float*)p;   p += Lb;
    c.Phi = (float*)p;  p += Lb;
    c.E = (float*)p;    p += sizeof(float) * (size_t)((a.p.n_bins + 3) & ~3);
    c.H = (float*)p;    p += sizeof(float) * (size_t)((a.p.n_bins + 3) & ~3);
    c.red = (double*)p; p += VB_N4_RED * sizeof(double);
    c.tw = (double*)p;  p += sizeof(double) * (size_t)a.fftN;
    float* smem_rowtab = (float*)p; p += (size_t)40 * a.rowtab_rows;
    float* smem_logI = (float*)p;
    c.nbins = a.p.n_bins; c.fftN = a.fftN; c.union_bytes = a.union_bytes;
    n4_twiddles(c);

    // cluster mode: cluster `cid` = CTAs [cid * K, (cid + 1) * K) of the grid works on subject cid; its CTAs
    // share workspace slot cid (the per-axis tables are per CTA: every CTA builds its own)
    c.K = CL ? max(a.K, 1) : 1; c.crank = (c.K > 1) ? (int)(blockIdx.x % (unsigned)c.K) : 0; c.istep = 1; c.ioff = 0;
    const int cid = (c.k() > 1) ? (int)(blockIdx.x / (unsigned)c.k()) : (int)blockIdx.x;
    if (CL) {
        c.q_max = a.q_max;
        c.cl_hist = a.cl_hist ? a.cl_hist + (size_t)cid * 4 * c.k() * a.p.n_bins : nullptr;
        c.cl_q = a.cl_q ? a.cl_q + (size_t)cid * 4 * c.k() * 2 * a.q_max : nullptr;
        c.cl_red = a.cl_red ? a.cl_red + (size_t)cid * 2 * c.k() * 4 : nullptr;
        c.cl_bar = a.cl_bar ? a.cl_bar + (size_t)cid * 4 : nullptr;
        c.cl_bcast = a.cl_bcast ? a.cl_bcast + (size_t)cid * 16 : nullptr;
        c.gen[0] = c.gen[1] = c.gen[2] = 0u; c.par_hist0 = c.par_hist1 = c.par_q0 = c.par_q1 = c.par_red = 0;
    }
    const size_t slotx = (size_t)cid;

    for (int round = 0;; ++round) {
        __syncthreads();
        if (c.k() == 1) {
            if (tid == 0) s_next = atomicAdd(a.work_counter, 1);
            __syncthreads();
        }
        const int si = (c.k() == 1) ? s_next : cid;
        if (si >= a.n || (c.k() > 1 && round > 0)) break;
        const VbSubject S = a.subj[si];
        c.n0 = S.n0; c.n1 = S.n1; c.n2 = S.n2; c.ncols = S.n1 * S.n2; c.W = (S.n0 + 31) >> 5;
        c.nblk = 0; c.nact = 0; c.nseg = 1; c.wps = c.W;
        c.V = S.V;
        c.img = a.image + S.off; c.msk = a.mask + S.off; c.out = a.out + S.off;
        c.colmask = a.ws_colmask + slotx * a.cmw_max;
        c.rbb = a.ws_rbb + slotx * a.rbb_max;
        c.wbase = a.ws_wbase + slotx * a.wb_max;
        c.acol = a.ws_acol + slotx * a.ncols_max;
        c.iorder = a.ws_order + slotx * a.wb_max;
        c.ctr = s_ctr;
        c.axis_base = a.ws_axis + (size_t)blockIdx.x * a.axis_stride;
        c.ec = a.ws_ec + slotx * (size_t)((a.Vmax + 3) & ~3LL);
        c.code = (S.n0 <= 65536 && S.n1 * S.n2 <= 65536) ? a.ws_code + slotx * a.Vmax : nullptr;
        c.ncp[0] = a.p.ncp[0]; c.ncp[1] = a.p.ncp[1]; c.ncp[2] = a.p.ncp[2];
        {
            const int nn[3] = {S.n0, S.n1, S.n2};
            for (int ax = 0; ax < 3; ++ax) {
                int f = a.p.ncp[ax];
                for (int l = 1; l < a.p.n_levels; ++l) f = 2 * f - 3;
                c.slot[ax] = max(nn[ax], f);
            }
        }
        c.logI = a.ws_logI + slotx * a.Vmax;
        c.U = a.ws_u + slotx * a.Vmax;                       // n4_setup moves it to shared memory if it fits
        c.rowtab = (a.rowtab_rows > 0 && S.n0 <= a.rowtab_rows) ? smem_rowtab : nullptr;

        float bmin, bmax;
        int M;
        long long tph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        long long tq = vb_clock();
        c.t_chain = 0; c.n_sweeps = 0; c.t_sub[0] = c.t_sub[1] = c.t_sub[2] = c.t_sub[3] = 0;
        if (c.k() == 1) {
            n4_setup(c, a.p.mask_label, smem_logI, a.logI_cap, bmin, bmax, M);
        } else {
            // CTA 0 of the cluster compacts the volume (in-place scans of the shared tables: one CTA) and
            // hands the scalars to the others
            if (c.rank() == 0) {
                n4_setup(c, a.p.mask_label, smem_logI, 0, bmin, bmax, M);
                if (tid == 0) {
                    int* b = c.cl_bcast;
                    b[0] = M; b[1] = c.nact; b[2] = c.nblk; b[3] = c.nseg; b[4] = c.wps; b[5] = c.hshift;
                    b[6] = __float_as_int(bmin); b[7] = __float_as_int(bmax);
                }
            }
            n4_csync(c, full);
            M = vb_ld_cg(c.cl_bcast + 0); c.nact = vb_ld_cg(c.cl_bcast + 1); c.nblk = vb_ld_cg(c.cl_bcast + 2);
            c.nseg = vb_ld_cg(c.cl_bcast + 3); c.wps = vb_ld_cg(c.cl_bcast + 4); c.hshift = vb_ld_cg(c.cl_bcast + 5);
            bmin = __int_as_float(vb_ld_cg(c.cl_bcast + 6)); bmax = __int_as_float(vb_ld_cg(c.cl_bcast + 7));
            c.u_smem = false; c.U = a.ws_u + slotx * a.Vmax;
        }
        c.M = M;
        for (int e = tid; e < a.Lmax; e += (int)blockDim.x) { c.L[e] = 0.0f; c.Lold[e] = 0.0f; }
        {
            const int used = (c.U == smem_logI) ? ((M + 3) & ~3) : 0;
            c.gspace = smem_logI + used; c.gcap = a.logI_cap - used;
            c.gstride = (c.nact + 31) & ~31; c.G = nullptr; c.G2 = nullptr;
        }
        __syncthreads();
        tph[0] = vb_clock() - tq;

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
                {
                    const int need = c.ncp[0] * c.gstride;
                    c.G2 = (VB_N4_G2 && 2 * need <= c.gcap) ? c.gspace + need : nullptr;
                    c.G = (c.code && need <= c.gcap && (c.G2 || need * 4 <= a.union_bytes)) ? c.gspace : nullptr;
                    if (!c.G) c.G2 = nullptr;
                    if (c.G) n4_build_G(c, c.L, c.G);
                }
                n4_contract<true>(c, full, bmin, bmax);
                if (level > 0) n4_range(c, bmin, bmax);
                conv = FLT_MAX;
                // `pending`: c.ec holds iteration `it`'s exp(old - new) and the sequential
                // convergence chain has not been run on it yet.  Outside debug capture the chain
                // runs on the chain team WHILE the phase team already computes histogram,
                // sharpening and fit of iteration it + 1; the fit is only committed if the chain
                // says "continue", so the result is that of the sequential schedule.
                bool pending = false;
                const bool serial = dbg != nullptr || a.no_overlap == 1;
                for (;;) {
                    const bool room = it < a.p.max_iter[level];
                    const bool cap = a.p.max_total_iters > 0 && total >= a.p.max_total_iters;
                    bool fitted = false;
                    if (pending) {
                        tq = vb_clock();
                        if (serial) {                                // no overlap: keeps the captured stages in order
                            const float cv = n4_convergence(c, full, M);
                            if (tid == 0) s_conv = cv;
                        } else if (tid < TC.nt) {
                            const float cv = n4_convergence(c, TC, M);
                            if (tid == 0) s_conv = cv;
                        } else if (room && !cap) {
                            n4_histogram(c, TP, bmin, bmax);
                            n4_sharpen(c, TP, bmin, bmax, a.p.fwhm, a.p.wiener_noise);
                            n4_contract<false>(c, TP, bmin, bmax);
                        }
                        n4_csync(c, full);               // (cluster mode: every CTA's chain has read the sequence before the next update rewrites it)
                        conv = s_conv;
                        __syncthreads();
                        fitted = !serial && room && !cap;
                        pending = false;
                        tph[5] += vb_clock() - tq;
                        if (dbg_trace && tid == 0 && ntrace - 1 < trace_cap && ntrace > 0) dbg_trace[ntrace - 1] = conv;
                    }
                    if (!(room && conv > a.p.conv_threshold)) break;
                    if (cap) { capped = true; break; }
                    if (!fitted) {
                        if (dbg && tid == 0) { dbg[0] = bmin; dbg[1] = bmax; }
                        tq = vb_clock();
                        n4_histogram(c, full, bmin, bmax);
                        long long tn = vb_clock(); tph[1] += tn - tq; tq = tn;
                        n4_sharpen(c, full, bmin, bmax, a.p.fwhm, a.p.wiener_noise);
                        tn = vb_clock(); tph[2] += tn - tq; tq = tn;
                        n4_contract<false>(c, full, bmin, bmax);
                        tn = vb_clock(); tph[3] += tn - tq;
                    }
                    n4_commit(c);
                    ++it; ++total;
                    tq = vb_clock();
                    pending = n4_update(c, M, bmin, bmax, a.p.conv_threshold, dbg != nullptr || it >= a.p.max_iter[level], conv);
                    tph[4] += vb_clock() - tq;
                    if (!pending && dbg_trace && tid == 0 && ntrace < trace_cap) dbg_trace[ntrace] = conv;
                    ++ntrace;
                }
            }
            if (a.stats && tid == 0 && c.rank() == 0) a.stats[si].n4_iters[level] = it;
            if (level != a.p.n_levels - 1 && M >= 2 && !capped) n4_refine(c);
        }
        if (M < 2 || capped) n4_build_tables(c);   // tables of the final lattice (M < 2: zero lattice => out = in)
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
        tq = vb_clock();
        if (a.stats && tid == 0 && c.rank() == 0) {
            for (int l = a.p.n_levels; l < VB_MAX_LEVELS; ++l) a.stats[si].n4_iters[l] = 0;
            a.stats[si].n4_last_conv = conv;
            a.stats[si].n4_n_mask = M;
        }
        n4_finalize(c, a.logbias ? a.logbias + S.off : nullptr);
        if (dbg && tid == 0) {
            tph[6] = vb_clock() - tq;
            for (int k = 0; k < 8; ++k) dbg[8 + k] = (float)((double)tph[k] * (1.0 / 1965.0));   // cycles -> us at 1965 MHz
            dbg[15] = (float)c.n_sweeps;
            for (int k = 0; k < 4; ++k) dbg[16 + k] = (float)((double)c.t_sub[k] * (1.0 / 1965.0));
        }
    }
}

VB_KERNEL void VB_LAUNCH_BOUNDS(VB_N4_THREADS, 1) n4_subject_kernel(N4Args a) { n4_subject_body<false>(a); }
VB_KERNEL void VB_LAUNCH_BOUNDS(VB_N4_THREADS, 1) n4_cluster_kernel(N4Args a) { n4_subject_body<true>(a); }
