// Overlay compositions that consume the path's arrays:
//   Vent_Analysis.exportDICOM   Vent_Analysis.py:387-391   BW = uint8(normalize(|N4|) * 255), defects red
//   GUI colorBinary             Vent_Analysis.py:628-634   the same in float64, on montages of whole volumes
//   Vent_Analysis.cropToData    Vent_Analysis.py:430-456   box of the rows / cols / slices with signal
//   Vent_Analysis.screenShot    Vent_Analysis.py:466-494   7 x slices montage, parula CI colours
// (CPU restatement: oracle/overlay_oracle.py.)  All of it is streaming work: a min / max reduction
// per subject, then one pass that reads 5-16 bytes per voxel and writes 3 bytes per pixel.
#pragma once
#include <float.h>
#include "vb_common.cuh"
#include "ci_kernels.cuh"

#define VB_OVL_F32 0      /* normalisation in float32 arithmetic (float32 arrays)                     */
#define VB_OVL_F64 1      /* ... in float64 (integer or float64 arrays; values held exactly as float32) */

struct OvlStats {
    unsigned kmin[3], kmax[3];   // order-preserving keys of min / max of {n4, image, proton} (|.| for the RGB overlays)
    int nan[3];                  // a NaN was seen: numpy's min / max return NaN
    int box[6];                  // r0, r1, c0, c1, s0, s1 (crop box; whole volume for the RGB overlays)
    int error;                   // 1: cropToData found no index > 0 with signal (reference: IndexError); 2: CI colour index >= 64
};

// numpy's C cast of a float / double to uint8 on x86-64: truncate to int32 (0x80000000 when out of
// range or NaN), keep the low byte
VB_DEV unsigned vb_cast_u8(double v) {
    if (!(fabs(v) < 2147483648.0)) return 0u;
    return (unsigned)((int)v) & 0xFFu;
}

// (x - mn) / (mx - mn) as numpy evaluates normalize() for the array's dtype; `flat`: max == min,
// normalize() returns x itself
VB_DEV double ovl_norm(float x, float mn, float mx, bool flat, int mode) {
    if (flat) return (double)x;
    if (mode == VB_OVL_F32) return (double)__fdiv_rn(__fsub_rn(x, mn), __fsub_rn(mx, mn));
    return __ddiv_rn(__dsub_rn((double)x, (double)mn), __dsub_rn((double)mx, (double)mn));
}

VB_KERNEL void ovl_init_kernel(OvlStats* st, const VbSubject* subj, int n) {
    const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= n) return;
    for (int a = 0; a < 3; ++a) { st[i].kmin[a] = 0xffffffffu; st[i].kmax[a] = 0u; st[i].nan[a] = 0; }
    st[i].box[0] = 0; st[i].box[1] = subj[i].n0; st[i].box[2] = 0; st[i].box[3] = subj[i].n1; st[i].box[4] = 0; st[i].box[5] = subj[i].n2;
    st[i].error = 0;
}

// cropToData(mask, border): one CTA per subject.  An index counts as "has signal" when the sum of
// its plane is > 0 (any non-zero voxel for the non-negative masks used here) AND the index is > 0
// (np.multiply(flag, range) is 0 at index 0, Vent_Analysis.py:433-440).
VB_KERNEL void ovl_cropbox_kernel(const uint8_t* VB_RESTRICT mask, const VbSubject* subj, int border, int border_slices, OvlStats* st) {
    VB_SHARED int lo[3], hi[3];
    const int si = (int)blockIdx.x, tid = (int)threadIdx.x, nt = (int)blockDim.x;
    const VbSubject S = subj[si];
    if (tid < 3) { lo[tid] = 0x7fffffff; hi[tid] = -1; }
    __syncthreads();
    const int ncols = S.n1 * S.n2;
    int l0 = 0x7fffffff, h0 = -1, l1 = 0x7fffffff, h1 = -1, l2 = 0x7fffffff, h2 = -1;
    for (int v = tid; v < S.V; v += nt) {
        if (mask[S.off + v]) {
            const int i = v / ncols, rem = v - i * ncols, j = rem / S.n2, k = rem - j * S.n2;
            if (i > 0) { l0 = min(l0, i); h0 = max(h0, i); }
            if (j > 0) { l1 = min(l1, j); h1 = max(h1, j); }
            if (k > 0) { l2 = min(l2, k); h2 = max(h2, k); }
        }
    }
    atomicMin(&lo[0], l0); atomicMax(&hi[0], h0); atomicMin(&lo[1], l1); atomicMax(&hi[1], h1);
    atomicMin(&lo[2], l2); atomicMax(&hi[2], h2);
    __syncthreads();
    if (tid == 0) {
        OvlStats* o = st + si;
        if (hi[0] < 0 || hi[1] < 0 || hi[2] < 0) { o->error = 1; return; }
        const int sb = border_slices ? border : 0;
        o->box[0] = max(lo[0] - border, 0); o->box[1] = min(hi[0] + border + 1, S.n0);
        o->box[2] = max(lo[1] - border, 0); o->box[3] = min(hi[1] + border + 1, S.n1);
        o->box[4] = max(lo[2] - sb, 0);     o->box[5] = min(hi[2] + sb + 1, S.n2);
    }
}

// min / max of up to three float32 arrays over the subject's box.  grid = (chunks, subjects).
// ABS: of |x| (exportDICOM / colorBinary take np.abs first).
struct OvlMinMaxArgs {
    const float* a[3];            // null entries are skipped
    const VbSubject* subj;
    OvlStats* st;
    int use_abs;
};
VB_KERNEL void ovl_minmax_kernel(OvlMinMaxArgs p) {
    VB_SHARED int red[64];
    const int si = (int)blockIdx.y;
    const VbSubject S = p.subj[si];
    OvlStats* o = p.st + si;
    if (o->error) return;
    const int r0 = o->box[0], c0 = o->box[2], s0 = o->box[4];
    const int R = o->box[1] - r0, Cc = o->box[3] - c0, Sl = o->box[5] - s0;
    const int nbox = R * Cc * Sl;
    for (int a = 0; a < 3; ++a) {
        if (!p.a[a]) continue;
        const float* x = p.a[a] + S.off;
        unsigned kmin = 0xffffffffu, kmax = 0u;
        int nan = 0;
        const bool whole = nbox == S.V && (S.off & 3) == 0;
        if (whole) {
            // the whole volume: a flat loop, four independent 16-byte loads per thread in flight
            const int nq = S.V >> 2, nth = (int)(gridDim.x * blockDim.x), t0 = (int)(blockIdx.x * blockDim.x + threadIdx.x);
            float fmn = FLT_MAX, fmx = -FLT_MAX;                     // (fminf / fmaxf skip NaNs: they are flagged separately)
            for (int q = t0; q < nq; q += 4 * nth) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (q + u * nth < nq) v[u] = *(const float4*)(x + 4 * (size_t)(q + u * nth));
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (q + u * nth < nq) {
                        float f[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float w = p.use_abs ? fabsf(f[e]) : f[e];
                            if (w != w) nan = 1;
                            fmn = fminf(fmn, w); fmx = fmaxf(fmx, w);
                        }
                    }
            }
            for (int e = 4 * nq + t0; e < S.V; e += nth) {
                const float w = p.use_abs ? fabsf(x[e]) : x[e];
                if (w != w) nan = 1;
                fmn = fminf(fmn, w); fmx = fmaxf(fmx, w);
            }
            if (fmn <= fmx) { kmin = vb_float_key(fmn); kmax = vb_float_key(fmx); }
        }
        for (int e = (int)(blockIdx.x * blockDim.x + threadIdx.x); e < (whole ? 0 : nbox); e += (int)(gridDim.x * blockDim.x)) {
            const int k = e % Sl, t = e / Sl, j = t % Cc, i = t / Cc;
            float v = x[((size_t)(r0 + i) * S.n1 + (c0 + j)) * S.n2 + (s0 + k)];
            if (p.use_abs) v = fabsf(v);
            if (v != v) { nan = 1; continue; }
            const unsigned key = vb_float_key(v);
            kmin = min(kmin, key); kmax = max(kmax, key);
        }
        const unsigned lo = (unsigned)vb_block_reduce((int)(kmin ^ 0x80000000u), VbMinI(), 0x7fffffff, red) ^ 0x80000000u;
        const unsigned hi = (unsigned)vb_block_reduce((int)(kmax ^ 0x80000000u), VbMaxI(), (int)0x80000000, red) ^ 0x80000000u;
        nan = vb_block_reduce(nan, VbMaxI(), 0, red);
        if (threadIdx.x == 0) {
            atomicMin(&o->kmin[a], lo); atomicMax(&o->kmax[a], hi);
            if (nan) atomicMax(&o->nan[a], 1);
        }
    }
}

// RGB overlay of a 0/1 array B on the normalised |A|, interleaved uint8 [r][c][s][3].
//   gui == 0  exportDICOM: BW = uint8(float32(norm * 255)); R = B==1 ? 255 : (B==0 ? BW : 0); G = B = (B==0 ? BW : 0)
//   gui == 1  colorBinary + astype(uint8): float64; R = uint8((norm*(B==0) + B) * 255), G = B = uint8(norm*(B==0) * 255)
struct OvlRgbArgs {
    const float* a;
    const uint8_t* b;
    const VbSubject* subj;
    const OvlStats* st;
    uint8_t* rgb;                 // [3 * V_total]
    int gui, mode;
};
VB_DEV void ovl_rgb_one(const OvlRgbArgs& p, float x, unsigned b, float mn, float mx, bool flat, unsigned& r, unsigned& gb) {
    const double nv = ovl_norm(x, mn, mx, flat, p.mode);
    if (!p.gui) {
        const unsigned bw = (p.mode == VB_OVL_F32) ? vb_cast_u8((double)__fmul_rn((float)nv, 255.0f)) : vb_cast_u8(__dmul_rn(nv, 255.0));
        gb = (b == 0u) ? bw : 0u;
        r = (b == 1u) ? 255u : gb;
    } else {
        const double keep = (b == 0u) ? nv : __dmul_rn(nv, 0.0);      // A*(B==0): NaN stays NaN
        gb = vb_cast_u8(__dmul_rn(keep, 255.0));
        r = vb_cast_u8(__dmul_rn(__dadd_rn(keep, (double)b), 255.0));
    }
}
VB_KERNEL void ovl_rgb_kernel(OvlRgbArgs p) {
    const int si = (int)blockIdx.y;
    const VbSubject S = p.subj[si];
    const OvlStats o = p.st[si];
    const bool nan = o.nan[0] != 0;
    const float mn = nan ? vb_bits_float(0x7fc00000u) : vb_key_float(o.kmin[0]);
    const float mx = nan ? mn : vb_key_float(o.kmax[0]);
    const bool flat = !nan && (__fsub_rn(mx, mn) == 0.0f);
    // four voxels per thread and step: one 16-byte and one 4-byte load, three 4-byte stores (a warp
    // writes 384 contiguous bytes); volumes at an unaligned offset take the scalar tail loop
    const bool al = (S.off & 3) == 0;
    const int nq = al ? (S.V >> 2) : 0;
    for (int q = (int)(blockIdx.x * blockDim.x + threadIdx.x); q < nq; q += (int)(gridDim.x * blockDim.x)) {
        const size_t g = (size_t)S.off + 4 * (size_t)q;
        const float4 xv = *(const float4*)(p.a + g);
        const uint32_t bw4 = *(const uint32_t*)(p.b + g);
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
        unsigned bytes[12];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            unsigned r, gb;
            ovl_rgb_one(p, fabsf(xs[e]), (bw4 >> (8 * e)) & 0xFFu, mn, mx, flat, r, gb);
            bytes[3 * e] = r; bytes[3 * e + 1] = gb; bytes[3 * e + 2] = gb;
        }
        uint32_t* ow = (uint32_t*)(p.rgb + 3 * g);
#pragma unroll
        for (int w = 0; w < 3; ++w)
            ow[w] = bytes[4 * w] | (bytes[4 * w + 1] << 8) | (bytes[4 * w + 2] << 16) | (bytes[4 * w + 3] << 24);
    }
    for (int v = 4 * nq + (int)(blockIdx.x * blockDim.x + threadIdx.x); v < S.V; v += (int)(gridDim.x * blockDim.x)) {
        const size_t g = (size_t)S.off + (size_t)v;
        unsigned r, gb;
        ovl_rgb_one(p, fabsf(p.a[g]), p.b[g], mn, mx, flat, r, gb);
        uint8_t* q = p.rgb + 3 * g;
        q[0] = (uint8_t)r; q[1] = (uint8_t)gb; q[2] = (uint8_t)gb;
    }
}

// screenShot montage (before the text): uint8 [7 R][S C][3] per subject at out + out_off[si]; row
// band i: 0, 1 blank | 2 proton | 3 HPvent | 4 N4 with the mask border in cyan | 5 N4 with defects
// in red | 6 N4 with the CI map in parula colours (index int(CI * 64 / 40), Vent_Analysis.py:482-484).
struct OvlMontageArgs {
    const float* proton;          // may be null (blank band; the reference needs it)
    const float* image;
    const float* n4;
    const uint8_t* mborder;
    const uint8_t* defect;
    const uint16_t* shell;        // may be null: no CI band colours
    const VbSubject* subj;
    OvlStats* st;
    const CiTableDev* tables;     // shell radii per table slot (S.table); unused when shell is null
    const double* vox_min;        // [n]
    const double* parula;         // [64][3]
    const long long* out_off;     // [n] byte offsets
    uint8_t* out;
    int image_mode, proton_mode;
};
VB_KERNEL void ovl_montage_kernel(OvlMontageArgs p) {
    const int si = (int)blockIdx.y;
    const VbSubject S = p.subj[si];
    OvlStats* op = p.st + si;
    const OvlStats o = *op;
    if (o.error == 1) return;
    const int r0 = o.box[0], c0 = o.box[2], s0 = o.box[4];
    const int R = o.box[1] - r0, Cc = o.box[3] - c0, Sl = o.box[5] - s0;
    float mn[3], mx[3];
    bool flat[3];
    for (int a = 0; a < 3; ++a) {
        const bool nan = o.nan[a] != 0;
        mn[a] = nan ? vb_bits_float(0x7fc00000u) : vb_key_float(o.kmin[a]);
        mx[a] = nan ? mn[a] : vb_key_float(o.kmax[a]);
        flat[a] = !nan && (__fsub_rn(mx[a], mn[a]) == 0.0f);
    }
    const double* radii = (p.shell && p.tables && S.table >= 0) ? p.tables[S.table].radii : nullptr;
    const int nb = radii ? p.tables[S.table].n_bounds : 0;
    const double vmin = radii ? p.vox_min[si] : 0.0;
    uint8_t* out = p.out + p.out_off[si];
    const int W = Sl * Cc, npx = 7 * R * W;
    for (int e = (int)(blockIdx.x * blockDim.x + threadIdx.x); e < npx; e += (int)(gridDim.x * blockDim.x)) {
        const int x = e % W, y = e / W;
        const int band = y / R, r = y - band * R, s = x / Cc, cidx = x - s * Cc;
        const size_t g = (size_t)S.off + ((size_t)(r0 + r) * S.n1 + (c0 + cidx)) * S.n2 + (s0 + s);
        double red = 0.0, grn = 0.0, blu = 0.0;
        if (band == 2) {
            if (p.proton) red = grn = blu = ovl_norm(p.proton[g], mn[2], mx[2], flat[2], p.proton_mode);
        } else if (band == 3) {
            red = grn = blu = ovl_norm(p.image[g], mn[1], mx[1], flat[1], p.image_mode);
        } else if (band >= 4) {
            const double n4 = ovl_norm(p.n4[g], mn[0], mx[0], flat[0], VB_OVL_F32);
            if (band == 4) {
                const bool bd = p.mborder[g] != 0;
                const double keep = bd ? __dmul_rn(n4, 0.0) : n4;            // N4*(~border)
                red = keep; grn = blu = bd ? __dadd_rn(keep, 1.0) : keep;
            } else if (band == 5) {
                const bool d = p.defect[g] != 0;
                const double keep = d ? __dmul_rn(n4, 0.0) : n4;
                red = d ? __dadd_rn(keep, 1.0) : keep; grn = blu = keep;
            } else {
                double ci = 0.0;
                if (radii) {
                    const int t = (int)p.shell[g];
                    if (t < nb) ci = __dmul_rn(radii[t], vmin);
                }
                if (ci > 0.0) {
                    const int idx = (int)__ddiv_rn(__dmul_rn(ci, 64.0), 40.0);
                    if (idx >= 64) { op->error = 2; red = grn = blu = 0.0; }
                    else {
                        const double z = __dmul_rn(n4, 0.0);                  // N4*(CI==0)
                        red = __dadd_rn(z, p.parula[3 * idx]); grn = __dadd_rn(z, p.parula[3 * idx + 1]); blu = __dadd_rn(z, p.parula[3 * idx + 2]);
                    }
                } else {
                    red = grn = blu = n4;                                     // + colour * 0
                }
            }
        }
        uint8_t* q = out + 3 * (size_t)e;
        q[0] = (uint8_t)vb_cast_u8(__dmul_rn(red, 255.0));
        q[1] = (uint8_t)vb_cast_u8(__dmul_rn(grn, 255.0));
        q[2] = (uint8_t)vb_cast_u8(__dmul_rn(blu, 255.0));
    }
}
