// Compact host I/O of the path (vb_pipeline_batch_host_packed): what crosses PCIe is the image as
// the reference reads it from DICOM (uint16 pixel data, Vent_Analysis.py:169-181; cast to float32
// here exactly as N4_bias_correction does at :322), the 0/1 mask bit-packed, one byte of codes per
// voxel (defectArrayLB | defectArray | defectBorder | mask_border) and the CI map as a sparse
// (voxel, shell) list.  Pure streaming kernels; every one moves 16 bytes per thread and step.
#pragma once
#include "vb_common.cuh"

// image_u16 -> float32 and mask bits -> bytes.  grid = (chunks, subjects); a thread expands one
// 32-voxel mask word (32 bytes out) and converts the 32 pixels under it (64 bytes in, 128 out).
struct UnpackArgs {
    const uint16_t* image_u16;    // [V_total] or null (image already float32 on the device)
    const uint32_t* mask_bits;    // subject i at mask_word_off[i], ceil(V/32) words; or null
    const long long* bit_off;     // [n] first mask word of every subject
    const VbSubject* subj;
    float* image;                 // [V_total] out
    uint8_t* mask;                // [V_total] out
};
VB_KERNEL void io_unpack_kernel(UnpackArgs a) {
    const int si = (int)blockIdx.y;
    const VbSubject S = a.subj[si];
    const int words = (S.V + 31) >> 5;
    const uint32_t* bits = a.mask_bits ? a.mask_bits + a.bit_off[si] : nullptr;
    const bool al = (S.off & 7) == 0;                       // 16-byte aligned u16 loads, 32-byte aligned f32 stores
    for (int w = (int)(blockIdx.x * blockDim.x + threadIdx.x); w < words; w += (int)(gridDim.x * blockDim.x)) {
        const int v0 = w << 5;
        const bool full = v0 + 32 <= S.V;
        if (bits) {
            const uint32_t m = bits[w];
            uint8_t* o = a.mask + S.off + v0;
            if (full && (S.off & 3) == 0) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint32_t nib = (m >> (4 * q)) & 0xFu;
                    // bits b0..b3 -> bytes 0/1
                    *(uint32_t*)(o + 4 * q) = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
                }
            } else {
                for (int b = 0; b < 32 && v0 + b < S.V; ++b) o[b] = (uint8_t)((m >> b) & 1u);
            }
        }
        if (a.image_u16) {
            const uint16_t* p = a.image_u16 + S.off + v0;
            float* o = a.image + S.off + v0;
            if (full && al) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint4 t = *(const uint4*)(p + 8 * q);
                    const uint32_t ws[4] = {t.x, t.y, t.z, t.w};
                    float f[8];
#pragma unroll
                    for (int e = 0; e < 4; ++e) { f[2 * e] = (float)(ws[e] & 0xFFFFu); f[2 * e + 1] = (float)(ws[e] >> 16); }
                    *(float4*)(o + 8 * q) = make_float4(f[0], f[1], f[2], f[3]);
                    *(float4*)(o + 8 * q + 4) = make_float4(f[4], f[5], f[6], f[7]);
                }
            } else {
                for (int b = 0; b < 32 && v0 + b < S.V; ++b) o[b] = (float)p[b];
            }
        }
    }
}

// codes = LB | defect << 3 | defectBorder << 4 | mask_border << 5.  16 voxels per thread and step.
struct CodesArgs {
    const uint8_t* lb;
    const uint8_t* defect;
    const uint8_t* dborder;       // may be null
    const uint8_t* mborder;       // may be null
    uint8_t* codes;
    long long v0, v1;             // voxel range of the batch to convert
};
VB_KERNEL void io_codes_kernel(CodesArgs a) {
    const long long nt = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long head = min(a.v1, (a.v0 + 15) & ~15LL);   // unaligned head / tail: bytes
    for (long long v = a.v0 + t0; v < head; v += nt)
        a.codes[v] = (uint8_t)(a.lb[v] | (a.defect[v] << 3) | ((a.dborder ? a.dborder[v] : 0) << 4) | ((a.mborder ? a.mborder[v] : 0) << 5));
    const long long nq = (a.v1 - head) >> 4;
    for (long long q = t0; q < nq; q += nt) {
        const long long v = head + (q << 4);
        const uint4 l = *(const uint4*)(a.lb + v), d = *(const uint4*)(a.defect + v);
        uint4 r;
        r.x = l.x | (d.x << 3); r.y = l.y | (d.y << 3); r.z = l.z | (d.z << 3); r.w = l.w | (d.w << 3);
        if (a.dborder) { const uint4 b = *(const uint4*)(a.dborder + v); r.x |= b.x << 4; r.y |= b.y << 4; r.z |= b.z << 4; r.w |= b.w << 4; }
        if (a.mborder) { const uint4 b = *(const uint4*)(a.mborder + v); r.x |= b.x << 5; r.y |= b.y << 5; r.z |= b.z << 5; r.w |= b.w << 5; }
        *(uint4*)(a.codes + v) = r;
    }
    for (long long v = head + (nq << 4) + t0; v < a.v1; v += nt)
        a.codes[v] = (uint8_t)(a.lb[v] | (a.defect[v] << 3) | ((a.dborder ? a.dborder[v] : 0) << 4) | ((a.mborder ? a.mborder[v] : 0) << 5));
}

// Sparse CI map of the subjects [0, n) of a chunk: entries of subject i start at the sum of the
// defect counts of the chunk's earlier subjects.  One CTA per subject.
struct SparseArgs {
    const VbSubject* subj;        // chunk's first subject
    const int* count;             // [n] defect counts (ci_pack_kernel)
    const int* list;              // row-fastest flat indices, subject i at S.off
    const uint16_t* shell;        // [V_total]
    uint32_t* index;              // chunk's output base
    uint16_t* value;
    int n;
};
VB_KERNEL void io_ci_sparse_kernel(SparseArgs a) {
    VB_SHARED long long red[64];
    const int si = (int)blockIdx.x, tid = (int)threadIdx.x, nt = (int)blockDim.x;
    long long before = 0;
    for (int j = tid; j < si; j += nt) before += a.count[j];
    before = vb_block_reduce(before, VbAddLL(), 0LL, red);
    const VbSubject S = a.subj[si];
    const int D = a.count[si], n0 = S.n0, n1 = S.n1, n2 = S.n2;
    const int* list = a.list + S.off;
    for (int d = tid; d < D; d += nt) {
        const int f = list[d];
        const int i = f % n0, jk = f / n0, j = jk % n1, k = jk / n1;
        const int c = (i * n1 + j) * n2 + k;
        a.index[before + d] = (uint32_t)c;
        a.value[before + d] = a.shell[S.off + c];
    }
}
