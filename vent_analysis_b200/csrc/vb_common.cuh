// Shared device helpers: subject descriptors, block-wide reductions.
#pragma once
#include "vb_platform.h"
#include "../../include/ventb200.h"

// One subject of a ragged batch: a C-ordered [n0][n1][n2] volume at voxel offset `off`.
struct VbSubject {
    long long off;   // first voxel in the concatenated arrays
    int n0, n1, n2;  // rows, cols, slices
    int V;           // n0*n1*n2
    int table;       // CI table slot (device array index), -1 if none
    int pad;
};

#define VB_FULL 0xffffffffu

VB_DEV int vb_lane() { return (int)(threadIdx.x & 31u); }
VB_DEV int vb_warp() { return (int)(threadIdx.x >> 5); }

// Block-wide reductions through a small shared scratch (>= 32 elements of T).  Fixed
// reduction tree => bitwise reproducible for a given block size.  All threads must call.
template <class T, class Op>
VB_DEV T vb_block_reduce(T v, Op op, T identity, T* scratch) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v = op(v, __shfl_xor_sync(VB_FULL, v, m));
    const int nw = (int)((blockDim.x + 31u) >> 5);
    __syncthreads();                       // scratch may still be read from a previous call
    if (vb_lane() == 0) scratch[vb_warp()] = v;
    __syncthreads();
    T r = (vb_lane() < nw) ? scratch[vb_lane()] : identity;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) r = op(r, __shfl_xor_sync(VB_FULL, r, m));
    return r;                              // every thread holds the result
}

struct VbAddD { VB_DEV_MEMBER double operator()(double a, double b) const { return a + b; } };
struct VbAddF { VB_DEV_MEMBER float operator()(float a, float b) const { return a + b; } };
struct VbAddI { VB_DEV_MEMBER int operator()(int a, int b) const { return a + b; } };
struct VbAddLL { VB_DEV_MEMBER long long operator()(long long a, long long b) const { return a + b; } };
struct VbMinF { VB_DEV_MEMBER float operator()(float a, float b) const { return fminf(a, b); } };
struct VbMaxF { VB_DEV_MEMBER float operator()(float a, float b) const { return fmaxf(a, b); } };
struct VbMinI { VB_DEV_MEMBER int operator()(int a, int b) const { return a < b ? a : b; } };
struct VbMaxI { VB_DEV_MEMBER int operator()(int a, int b) const { return a > b ? a : b; } };

// order-preserving float <-> uint key (radix select).  The bit pattern is moved through an opaque
// register move: seen as `bitcast(f) | signbit`, the compiler emits -|f| as a FLOAT add, which
// replaces a NaN by the canonical positive NaN and so changes its key (found on the B200 by the
// NaN-in-mask parity case).
VB_DEV unsigned vb_float_bits(float f) {
#if defined(VB_EMULATE)
    return __float_as_uint(f);
#else
    unsigned u;
    asm("mov.b32 %0, %1;" : "=r"(u) : "f"(f));
    return u;
#endif
}
VB_DEV float vb_bits_float(unsigned u) {
#if defined(VB_EMULATE)
    return __uint_as_float(u);
#else
    float f;
    asm("mov.b32 %0, %1;" : "=f"(f) : "r"(u));
    return f;
#endif
}
VB_DEV unsigned vb_float_key(float f) {
    const unsigned u = vb_float_bits(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
VB_DEV float vb_key_float(unsigned k) {
    const unsigned u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return vb_bits_float(u);
}
