// libventb200: C ABI (include/ventb200.h) + host orchestration of the sm_100a kernels.
// Also compiled for the host with -DVB_EMULATE by tests/emu (test infrastructure only).
#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include <unordered_set>
#include <vector>

#include "vb_rt.h"
#include "vb_common.cuh"
#include "n4_kernels.cuh"
#include "vdp_kernels.cuh"
#include "ci_kernels.cuh"
#include "io_kernels.cuh"
#include "overlay_kernels.cuh"
#include "recon_kernels.cuh"

namespace {

struct Buf {
    void* p = nullptr;
    size_t cap = 0;
};

// per-resident-CTA workspace of the N4 kernel; two sets so that the launches of consecutive chunks
// of the host pipeline (two compute streams) can overlap
#define VB_MAX_CSTREAMS 4
struct N4Ws {
    Buf logI, u, colmask, rbb, wbase, acol, order, ec, code, axis, counter, cl_hist, cl_q, cl_red, cl_bar, cl_bcast;
};

struct TableHost {
    int n0, n1, n2;
    int n_unique, n_bounds;
    void* d_rows = nullptr;
    void* d_ii = nullptr;
    void* d_up = nullptr;      // per 32-row step: largest row count of a boundary that ends inside it (0: none)
    void* d_radii = nullptr;   // shell radii (double), for vb_pack6_*
    std::vector<double> radii;
};

}  // namespace

struct vb_ctx {
    int device = 0, sms = 0, max_smem = 0;
    vbrt::stream_t stream = nullptr, stream_in = nullptr, stream_out = nullptr;
    std::vector<vbrt::event_t> ev_copy;   // host pipeline: inputs-ready / outputs-ready events per chunk
    int r0 = 0;                          // first subject of the range the stage runners work on
    uint64_t launches = 0;
    bool debug = false;
    int n4_cluster_max = 0;   // vb_set_n4_cluster (0: automatic)
    int n4_cluster = 1;       // CTAs per subject of the last N4 launch
    int n4_parts = 1;         // chunks / parts of the schedule the next N4 launch belongs to (> 1: several N4 kernels in flight)
    // subject descriptors
    std::vector<VbSubject> last_subj;
    std::vector<long long> last_wordoff;
    Buf h_desc, h_stats, d_subj, d_wordoff;   // h_*: pinned host staging
    long long v_total = 0, words_total = 0;
    // workspaces
    N4Ws n4ws[VB_MAX_CSTREAMS];
    int wsi = 0;                         // N4 workspace set the next N4 launch uses (= compute stream of the chunk)
    int ksi = 0;                         // sort-scratch set the next vdp_stats launch uses (= parity of the chunk)
    vbrt::stream_t cextra[VB_MAX_CSTREAMS - 2] = {};   // further compute streams of the host pipelines (after stream, stream2)
    int n_cstreams = 4;                  // compute streams the host pipelines rotate over (measured on 1776 C2 subjects: 2: 88.2, 3: 86.4, 4: 85.8 ms)
    vbrt::stream_t stream2 = nullptr;    // second compute stream (host pipeline chunks; parts of a resident batch)
    vbrt::stream_t post[2] = {nullptr, nullptr};   // high-priority streams of the stages after N4 (VDP, borders, CI, output packing)
    vbrt::stream_t pre[2] = {nullptr, nullptr};    // high-priority streams of the stages before N4 (input unpacking, SNR)
    std::vector<vbrt::event_t> ev_pre;   // inputs of chunk k unpacked, SNR done -> its N4 may start
    std::vector<vbrt::event_t> ev_post;  // N4 of chunk k done -> its later stages may start
    vbrt::event_t ev_fork = nullptr, ev_joins[3] = {nullptr, nullptr, nullptr};
    Buf ws_d0, ws_bits, ws_list, ws_count, dbg;
    Buf ws_keys_a[2], ws_keys_b[2];     // sort scratch of vdp_stats_kernel, one set per compute stream (like n4ws)
    int dbg_stride = 0, dbg_nbins = 0, dbg_Lmax = 0, dbg_n = 0;
    // CI tables
    std::vector<TableHost> tables;
    Buf d_tables;
    bool tables_dirty = true;
    // per-stage CUDA-event timing of vb_pipeline_batch_dev (vb_set_profiling)
    bool prof = false;
    std::vector<vbrt::event_t> ev;      // (VB_N_STAGES + 1) events per profiled call
    size_t ev_used = 0;
    // host-API mirrors
    Buf m_image, m_mask, m_n4, m_logbias, m_defect, m_lb, m_dborder, m_mborder, m_shell, m_stats;
    Buf m_proton, m_pack, d_voxmin;
    Buf m_u16, m_bits, m_codes, m_ci_idx, m_ci_val, d_bitoff;   // compact host I/O
    std::vector<long long> last_bitoff;
    Buf m_k[2], d_ktmp;                                     // k-space reconstruction
    Buf d_ovl, m_rgb, m_ovl_in[6], d_parula, d_outoff;   // overlay compositions
};

namespace {

int fail(int code, const std::string& msg) {
    vbrt::set_error(msg);
    return code;
}

bool ensure(Buf& b, size_t bytes) {
    if (bytes <= b.cap && b.p) return true;
    vbrt::dev_free(b.p);
    b.p = nullptr; b.cap = 0;
    const size_t want = bytes + bytes / 8 + 256;
    if (!vbrt::dev_malloc(&b.p, want)) return false;
    b.cap = want;
    return true;
}

// Upload the batch description (cached: a repeated batch costs nothing).
int prepare(vb_ctx* c, int n, const int32_t* shapes, const int32_t* handles, vbrt::stream_t s) {
    if (n <= 0 || !shapes) return fail(VB_ERR_ARG, "empty batch");
    if (n > 65535) return fail(VB_ERR_UNSUPPORTED, "more than 65535 subjects in one batch");
    std::vector<VbSubject> subj((size_t)n);
    std::vector<long long> woff((size_t)n + 1);
    long long off = 0, w = 0;
    for (int i = 0; i < n; ++i) {
        const long long n0 = shapes[3 * i], n1 = shapes[3 * i + 1], n2 = shapes[3 * i + 2];
        if (n0 < 1 || n1 < 1 || n2 < 1) return fail(VB_ERR_ARG, "non-positive shape");
        const long long V = n0 * n1 * n2;
        if (V > (1LL << 30)) return fail(VB_ERR_UNSUPPORTED, "volume larger than 2^30 voxels");
        VbSubject S;
        S.off = off; S.n0 = (int)n0; S.n1 = (int)n1; S.n2 = (int)n2; S.V = (int)V;
        S.table = handles ? handles[i] : -1; S.pad = 0;
        if (handles) {
            if (S.table < 0 || S.table >= (int)c->tables.size()) return fail(VB_ERR_ARG, "bad CI table handle");
            const TableHost& T = c->tables[(size_t)S.table];
            if (T.n0 != S.n0 || T.n1 != S.n1 || T.n2 != S.n2)
                return fail(VB_ERR_ARG, "CI table was prepared for a different volume shape");
        }
        subj[(size_t)i] = S;
        woff[(size_t)i] = w;
        off += V;
        w += (V + 31) / 32 + 1;
    }
    woff[(size_t)n] = w;
    const bool same = subj.size() == c->last_subj.size() &&
                      memcmp(subj.data(), c->last_subj.data(), subj.size() * sizeof(VbSubject)) == 0;
    c->v_total = off; c->words_total = w;
    if (same) return VB_OK;
    // a previous upload may still be in flight from the staging buffer
    if (!vbrt::stream_sync(s)) return VB_ERR_CUDA;
    const size_t sb = subj.size() * sizeof(VbSubject), wb = woff.size() * sizeof(long long);
    if (c->h_desc.cap < sb + wb) {
        vbrt::host_free(c->h_desc.p);
        c->h_desc.p = nullptr; c->h_desc.cap = 0;
        if (!vbrt::host_malloc(&c->h_desc.p, (sb + wb) * 2)) return VB_ERR_CUDA;
        c->h_desc.cap = (sb + wb) * 2;
    }
    if (!ensure(c->d_subj, sb) || !ensure(c->d_wordoff, wb)) return VB_ERR_CUDA;
    memcpy(c->h_desc.p, subj.data(), sb);
    memcpy((char*)c->h_desc.p + sb, woff.data(), wb);
    if (!vbrt::h2d(c->d_subj.p, c->h_desc.p, sb, s)) return VB_ERR_CUDA;
    if (!vbrt::h2d(c->d_wordoff.p, (char*)c->h_desc.p + sb, wb, s)) return VB_ERR_CUDA;
    // the descriptors are read by kernels on OTHER streams too (the host pipeline's second compute
    // stream, a later *_dev call on a different caller stream): complete the upload here -- it
    // happens once per distinct batch
    if (!vbrt::stream_sync(s)) return VB_ERR_CUDA;
    c->last_subj.swap(subj);
    c->last_wordoff.swap(woff);
    return VB_OK;
}

const VbSubject* subj_ptr(const vb_ctx* c) { return (const VbSubject*)c->d_subj.p + c->r0; }

int grid_x_for(const vb_ctx* c, int n, long long vmax, int per_block) {
    long long want = (c->sms * 8LL + n - 1) / n;
    const long long cap = (vmax + per_block - 1) / per_block;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    if (want > 4096) want = 4096;
    return (int)want;
}

long long vmax_of(const vb_ctx* c) {
    long long m = 0;
    for (const VbSubject& S : c->last_subj) m = std::max<long long>(m, S.V);
    return m;
}

int upload_tables(vb_ctx* c, vbrt::stream_t s) {
    if (!c->tables_dirty) return VB_OK;
    std::vector<CiTableDev> t(c->tables.size());
    for (size_t i = 0; i < t.size(); ++i) {
        t[i].rows = (const int2*)c->tables[i].d_rows;
        t[i].ii = (const int*)c->tables[i].d_ii;
        t[i].up = (const int*)c->tables[i].d_up;
        t[i].radii = (const double*)c->tables[i].d_radii;
        t[i].n_unique = c->tables[i].n_unique;
        t[i].n_bounds = c->tables[i].n_bounds;
    }
    if (!vbrt::stream_sync(s)) return VB_ERR_CUDA;
    if (!ensure(c->d_tables, t.size() * sizeof(CiTableDev))) return VB_ERR_CUDA;
    if (!vbrt::h2d(c->d_tables.p, t.data(), t.size() * sizeof(CiTableDev), s)) return VB_ERR_CUDA;
    if (!vbrt::stream_sync(s)) return VB_ERR_CUDA;      // `t` is a stack object
    c->tables_dirty = false;
    return VB_OK;
}

// ------------------------------------------------------------------------------------- stages
int run_snr(vb_ctx* c, const float* image, const uint8_t* mask, int n, int fov, vb_subject_stats* stats,
            vbrt::stream_t s) {
    vbrt::clear_stale_error();
    for (const VbSubject& S : c->last_subj)
        if (S.n0 + S.n1 + S.n2 > VB_SNR_MAX_AXIS) return fail(VB_ERR_UNSUPPORTED, "SNR: n0+n1+n2 too large");
    const int grid = std::min(n, c->sms * 4);
    int threads = VB_SNR_THREADS;
    if (const char* e = getenv("VB_SNR_THREADS_RT")) threads = std::max(32, std::min(VB_VDP_THREADS, atoi(e) & ~31));   // tuning
    // ring of bulk-copy stages (two CTAs per SM): 6 stages of whole rows; rows that do not fit keep
    // the direct-load path inside the kernel
    int ring = 0;
    if (!getenv("VB_SNR_NO_BULK")) {
        const long long statics = vbrt::static_smem(snr_kernel);
        ring = (int)std::max<long long>(0, std::min<long long>(100 * 1024, ((long long)c->max_smem - 4096) / 2 - statics)) & ~255;
        if (!vbrt::allow_smem(snr_kernel, ring)) return VB_ERR_CUDA;
    }
    VB_LAUNCH(snr_kernel, dim3(grid), dim3(threads), (size_t)ring, s, image, mask, subj_ptr(c), n,
              fov, stats, ring);
    ++c->launches;
    return vbrt::last_launch_ok("snr_kernel") ? VB_OK : VB_ERR_CUDA;
}

int run_n4(vb_ctx* c, const float* image, const uint8_t* mask, int n, const vb_n4_params* pp, float* out,
           float* logbias, vb_subject_stats* stats, vbrt::stream_t s) {
    vbrt::clear_stale_error();
    vb_n4_params p = *pp;
    if (p.n_levels < 1 || p.n_levels > VB_MAX_LEVELS) return fail(VB_ERR_ARG, "N4: n_levels out of range");
    if (p.n_bins < 2) return fail(VB_ERR_ARG, "N4: n_bins < 2");
    int ncpf[3];
    for (int a = 0; a < 3; ++a) {
        if (p.ncp[a] < 4) return fail(VB_ERR_ARG, "N4: cubic splines need >= 4 control points");
        long long v = p.ncp[a];
        for (int l = 1; l < p.n_levels; ++l) v = 2 * v - 3;
        if (v > 4096) return fail(VB_ERR_UNSUPPORTED, "N4: lattice too large");
        ncpf[a] = (int)v;
    }
    const long long Lmax = (long long)ncpf[0] * ncpf[1] * ncpf[2];
    int e = 0;
    while ((1 << e) < p.n_bins) ++e;
    const int fftN = 1 << (e + 1);
    long long vmax = 0, cmw = 0, axis = 0, n2max = 0, rbb = 0, wb = 0, ncmax = 0, n0max = 0;
    for (const VbSubject& S : c->last_subj) {
        const long long ncols = (long long)S.n1 * S.n2, nblk = (ncols + 31) / 32;
        vmax = std::max<long long>(vmax, S.V);
        cmw = std::max<long long>(cmw, (long long)((S.n0 + 31) / 32) * ncols);
        axis = std::max<long long>(axis, 16LL * (std::max(S.n0, ncpf[0]) + std::max(S.n1, ncpf[1]) + std::max(S.n2, ncpf[2])));
        n2max = std::max<long long>(n2max, S.n2);
        const long long W = (S.n0 + 31) / 32;
        rbb = std::max<long long>(rbb, std::max((long long)S.n0 * nblk + 1, ncols + 2));
        wb = std::max<long long>(wb, nblk * W + 1);
        ncmax = std::max<long long>(ncmax, ncols);
        n0max = std::max<long long>(n0max, S.n0);
    }
    // shared memory: [union][L][Lold][Om][Phi][E][H][red][row tables][compact log image]
    const long long Lb = 4 * ((Lmax + 3) & ~3LL), Hb = 4 * ((p.n_bins + 3) & ~3);
    const long long rowtab_rows = (40 * n0max <= 20480) ? ((n0max + 3) & ~3LL) : 0;
    const long long fixed = 4 * Lb + 2 * Hb + VB_N4_RED * 8 + 8LL * fftN + 40 * rowtab_rows;
    long long ub = 4LL * VB_N4_HCOPIES * (p.n_bins | 1);                 // histogram copies
    ub = std::max(ub, 8LL * 6 * fftN);                                   // FFT work arrays (the twiddles have their own place)
    ub = std::max(ub, 8LL * (long long)ncpf[0] * ncpf[1] * n2max);       // Q[a][b][s], 64-bit fixed point
    ub = std::max(ub, 4LL * Lmax);                                       // refinement scratch
    ub = (ub + 15) & ~15LL;
    const long long static_smem = vbrt::static_smem(n4_subject_kernel);    // s_next, chain state
    if (ub + fixed + static_smem > c->max_smem)
        return fail(VB_ERR_UNSUPPORTED, "N4: working set does not fit the 227 KB of shared memory");
    long long cap = ((long long)c->max_smem - ub - fixed - static_smem) / 4;
    if (const char* e = getenv("VB_N4_SMEM_CAP")) cap = std::min<long long>(cap, std::max(0, atoi(e)));   // tests: U in global memory / the column-walk update
    // A batch that leaves most SMs idle runs in CLUSTER MODE: a thread-block cluster of K CTAs per subject
    // (n4_kernels.cuh; latency of single subjects and small batches).  Not with the debug capture.
    // The automatic choice follows the B200 timings in profiles/r02_n4_cluster_timings.jsonl: volumes below
    // ~200 k voxels lose (2.0 -> 2.8 ms at 48x40x6: the cross-CTA barriers cost more than the split saves); a
    // grid that fills the last SM loses too (18 subjects x 8 CTAs: 7.6 ms against 6.1 with 4) because clusters
    // are placed within a GPC, so one eighth of the SMs stays free; chunks of a multi-stream schedule keep one
    // CTA per subject (their kernels share the SMs).  vb_set_n4_cluster(2 / 4 / 8) overrides the first two.
    int K = 1;
    {
        const bool forced = c->n4_cluster_max > 1;
        int want = (c->n4_cluster_max > 0) ? c->n4_cluster_max : 8;
        long long min_v = 200000, limit = forced ? c->sms : (long long)c->sms * 7 / 8;
        if (const char* e = getenv("VB_N4_CLUSTER_MIN_V")) min_v = atoll(e);                      // tests
        if (!forced && vmax < min_v) want = 1;
        if (c->n4_parts > 1) want = 1;
        if (const char* e = getenv("VB_N4_CLUSTER")) { want = std::max(1, std::min(8, atoi(e))); limit = c->sms; }   // tuning
        while (want & (want - 1)) want &= want - 1;          // a power of two
        while (want > 1 && (long long)n * want > limit) want >>= 1;
        if (!c->debug) K = want;
    }
    const int grid = (K > 1) ? n * K : std::min(n, c->sms);
    const int slots = (K > 1) ? n : grid;                 // workspace slots shared by the CTAs of a cluster
    N4Ws& W = c->n4ws[c->wsi % VB_MAX_CSTREAMS];
    if (!ensure(W.logI, sizeof(float) * (size_t)slots * (size_t)vmax) || !ensure(W.u, sizeof(float) * (size_t)slots * (size_t)vmax) ||
        !ensure(W.colmask, 4 * (size_t)slots * (size_t)cmw) || !ensure(W.rbb, 4 * (size_t)slots * (size_t)rbb) ||
        !ensure(W.wbase, 4 * (size_t)slots * (size_t)wb) || !ensure(W.acol, 4 * (size_t)slots * (size_t)ncmax) ||
        !ensure(W.order, 4 * (size_t)slots * (size_t)wb) || !ensure(W.ec, 4 * (size_t)slots * (size_t)((vmax + 3) & ~3LL)) ||
        !ensure(W.code, 4 * (size_t)slots * (size_t)vmax) ||
        !ensure(W.axis, 4 * (size_t)grid * (size_t)axis) || !ensure(W.counter, 64))
        return VB_ERR_CUDA;
    const long long q_max = (long long)ncpf[0] * ncpf[1] * n2max;          // Q[a][b][s] entries at the last level
    if (K > 1) {
        if (!ensure(W.cl_hist, 4 * 8 * (size_t)slots * K * (size_t)p.n_bins) || !ensure(W.cl_q, 4 * 8 * (size_t)slots * K * (size_t)q_max) ||
            !ensure(W.cl_red, 2 * 32 * (size_t)slots * K) || !ensure(W.cl_bar, 16 * (size_t)slots) || !ensure(W.cl_bcast, 64 * (size_t)slots))
            return VB_ERR_CUDA;
        if (!vbrt::dev_memset(W.cl_bar.p, 0, 16 * (size_t)slots, s)) return VB_ERR_CUDA;
    }
    int trace = 0;
    for (int l = 0; l < p.n_levels; ++l) trace += p.max_iter[l];
    float* dbg = nullptr;
    int dbg_stride = 0;
    if (c->debug) {
        dbg_stride = VB_DBG_HDR + 2 * p.n_bins + (int)Lmax + trace;
        // one record per subject of the WHOLE batch (the host pipeline runs it in chunks)
        const size_t ntot = c->last_subj.size();
        if (!ensure(c->dbg, sizeof(float) * ntot * dbg_stride)) return VB_ERR_CUDA;
        dbg = (float*)c->dbg.p + (size_t)c->r0 * dbg_stride;
        c->dbg_stride = dbg_stride; c->dbg_nbins = p.n_bins; c->dbg_Lmax = (int)Lmax; c->dbg_n = (int)ntot;
        if (!vbrt::dev_memset(dbg, 0, sizeof(float) * (size_t)n * dbg_stride, s)) return VB_ERR_CUDA;
    }
    if (!vbrt::dev_memset(W.counter.p, 0, 64, s)) return VB_ERR_CUDA;
    N4Args a;
    a.image = image; a.mask = mask; a.out = out; a.logbias = logbias;
    a.subj = subj_ptr(c); a.n = n; a.p = p; a.stats = stats;
    a.ws_logI = (float*)W.logI.p; a.ws_u = (float*)W.u.p; a.ws_colmask = (uint32_t*)W.colmask.p; a.ws_rbb = (uint32_t*)W.rbb.p;
    a.ws_wbase = (uint32_t*)W.wbase.p; a.ws_acol = (uint32_t*)W.acol.p; a.ws_order = (uint32_t*)W.order.p; a.ws_axis = (float*)W.axis.p;
    a.ws_ec = (float*)W.ec.p; a.ws_code = (uint32_t*)W.code.p; a.rowtab_rows = (int)rowtab_rows;
    { const char* e = getenv("VB_N4_NO_OVERLAP"); a.no_overlap = (e && e[0] == '1') ? 1 : 0; }
    a.work_counter = (int*)W.counter.p; a.debug = dbg;
    a.K = K; a.q_max = (int)q_max;
    // chain team of the cluster kernel: the phase team of a CTA has 1 / K of the one-CTA work while every CTA runs the
    // whole chain, so the chain gets more of the 28 warps as K grows (profiles/r02_n4_cluster_chain_sweep.jsonl:
    // C4 subject, K = 8: 12 warps 58.7 ms, 16: 51.1, 20: 47.1, 24: 60.6; K = 4: 16 warps best; K = 2: 12)
    a.cl_chain_warps = (K >= 8) ? 20 : (K >= 4) ? 16 : 12;
    if (const char* e = getenv("VB_N4_CL_CHAIN_WARPS")) a.cl_chain_warps = std::max(4, std::min(VB_N4_THREADS / 32 - 1, atoi(e)));   // tuning (the chain needs a team of >= 4 warps)
    a.cl_hist = (K > 1) ? (unsigned long long*)W.cl_hist.p : nullptr; a.cl_q = (K > 1) ? (unsigned*)W.cl_q.p : nullptr;
    a.cl_red = (K > 1) ? (double*)W.cl_red.p : nullptr; a.cl_bar = (K > 1) ? (unsigned*)W.cl_bar.p : nullptr;
    a.cl_bcast = (K > 1) ? (int*)W.cl_bcast.p : nullptr;
    a.Vmax = vmax; a.cmw_max = cmw; a.rbb_max = rbb; a.wb_max = wb; a.ncols_max = ncmax;
    a.axis_stride = (int)axis; a.debug_stride = dbg_stride;
    a.Lmax = (int)Lmax; a.fftN = fftN; a.union_bytes = (int)ub;
    // a mask that does not fit keeps its compact log image in global memory (L2)
    a.logI_cap = (int)std::min<long long>(std::max<long long>(cap, 0), vmax);
    const int smem = (int)(ub + fixed + 4LL * a.logI_cap);
    if (!vbrt::allow_smem(n4_subject_kernel, smem)) return VB_ERR_CUDA;
    bool launched = false;
    if (K > 1) {
        launched = vbrt::allow_smem(n4_cluster_kernel, smem) &&
                   VB_LAUNCH_CLUSTER(n4_cluster_kernel, dim3(grid), dim3(VB_N4_THREADS), (size_t)smem, s, K, a);
        if (!launched) { vbrt::clear_stale_error(); a.K = 1; }          // no room for such a cluster: one CTA per subject
    }
    if (!launched) { VB_LAUNCH(n4_subject_kernel, dim3(std::min(n, c->sms)), dim3(VB_N4_THREADS), (size_t)smem, s, a); }
    c->n4_cluster = launched ? K : 1;
    ++c->launches;
    return vbrt::last_launch_ok("n4_subject_kernel") ? VB_OK : VB_ERR_CUDA;
}

int run_border(vb_ctx* c, const uint8_t* a, int n, uint8_t* border, vbrt::stream_t s) {
    vbrt::clear_stale_error();
    const int gx = grid_x_for(c, n, vmax_of(c), 256);
    VB_LAUNCH(border_kernel, dim3(gx, n), dim3(256), 0, s, a, subj_ptr(c), border);
    ++c->launches;
    return vbrt::last_launch_ok("border_kernel") ? VB_OK : VB_ERR_CUDA;
}

int run_vdp(vb_ctx* c, const float* n4, const uint8_t* mask, int n, float thresh, uint8_t* defect, uint8_t* lb,
            uint8_t* dborder, vb_subject_stats* stats, vbrt::stream_t s) {
    vbrt::clear_stale_error();
    const VbSubject* subj = subj_ptr(c);
    {
        // sort buckets + the keys of a subject in shared memory (one CTA per SM); the bucketed copy
        // and the fallback for masks that do not fit live in a per-CTA scratch in global memory (L2)
        const long long vmax = vmax_of(c);
        const long long statics = vbrt::static_smem(vdp_stats_kernel);
        // words: 2^11 coarse + (2^11 + key_cap / LOAD) fine buckets + key_cap keys
        const long long avail = ((long long)c->max_smem - statics) / 4 - (2 << VB_VDP_LOG_COARSE);
        long long key_cap = std::max<long long>(64, std::min<long long>(vmax, avail * VB_VDP_LOAD / (VB_VDP_LOAD + 1))) & ~31LL;
        if (const char* e = getenv("VB_VDP_KEY_CAP"))          // tests: force the global-memory path
            key_cap = std::max<long long>(64, std::min<long long>(key_cap, atoll(e))) & ~31LL;
        const long long dyn_words = (2 << VB_VDP_LOG_COARSE) + key_cap / VB_VDP_LOAD + key_cap;
        const int grid = std::min(n, c->sms * 2);
        const size_t stride = (size_t)((vmax + 3) & ~3LL);
        Buf& ka = c->ws_keys_a[c->ksi & 1];
        Buf& kb = c->ws_keys_b[c->ksi & 1];
        if (!ensure(ka, 4 * stride * (size_t)grid) || !ensure(kb, 4 * stride * (size_t)grid)) return VB_ERR_CUDA;
        if (!vbrt::allow_smem(vdp_stats_kernel, (int)(4 * dyn_words))) return VB_ERR_CUDA;
        VB_LAUNCH(vdp_stats_kernel, dim3(grid), dim3(VB_VDP_THREADS), (size_t)(4 * dyn_words), s, n4, mask,
                  subj, n, stats, (int)key_cap, (int)dyn_words, (unsigned*)ka.p, (unsigned*)kb.p, (long long)stride);
        ++c->launches;
    }
    // fused threshold + linear binning + median + border on row tiles held in shared memory
    long long ncmax = 1, n0max = 1, n2max = 1;
    for (const VbSubject& S : c->last_subj) {
        ncmax = std::max<long long>(ncmax, (long long)S.n1 * S.n2);
        n0max = std::max<long long>(n0max, S.n0);
        n2max = std::max<long long>(n2max, S.n2);
    }
    const long long P = (n2max + 3) / 4 + 1;                   // padding words around the tiles (word version)
    const long long budget = std::min<long long>(c->max_smem - 1024, 96 * 1024) - 16 * P;
    long long R = (budget / ncmax - 6) / 2;                    // (R + 4) + (R + 2) rows of bytes
    if (R >= 1) {
        R = std::min<long long>(R, 32);
        R = std::min<long long>(R, n0max);
        const int smem = (int)((2 * R + 6) * ncmax + 16 * P);
        const int gx = (int)((n0max + R - 1) / R);
        bool any_words = false, any_bytes = false;             // subjects of this range per kernel flavour
        for (int i = 0; i < n; ++i) {
            const VbSubject& S = c->last_subj[(size_t)(c->r0 + i)];
            const bool w = (((long long)S.n1 * S.n2) & 3) == 0 && (S.off & 3) == 0 && (S.n2 + 3) / 4 + 1 <= P;
            (w ? any_words : any_bytes) = true;
        }
        if (any_words) {
            if (!vbrt::allow_smem(vdp_classify4_kernel, smem)) return VB_ERR_CUDA;
            VB_LAUNCH(vdp_classify4_kernel, dim3(gx, n), dim3(VB_CLS_THREADS), (size_t)smem, s, n4, mask, subj, thresh, (int)R,
                      (int)P, lb, defect, dborder, stats);
            ++c->launches;
        }
        if (any_bytes) {
            if (!vbrt::allow_smem(vdp_classify_kernel, smem)) return VB_ERR_CUDA;
            VB_LAUNCH(vdp_classify_kernel, dim3(gx, n), dim3(VB_CLS_THREADS), (size_t)smem, s, n4, mask, subj, thresh, (int)R,
                      (int)P, lb, defect, dborder, stats);
            ++c->launches;
        }
        return vbrt::last_launch_ok("vdp kernels") ? VB_OK : VB_ERR_CUDA;
    }
    // rows too long for a shared-memory tile: three streaming kernels through a d0 array
    if (!ensure(c->ws_d0, (size_t)c->v_total)) return VB_ERR_CUDA;
    const int gx = grid_x_for(c, n, vmax_of(c), 256);
    VB_LAUNCH(vdp_threshold_kernel, dim3(gx, n), dim3(256), 0, s, n4, mask, subj, thresh, (uint8_t*)c->ws_d0.p, lb,
              stats);
    VB_LAUNCH(vdp_median_kernel, dim3(gx, n), dim3(256), 0, s, (const uint8_t*)c->ws_d0.p, subj, defect, stats);
    c->launches += 2;
    if (!vbrt::last_launch_ok("vdp kernels")) return VB_ERR_CUDA;
    return dborder ? run_border(c, defect, n, dborder, s) : VB_OK;
}

int run_ci(vb_ctx* c, const uint8_t* defect, int n, uint16_t* shell, vb_subject_stats* stats, vbrt::stream_t s) {
    vbrt::clear_stale_error();
    int rc = upload_tables(c, s);
    if (rc) return rc;
    if (!ensure(c->ws_bits, 4 * (size_t)c->words_total) || !ensure(c->ws_list, 4 * (size_t)c->v_total) ||
        !ensure(c->ws_count, 4 * c->last_subj.size()))
        return VB_ERR_CUDA;
    const long long voff = c->last_subj[(size_t)c->r0].off;
    const long long vend = c->last_subj[(size_t)(c->r0 + n - 1)].off + c->last_subj[(size_t)(c->r0 + n - 1)].V;
    if (!vbrt::dev_memset(shell + voff, 0xFF, 2 * (size_t)(vend - voff), s) ||
        !vbrt::dev_memset((int*)c->ws_count.p + c->r0, 0, 4 * (size_t)n, s))
        return VB_ERR_CUDA;
    CiArgs a;
    a.defect = defect; a.subj = subj_ptr(c); a.n = n;
    a.tables = (const CiTableDev*)c->d_tables.p; a.bits = (uint32_t*)c->ws_bits.p;
    a.word_off = (const long long*)c->d_wordoff.p + c->r0; a.list = (int*)c->ws_list.p; a.count = (int*)c->ws_count.p + c->r0;
    a.shell = shell; a.stats = stats;
    const long long vmax = vmax_of(c);
    const int words = (int)((vmax + 31) / 32);
    int nbmax = 0;
    for (const VbSubject& S : c->last_subj) nbmax = std::max(nbmax, c->tables[(size_t)S.table].n_bounds);
    VB_LAUNCH(ci_pack_kernel, dim3(grid_x_for(c, n, words, 256), n), dim3(256), 0, s, a);
    const int gx = std::max(1, std::min(64, (c->sms * 8 + n - 1) / n));
    if (words <= VB_CI_SMEM_WORDS_MAX) {
        if (!vbrt::allow_smem(ci_grow_kernel<true>, 4 * words)) return VB_ERR_CUDA;
        VB_LAUNCH(ci_grow_kernel<true>, dim3(gx, n), dim3(VB_CI_THREADS), (size_t)4 * words, s, a);
    } else {
        VB_LAUNCH(ci_grow_kernel<false>, dim3(gx, n), dim3(VB_CI_THREADS), 0, s, a);
    }
    if (!vbrt::allow_smem(ci_percentile_kernel, 4 * nbmax)) return VB_ERR_CUDA;
    VB_LAUNCH(ci_percentile_kernel, dim3(std::min(n, c->sms * 4)), dim3(VB_CI_THREADS), (size_t)4 * nbmax, s, a);
    c->launches += 3;
    return vbrt::last_launch_ok("ci kernels") ? VB_OK : VB_ERR_CUDA;
}

}  // namespace

// ============================================================================================ C ABI
extern "C" {

int vb_version(void) { return 100; }
const char* vb_last_error(void) { return vbrt::g_err.c_str(); }
int vb_device_count(void) { return vbrt::device_count(); }

int vb_create(int device, vb_ctx** out) {
    if (!out) return fail(VB_ERR_ARG, "null ctx pointer");
    *out = nullptr;
    const int nd = vbrt::device_count();
    if (nd <= 0) return fail(VB_ERR_CUDA, "no CUDA device: ventb200 has no CPU fallback");
    if (device < 0 || device >= nd) return fail(VB_ERR_ARG, "bad device index");
    if (!vbrt::set_device(device)) return VB_ERR_CUDA;
    vb_ctx* c = new vb_ctx();
    c->device = device;
    c->sms = vbrt::sm_count(device);
    c->max_smem = vbrt::max_smem_optin(device);
    if (!vbrt::stream_create(&c->stream) || !vbrt::stream_create(&c->stream_in) || !vbrt::stream_create(&c->stream_out) ||
        !vbrt::stream_create(&c->stream2) || !vbrt::stream_create_high(&c->post[0]) || !vbrt::stream_create_high(&c->post[1]) ||
        !vbrt::stream_create_high(&c->pre[0]) || !vbrt::stream_create_high(&c->pre[1]) ||
        !vbrt::stream_create(&c->cextra[0]) || !vbrt::stream_create(&c->cextra[1])) {
        delete c;
        return VB_ERR_CUDA;
    }
    if (const char* e = getenv("VB_N4_STREAMS")) c->n_cstreams = std::max(1, std::min(VB_MAX_CSTREAMS, atoi(e)));   // tuning
    *out = c;
    return VB_OK;
}

void vb_destroy(vb_ctx* c) {
    if (!c) return;
    vbrt::set_device(c->device);
    {   // every stream of the context is idle before anything is freed
        vbrt::stream_t all[] = {c->stream, c->stream2, c->cextra[0], c->cextra[1], c->pre[0], c->pre[1], c->post[0], c->post[1],
                                c->stream_in, c->stream_out};
        for (vbrt::stream_t st : all) vbrt::stream_sync(st);
    }
    for (N4Ws& W : c->n4ws) {
        Buf* wb[] = {&W.logI, &W.u, &W.colmask, &W.rbb, &W.wbase, &W.acol, &W.order, &W.ec, &W.code, &W.axis, &W.counter, &W.cl_hist, &W.cl_q, &W.cl_red, &W.cl_bar, &W.cl_bcast};
        for (Buf* b : wb) vbrt::dev_free(b->p);
    }
    vbrt::stream_destroy(c->stream2);
    vbrt::stream_destroy(c->post[0]); vbrt::stream_destroy(c->post[1]);
    for (vbrt::event_t e : c->ev_post) vbrt::event_destroy(e);
    vbrt::stream_destroy(c->pre[0]); vbrt::stream_destroy(c->pre[1]);
    vbrt::stream_destroy(c->cextra[0]); vbrt::stream_destroy(c->cextra[1]);
    for (vbrt::event_t e : c->ev_pre) vbrt::event_destroy(e);
    Buf* bufs[] = {&c->d_subj, &c->d_wordoff, &c->ws_d0, &c->ws_keys_a[0], &c->ws_keys_a[1], &c->ws_keys_b[0], &c->ws_keys_b[1],
                   &c->ws_bits, &c->ws_list, &c->ws_count, &c->dbg, &c->d_tables, &c->m_image, &c->m_mask, &c->m_n4,
                   &c->m_logbias, &c->m_defect, &c->m_lb, &c->m_dborder, &c->m_mborder, &c->m_shell, &c->m_stats, &c->m_proton, &c->m_pack, &c->d_voxmin,
                   &c->m_u16, &c->m_bits, &c->m_codes, &c->m_ci_idx, &c->m_ci_val, &c->d_bitoff,
                   &c->m_k[0], &c->m_k[1], &c->d_ktmp,
                   &c->d_ovl, &c->m_rgb, &c->m_ovl_in[0], &c->m_ovl_in[1], &c->m_ovl_in[2], &c->m_ovl_in[3], &c->m_ovl_in[4], &c->m_ovl_in[5], &c->d_parula, &c->d_outoff};
    for (Buf* b : bufs) vbrt::dev_free(b->p);
    for (TableHost& t : c->tables) { vbrt::dev_free(t.d_rows); vbrt::dev_free(t.d_ii); vbrt::dev_free(t.d_up); vbrt::dev_free(t.d_radii); }
    vbrt::host_free(c->h_desc.p);
    vbrt::host_free(c->h_stats.p);
    for (vbrt::event_t e : c->ev) vbrt::event_destroy(e);
    for (vbrt::event_t e : c->ev_copy) vbrt::event_destroy(e);
    vbrt::event_destroy(c->ev_fork);
    for (vbrt::event_t e : c->ev_joins) vbrt::event_destroy(e);
    vbrt::stream_destroy(c->stream_in);
    vbrt::stream_destroy(c->stream_out);
    vbrt::stream_destroy(c->stream);
    delete c;
}

uint64_t vb_launch_count(const vb_ctx* c) { return c ? c->launches : 0; }

void vb_n4_default_params(vb_n4_params* p) {
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->n_levels = 4;
    for (int i = 0; i < 4; ++i) p->max_iter[i] = 50;
    p->conv_threshold = 0.001f; p->fwhm = 0.15f; p->wiener_noise = 0.01f; p->n_bins = 200;
    p->ncp[0] = p->ncp[1] = p->ncp[2] = 4;
    p->mask_label = 1; p->max_total_iters = 0;
}

int vb_ci_table_create(vb_ctx* c, const double* table, int64_t nrows, int32_t n0, int32_t n1, int32_t n2,
                       int32_t* handle) {
    if (!c || !table || nrows < 2 || !handle || n0 < 1 || n1 < 1 || n2 < 1) return fail(VB_ERR_ARG, "ci_table_create");
    vbrt::set_device(c->device);
    // shell boundaries: getRadiiIndices (CI.py:79-85) == where(diff(radius) > 0) + 1
    std::vector<int> bnd;
    for (int64_t r = 0; r + 1 < nrows; ++r)
        if (table[4 * (r + 1)] - table[4 * r] > 0) bnd.push_back((int)(r + 1));
    if (bnd.empty()) return fail(VB_ERR_ARG, "sphere table has no shell boundary");
    if (bnd.size() >= 0xFFFE) return fail(VB_ERR_UNSUPPORTED, "more than 65533 shells");
    const int last = bnd.back();                 // rows >= last are never looked at (CI.py:94)
    std::vector<int2> rows;
    rows.reserve((size_t)last);
    std::unordered_set<long long> seen;
    seen.reserve((size_t)last * 2);
    std::vector<int> ucount(bnd.size());
    size_t bi = 0;
    for (int r = 0; r < last; ++r) {
        while (bi < bnd.size() && bnd[bi] == r) ucount[bi++] = (int)rows.size();
        const long long off = (long long)table[4 * r + 1] + (long long)table[4 * r + 2] * n0 +
                              (long long)table[4 * r + 3] * n0 * (long long)n1;
        if (off > 0x3fffffffLL || off < -0x3fffffffLL) continue;   // can never land inside a <=2^30 volume
        if (seen.insert(off).second) { int2 e; e.x = (int)off; e.y = 0; rows.push_back(e); }
    }
    while (bi < bnd.size()) ucount[bi++] = (int)rows.size();
    for (size_t t = 0; t < bnd.size(); ++t) {
        const int u = ucount[t] - 1;             // >= 0: row 0 (the centre) is always unique
        if (u < 0) return fail(VB_ERR_ARG, "sphere table must start with the centre voxel");
        const int cnt = rows[(size_t)u].y & 0xFFFF;
        if (cnt == 0) rows[(size_t)u].y = (int)(((unsigned)t << 16) | 1u);
        else rows[(size_t)u].y = (int)(((unsigned)rows[(size_t)u].y & 0xFFFF0000u) | (unsigned)(cnt + 1));
    }
    TableHost T;
    T.n0 = n0; T.n1 = n1; T.n2 = n2;
    T.n_unique = (int)rows.size(); T.n_bounds = (int)bnd.size();
    T.radii.resize(bnd.size());
    for (size_t t = 0; t < bnd.size(); ++t) T.radii[t] = table[4 * (size_t)(bnd[t] - 1)];   // CI.py:105
    std::vector<int> up((rows.size() + 31) / 32, 0);
    for (size_t u = 0; u < rows.size(); ++u) {
        const int cnt = rows[u].y & 0xFFFF;
        if (cnt) up[u >> 5] = std::max(up[u >> 5], bnd[(size_t)((unsigned)rows[u].y >> 16) + (size_t)cnt - 1]);
    }
    if (!vbrt::dev_malloc(&T.d_rows, rows.size() * sizeof(int2)) || !vbrt::dev_malloc(&T.d_ii, bnd.size() * sizeof(int)) ||
        !vbrt::dev_malloc(&T.d_up, up.size() * sizeof(int)) || !vbrt::dev_malloc(&T.d_radii, bnd.size() * sizeof(double)))
        return VB_ERR_CUDA;
    if (!vbrt::h2d(T.d_rows, rows.data(), rows.size() * sizeof(int2), c->stream) ||
        !vbrt::h2d(T.d_ii, bnd.data(), bnd.size() * sizeof(int), c->stream) ||
        !vbrt::h2d(T.d_up, up.data(), up.size() * sizeof(int), c->stream) ||
        !vbrt::h2d(T.d_radii, T.radii.data(), bnd.size() * sizeof(double), c->stream) || !vbrt::stream_sync(c->stream))
        return VB_ERR_CUDA;
    c->tables.push_back(T);
    c->tables_dirty = true;
    *handle = (int32_t)c->tables.size() - 1;
    return VB_OK;
}

int vb_ci_table_radii(vb_ctx* c, int32_t handle, double* radii, int32_t capacity, int32_t* n_bounds) {
    if (!c || handle < 0 || handle >= (int)c->tables.size()) return fail(VB_ERR_ARG, "bad CI table handle");
    const TableHost& T = c->tables[(size_t)handle];
    if (n_bounds) *n_bounds = T.n_bounds;
    if (radii)
        for (int t = 0; t < T.n_bounds && t < capacity; ++t) radii[t] = T.radii[(size_t)t];
    return VB_OK;
}

int vb_snr_batch_dev(vb_ctx* c, const float* image, const uint8_t* mask, int32_t n, const int32_t* shapes,
                     int32_t fov, vb_subject_stats* stats, void* stream) {
    if (!c || !image || !mask || !stats) return fail(VB_ERR_ARG, "snr: null argument");
    vbrt::set_device(c->device);
    int rc = prepare(c, n, shapes, nullptr, (vbrt::stream_t)stream);
    return rc ? rc : run_snr(c, image, mask, n, fov, stats, (vbrt::stream_t)stream);
}

int vb_n4_batch_dev(vb_ctx* c, const float* image, const uint8_t* mask, int32_t n, const int32_t* shapes,
                    const vb_n4_params* params, float* n4, float* logbias, vb_subject_stats* stats, void* stream) {
    if (!c || !image || !mask || !n4 || !params) return fail(VB_ERR_ARG, "n4: null argument");
    vbrt::set_device(c->device);
    int rc = prepare(c, n, shapes, nullptr, (vbrt::stream_t)stream);
    return rc ? rc : run_n4(c, image, mask, n, params, n4, logbias, stats, (vbrt::stream_t)stream);
}

int vb_vdp_batch_dev(vb_ctx* c, const float* n4, const uint8_t* mask, int32_t n, const int32_t* shapes, float thresh,
                     uint8_t* defect, uint8_t* lb, uint8_t* dborder, vb_subject_stats* stats, void* stream) {
    if (!c || !n4 || !mask || !defect || !lb || !stats) return fail(VB_ERR_ARG, "vdp: null argument");
    vbrt::set_device(c->device);
    int rc = prepare(c, n, shapes, nullptr, (vbrt::stream_t)stream);
    return rc ? rc : run_vdp(c, n4, mask, n, thresh, defect, lb, dborder, stats, (vbrt::stream_t)stream);
}

int vb_border_batch_dev(vb_ctx* c, const uint8_t* a, int32_t n, const int32_t* shapes, uint8_t* border, void* stream) {
    if (!c || !a || !border) return fail(VB_ERR_ARG, "border: null argument");
    vbrt::set_device(c->device);
    int rc = prepare(c, n, shapes, nullptr, (vbrt::stream_t)stream);
    return rc ? rc : run_border(c, a, n, border, (vbrt::stream_t)stream);
}

int vb_ci_batch_dev(vb_ctx* c, const uint8_t* defect, int32_t n, const int32_t* shapes, const int32_t* handles,
                    uint16_t* shell, vb_subject_stats* stats, void* stream) {
    if (!c || !defect || !handles || !shell || !stats) return fail(VB_ERR_ARG, "ci: null argument");
    vbrt::set_device(c->device);
    int rc = prepare(c, n, shapes, handles, (vbrt::stream_t)stream);
    return rc ? rc : run_ci(c, defect, n, shell, stats, (vbrt::stream_t)stream);
}

// Vent_Analysis.build4DdataArray for a batch (device pointers; null inputs leave their channel 0)
int vb_pack6_batch_dev(vb_ctx* c, int32_t n, const int32_t* shapes, const int32_t* handles, const double* vox_min,
                       const float* proton, const float* image, const uint8_t* mask, const float* n4,
                       const uint8_t* defect, const uint16_t* shell, float* out, void* stream) {
    if (!c || !vox_min || !out) return fail(VB_ERR_ARG, "pack6: null argument");
    if (shell && !handles) return fail(VB_ERR_ARG, "pack6: ci_shell needs the CI table handles");
    vbrt::set_device(c->device);
    vbrt::stream_t s = (vbrt::stream_t)stream;
    int rc = prepare(c, n, shapes, shell ? handles : nullptr, s);
    if (rc) return rc;
    if (shell && (rc = upload_tables(c, s))) return rc;
    if (!ensure(c->d_voxmin, sizeof(double) * (size_t)n)) return VB_ERR_CUDA;
    if (!vbrt::stream_sync(s)) return VB_ERR_CUDA;             // a previous upload from the caller's array
    if (!vbrt::h2d(c->d_voxmin.p, vox_min, sizeof(double) * (size_t)n, s) || !vbrt::stream_sync(s)) return VB_ERR_CUDA;
    vbrt::clear_stale_error();
    c->r0 = 0;
    Pack6Args a;
    a.proton = proton; a.image = image; a.mask = mask; a.n4 = n4; a.defect = defect; a.shell = shell;
    a.subj = subj_ptr(c); a.tables = (const CiTableDev*)c->d_tables.p; a.vox_min = (const double*)c->d_voxmin.p; a.out = out;
    VB_LAUNCH(pack6_kernel, dim3(grid_x_for(c, n, vmax_of(c), 256), n), dim3(256), 0, s, a);
    ++c->launches;
    return vbrt::last_launch_ok("pack6_kernel") ? VB_OK : VB_ERR_CUDA;
}

// the same on host arrays (pageable or pinned); blocks until `out` is filled
int vb_pack6_batch_host(vb_ctx* c, int32_t n, const int32_t* shapes, const int32_t* handles, const double* vox_min,
                        const float* proton, const float* image, const uint8_t* mask, const float* n4,
                        const uint8_t* defect, const uint16_t* shell, float* out) {
    if (!c || !shapes || n <= 0 || !out) return fail(VB_ERR_ARG, "pack6: null argument");
    vbrt::set_device(c->device);
    vbrt::stream_t s = c->stream;
    size_t V = 0;
    for (int i = 0; i < n; ++i) V += (size_t)shapes[3 * i] * shapes[3 * i + 1] * shapes[3 * i + 2];
    bool ok = ensure(c->m_pack, 24 * V);
    if (proton) ok = ok && ensure(c->m_proton, 4 * V) && vbrt::h2d(c->m_proton.p, proton, 4 * V, s);
    if (image) ok = ok && ensure(c->m_image, 4 * V) && vbrt::h2d(c->m_image.p, image, 4 * V, s);
    if (mask) ok = ok && ensure(c->m_mask, V) && vbrt::h2d(c->m_mask.p, mask, V, s);
    if (n4) ok = ok && ensure(c->m_n4, 4 * V) && vbrt::h2d(c->m_n4.p, n4, 4 * V, s);
    if (defect) ok = ok && ensure(c->m_defect, V) && vbrt::h2d(c->m_defect.p, defect, V, s);
    if (shell) ok = ok && ensure(c->m_shell, 2 * V) && vbrt::h2d(c->m_shell.p, shell, 2 * V, s);
    if (!ok) return VB_ERR_CUDA;
    int rc = vb_pack6_batch_dev(c, n, shapes, handles, vox_min, proton ? (const float*)c->m_proton.p : nullptr,
                                image ? (const float*)c->m_image.p : nullptr, mask ? (const uint8_t*)c->m_mask.p : nullptr,
                                n4 ? (const float*)c->m_n4.p : nullptr, defect ? (const uint8_t*)c->m_defect.p : nullptr,
                                shell ? (const uint16_t*)c->m_shell.p : nullptr, (float*)c->m_pack.p, s);
    if (rc) return rc;
    if (!vbrt::d2h(out, c->m_pack.p, 24 * V, s) || !vbrt::stream_sync(s)) return VB_ERR_CUDA;
    return VB_OK;
}

// the selected stages on subjects [r0, r0 + cnt) of the prepared batch (buffers are whole-batch)
static int pipeline_range(vb_ctx* c, int r0, int cnt, const vb_n4_params* params, float thresh, uint32_t flags,
                          const vb_buffers* d, vbrt::stream_t s, vbrt::event_t* ev, bool clear_stats = true) {
    int rc = VB_OK;
    c->r0 = r0;
    vb_subject_stats* st = d->stats + r0;
    do {
        if (clear_stats && !vbrt::dev_memset(st, 0, sizeof(vb_subject_stats) * (size_t)cnt, s)) { rc = VB_ERR_CUDA; break; }
        if (ev && !vbrt::event_record(ev[0], s)) { rc = VB_ERR_CUDA; break; }
        if (flags & VB_DO_SNR)
            if ((rc = run_snr(c, d->image, d->mask, cnt, 20, st, s))) break;
        if (ev && !vbrt::event_record(ev[1], s)) { rc = VB_ERR_CUDA; break; }
        if (flags & VB_DO_N4)
            if ((rc = run_n4(c, d->image, d->mask, cnt, params, d->n4, d->logbias, st, s))) break;
        if (ev && !vbrt::event_record(ev[2], s)) { rc = VB_ERR_CUDA; break; }
        if (flags & VB_DO_VDP)
            if ((rc = run_vdp(c, d->n4, d->mask, cnt, thresh, d->defect, d->defect_lb, d->defect_border, st, s))) break;
        if (ev && !vbrt::event_record(ev[3], s)) { rc = VB_ERR_CUDA; break; }
        if ((flags & VB_DO_MASK_BORDER) && d->mask_border)
            if ((rc = run_border(c, d->mask, cnt, d->mask_border, s))) break;
        if (ev && !vbrt::event_record(ev[4], s)) { rc = VB_ERR_CUDA; break; }
        if (flags & VB_DO_CI)
            if ((rc = run_ci(c, d->defect, cnt, d->ci_shell, st, s))) break;
        if (ev && !vbrt::event_record(ev[5], s)) { rc = VB_ERR_CUDA; break; }
    } while (0);
    c->r0 = 0;
    return rc;
}

// Chunk k of a host pipeline (or part k of a large resident batch), on three streams:
//   `pre`  the chunk's inputs are ready there (the caller has queued the unpacking); SNR runs there too.
//          The host pipelines pass a HIGH-PRIORITY stream: these small kernels would otherwise sit in
//          front of the chunk's N4 launch on the compute stream and could only start once SMs held by
//          the other chunk's N4 CTAs retire -- N4 of chunk k then waited for a trickle of SNR CTAs.
//   `cs`   the persistent N4 kernel; consecutive chunks alternate between two compute streams (and N4
//          workspace sets), so the stragglers of one chunk's N4 overlap the next chunk's CTAs.
//   *ps    (returned) the stages after N4 on a high-priority stream of their own: N4 of chunk k + 2
//          only waits for N4 of chunk k, not for chunk k's VDP / CI kernels, and those get a retiring
//          SM before the next N4 CTA does.  The chunk's outputs are complete on *ps.
static bool chunk_edge(std::vector<vbrt::event_t>& evs, int k, vbrt::stream_t from, vbrt::stream_t to) {
    if (from == to) return true;
    while (evs.size() <= (size_t)k) {
        vbrt::event_t e;
        if (!vbrt::event_create(&e)) return false;
        evs.push_back(e);
    }
    return vbrt::event_record(evs[(size_t)k], from) && vbrt::stream_wait(to, evs[(size_t)k]);
}
static int chunk_stages(vb_ctx* c, int k, int r0, int cnt, const vb_n4_params* params, float thresh, uint32_t flags,
                        const vb_buffers* d, vbrt::stream_t pre, vbrt::stream_t cs, int ws, vbrt::stream_t* ps, bool clear_stats) {
    const uint32_t f_pre = flags & (uint32_t)VB_DO_SNR, f_n4 = flags & (uint32_t)VB_DO_N4, f_post = flags & ~(f_pre | f_n4);
    const bool split = f_n4 && !getenv("VB_NO_POST_STREAMS");
    *ps = cs;
    int rc = VB_OK;
    c->wsi = ws; c->ksi = k & 1;
    if (!split) {                                            // no N4 kernel to work around: one stream
        if (!chunk_edge(c->ev_pre, k, pre, cs)) rc = VB_ERR_CUDA;
        if (!rc) rc = pipeline_range(c, r0, cnt, params, thresh, flags, d, cs, nullptr, clear_stats);
        c->wsi = 0; c->ksi = 0;
        return rc;
    }
    rc = pipeline_range(c, r0, cnt, params, thresh, f_pre, d, pre, nullptr, clear_stats);       // (clears the records,) SNR
    if (!rc && !chunk_edge(c->ev_pre, k, pre, cs)) rc = VB_ERR_CUDA;
    c->wsi = ws; c->ksi = k & 1;
    if (!rc) rc = pipeline_range(c, r0, cnt, params, thresh, f_n4, d, cs, nullptr, false);
    if (!rc && f_post) {
        *ps = c->post[k & 1];
        if (!chunk_edge(c->ev_post, k, cs, *ps)) rc = VB_ERR_CUDA;
        c->wsi = ws; c->ksi = k & 1;
        if (!rc) rc = pipeline_range(c, r0, cnt, params, thresh, f_post, d, *ps, nullptr, false);
    }
    c->wsi = 0; c->ksi = 0;
    return rc;
}

// marks the N4 launches of a schedule of `parts` chunks (run_n4 keeps one CTA per subject when several N4 kernels share the SMs)
struct PartsScope {
    vb_ctx* c;
    PartsScope(vb_ctx* ctx, int parts) : c(ctx) { c->n4_parts = parts; }
    ~PartsScope() { c->n4_parts = 1; }
};

static int pipeline_check(uint32_t flags, const int32_t* handles, const vb_n4_params* params, const vb_buffers* d) {
    if ((flags & VB_DO_CI) && !handles) return fail(VB_ERR_ARG, "pipeline: CI needs table handles");
    if ((flags & VB_DO_SNR) && !d->image) return fail(VB_ERR_ARG, "pipeline: SNR needs image");
    if ((flags & VB_DO_N4) && (!d->image || !d->n4 || !params)) return fail(VB_ERR_ARG, "pipeline: N4 needs image, n4, params");
    if ((flags & VB_DO_VDP) && (!d->n4 || !d->defect || !d->defect_lb))
        return fail(VB_ERR_ARG, "pipeline: VDP needs n4, defect, defect_lb");
    if ((flags & VB_DO_CI) && (!d->defect || !d->ci_shell)) return fail(VB_ERR_ARG, "pipeline: CI needs defect, ci_shell");
    return VB_OK;
}

int vb_pipeline_batch_dev(vb_ctx* c, int32_t n, const int32_t* shapes, const int32_t* handles,
                          const vb_n4_params* params, float thresh, uint32_t flags, const vb_buffers* d, void* stream) {
    if (!c || !d || !d->mask || !d->stats) return fail(VB_ERR_ARG, "pipeline: null argument");
    vbrt::stream_t s = (vbrt::stream_t)stream;
    vbrt::set_device(c->device);
    int rc = pipeline_check(flags, handles, params, d);
    if (rc) return rc;
    rc = prepare(c, n, shapes, (flags & VB_DO_CI) ? handles : nullptr, s);
    if (rc) return rc;
    // stage boundaries: one event before the first stage and one after each stage
    vbrt::event_t* ev = nullptr;
    if (c->prof) {
        while (c->ev.size() < c->ev_used + VB_N_STAGES + 1) {
            vbrt::event_t e;
            if (!vbrt::event_create(&e)) return VB_ERR_CUDA;
            c->ev.push_back(e);
        }
        ev = c->ev.data() + c->ev_used;
        c->ev_used += VB_N_STAGES + 1;
    }
    // Large batches run in parts on two streams: while the persistent N4 kernel of part k + 1 holds
    // the SMs, the CTAs of part k that run out of subjects free theirs one by one -- part k + 1's
    // CTAs and the (small) VDP / CI kernels of the earlier parts fill that tail instead of leaving
    // it idle, and only the last part's VDP / CI are exposed.  SNR of the whole batch goes first.
    int nparts = (n >= 6 * std::max(c->sms, 1)) ? 2 : 1;      // measured on 1776 C2 subjects: 2 parts 85.6 ms, 1: 87.6, 4: 86.4, 6: 86.6
    if (const char* e = getenv("VB_PARTS")) nparts = std::max(1, std::min(8, atoi(e)));
    if (!(flags & VB_DO_N4) || c->debug) nparts = 1;
    if (nparts == 1) return pipeline_range(c, 0, n, params, thresh, flags, d, s, ev);
    const PartsScope parts_scope(c, nparts);
    if (!c->ev_fork) {
        if (!vbrt::event_create(&c->ev_fork)) return VB_ERR_CUDA;
    }
    if (!vbrt::dev_memset(d->stats, 0, sizeof(vb_subject_stats) * (size_t)n, s)) return VB_ERR_CUDA;
    if (ev && !vbrt::event_record(ev[0], s)) return VB_ERR_CUDA;
    c->r0 = 0;
    if ((flags & VB_DO_SNR) && (rc = run_snr(c, d->image, d->mask, n, 20, d->stats, s))) return rc;
    if (ev && !vbrt::event_record(ev[1], s)) return VB_ERR_CUDA;
    if (!vbrt::event_record(c->ev_fork, s) || !vbrt::stream_wait(c->stream2, c->ev_fork)) return VB_ERR_CUDA;
    for (int k = 0; k < nparts; ++k) {
        // two parts: the last one is a single wave of subjects -- its VDP / CI kernels are the only ones
        // nothing overlaps
        int r0 = (int)((long long)n * k / nparts), r1 = (int)((long long)n * (k + 1) / nparts);
        // (measured on 1776 C2 subjects, first part 50 / 67 / 75 / 84 / 92 %: 74.5 / 73.9 / 73.5 / 73.2 / 73.0 ms)
        const int sm = std::max(c->sms, 1);
        if (nparts == 2 && n >= 2 * sm) {
            int cut = n - sm;                                  // last part: one wave
            if (const char* e = getenv("VB_PART_SPLIT"))       // tuning: first part in per cent
                cut = std::max(sm, std::min(n - sm, (int)((long long)n * std::max(10, std::min(95, atoi(e))) / 100) / sm * sm));
            r0 = k == 0 ? 0 : cut; r1 = k == 0 ? cut : n;
        }
        vbrt::stream_t ps;
        vbrt::stream_t cs = (k & 1) ? c->stream2 : s;
        rc = chunk_stages(c, k, r0, r1 - r0, params, thresh, flags & ~(uint32_t)VB_DO_SNR, d, cs, cs, k & 1, &ps, false);
        if (rc) break;
    }
    // join: the caller's stream continues when all of them are done
    {
        vbrt::stream_t js[3] = {c->stream2, c->post[0], c->post[1]};
        for (int j = 0; j < 3; ++j) {
            if (!c->ev_joins[j] && !vbrt::event_create(&c->ev_joins[j])) return VB_ERR_CUDA;
            if (!vbrt::event_record(c->ev_joins[j], js[j]) || !vbrt::stream_wait(s, c->ev_joins[j])) return VB_ERR_CUDA;
        }
    }
    if (ev) {                                                  // the stages overlap: N4 = everything up to the join, the rest reads 0
        for (int k = 2; k <= VB_N_STAGES; ++k)
            if (!vbrt::event_record(ev[k], s)) return VB_ERR_CUDA;
    }
    return rc;
}

// Host pipelines work in chunks of a bit less than one wave of subjects: inputs of chunk k+1 go up
// and outputs of chunk k-1 come down (two copy streams) while chunk k computes; up to four chunks'
// N4 kernels are in flight (compute streams), so an SM that retires a subject always finds a CTA of
// a later chunk waiting.  Large batches start with two smaller chunks (1/4 and 1/2 wave): nothing
// overlaps the first upload.  (Small chunks at the END were measured slower: a chunk cannot take
// less than one subject's latency.)
static std::vector<int> chunk_bounds(const vb_ctx* c, int n) {
    int per = std::max(c->sms, 1);
    if (const char* e = getenv("VB_CHUNK_PER")) per = std::max(1, atoi(e));     // tuning: subjects per chunk
    std::vector<int> bounds;
    if (n >= 6 * per) {
        const int a = std::max(1, per / 4), b = std::max(1, per / 2);
        const int rest = n - (a + b), nm = std::max(1, std::min(14, 2 * rest / per));       // about 0.8 waves per chunk
        bounds.push_back(0); bounds.push_back(a); bounds.push_back(a + b);
        for (int k = 1; k <= nm; ++k) bounds.push_back(a + b + (int)((long long)rest * k / nm));
        {   // the last chunk in two halves: its download is the one nothing overlaps (measured 85.4 -> 84.3 ms
            // per 1776 C2 subjects; four quarters: 87.7)
            int t = 2;
            if (const char* e = getenv("VB_CHUNK_TAIL")) t = std::max(1, std::min(8, atoi(e)));      // tuning
            const int lo = bounds[bounds.size() - 2], hi = bounds.back();
            if (t > 1 && hi - lo >= 2 * t) {
                bounds.pop_back();
                for (int k = 1; k <= t; ++k) bounds.push_back(lo + (int)((long long)(hi - lo) * k / t));
            }
        }
    } else {
        const int nc = std::max(1, std::min(16, n / per));
        for (int k = 0; k <= nc; ++k) bounds.push_back((int)((long long)n * k / nc));
    }
    return bounds;
}
// every exit path of a host pipeline waits for the copies it has enqueued: the caller may free its
// buffers, and the next call reuses the device mirrors
static bool drain(vb_ctx* c) {
    bool ok = vbrt::stream_sync(c->stream_in);
    ok = vbrt::stream_sync(c->stream) && ok;
    ok = vbrt::stream_sync(c->stream2) && ok;
    ok = vbrt::stream_sync(c->post[0]) && ok;
    ok = vbrt::stream_sync(c->post[1]) && ok;
    ok = vbrt::stream_sync(c->pre[0]) && ok;
    ok = vbrt::stream_sync(c->pre[1]) && ok;
    ok = vbrt::stream_sync(c->cextra[0]) && ok;
    ok = vbrt::stream_sync(c->cextra[1]) && ok;
    ok = vbrt::stream_sync(c->stream_out) && ok;
    return ok;
}

int vb_pipeline_batch_host(vb_ctx* c, int32_t n, const int32_t* shapes, const int32_t* handles,
                           const vb_n4_params* params, float thresh, uint32_t flags, const vb_buffers* h) {
    if (!c || !h || !h->mask || !h->stats || !shapes || n <= 0) return fail(VB_ERR_ARG, "pipeline_host: null argument");
    vbrt::set_device(c->device);
    vbrt::stream_t s = c->stream;
    size_t V = 0;
    for (int i = 0; i < n; ++i) V += (size_t)shapes[3 * i] * shapes[3 * i + 1] * shapes[3 * i + 2];
    const bool so_only = (flags & VB_HOST_STATS_ONLY) != 0;     // per-subject records only: no array comes back
    const bool need_n4 = (flags & (VB_DO_N4 | VB_DO_VDP)) != 0;
    const bool need_def = (flags & (VB_DO_VDP | VB_DO_CI)) != 0;
    vb_buffers d;
    memset(&d, 0, sizeof(d));
    if (!ensure(c->m_mask, V) || !ensure(c->m_stats, sizeof(vb_subject_stats) * (size_t)n)) return VB_ERR_CUDA;
    d.mask = (const uint8_t*)c->m_mask.p; d.stats = (vb_subject_stats*)c->m_stats.p;
    if (h->image) {
        if (!ensure(c->m_image, 4 * V)) return VB_ERR_CUDA;
        d.image = (const float*)c->m_image.p;
    }
    if (need_n4) {
        if (!h->n4 && (!so_only || !(flags & VB_DO_N4))) return fail(VB_ERR_ARG, "pipeline_host: n4 buffer required");
        if (!ensure(c->m_n4, 4 * V)) return VB_ERR_CUDA;
        d.n4 = (float*)c->m_n4.p;
    }
    if (h->logbias && (flags & VB_DO_N4)) { if (!ensure(c->m_logbias, 4 * V)) return VB_ERR_CUDA; d.logbias = (float*)c->m_logbias.p; }
    if (need_def) {
        if (!h->defect && (!so_only || !(flags & VB_DO_VDP))) return fail(VB_ERR_ARG, "pipeline_host: defect buffer required");
        if (!ensure(c->m_defect, V)) return VB_ERR_CUDA;
        d.defect = (uint8_t*)c->m_defect.p;
    }
    if (flags & VB_DO_VDP) {
        if (!h->defect_lb && !so_only) return fail(VB_ERR_ARG, "pipeline_host: defect_lb buffer required");
        if (!ensure(c->m_lb, V)) return VB_ERR_CUDA;
        d.defect_lb = (uint8_t*)c->m_lb.p;
        if (h->defect_border) { if (!ensure(c->m_dborder, V)) return VB_ERR_CUDA; d.defect_border = (uint8_t*)c->m_dborder.p; }
    }
    if ((flags & VB_DO_MASK_BORDER) && h->mask_border) { if (!ensure(c->m_mborder, V)) return VB_ERR_CUDA; d.mask_border = (uint8_t*)c->m_mborder.p; }
    if (flags & VB_DO_CI) {
        if (!h->ci_shell && !so_only) return fail(VB_ERR_ARG, "pipeline_host: ci_shell buffer required");
        if (!ensure(c->m_shell, 2 * V)) return VB_ERR_CUDA;
        d.ci_shell = (uint16_t*)c->m_shell.p;
    }
    int rc = pipeline_check(flags, handles, params, &d);
    if (rc) return rc;
    rc = prepare(c, n, shapes, (flags & VB_DO_CI) ? handles : nullptr, s);
    if (rc) return rc;
    const std::vector<int> bounds = chunk_bounds(c, n);        // first subject of every chunk, then n
    const int nchunks = (int)bounds.size() - 1;
    const PartsScope parts_scope(c, nchunks);
    while (c->ev_copy.size() < 2 * (size_t)nchunks) {
        vbrt::event_t e;
        if (!vbrt::event_create(&e)) return VB_ERR_CUDA;
        c->ev_copy.push_back(e);
    }
    // the (small) stats records go through pinned staging: a copy into pageable user memory would
    // block the host until the chunk has finished and serialise the pipeline
    const size_t stb = sizeof(vb_subject_stats) * (size_t)n;
    if (c->h_stats.cap < stb) {
        vbrt::host_free(c->h_stats.p);
        c->h_stats.p = nullptr; c->h_stats.cap = 0;
        if (!vbrt::host_malloc(&c->h_stats.p, stb * 2)) return VB_ERR_CUDA;
        c->h_stats.cap = stb * 2;
    }
    bool ok = true;
    for (int k = 0; k < nchunks && ok; ++k) {
        const int r0 = bounds[(size_t)k], r1 = bounds[(size_t)k + 1];
        const size_t v0 = (size_t)c->last_subj[(size_t)r0].off;
        const size_t v1 = (size_t)c->last_subj[(size_t)(r1 - 1)].off + (size_t)c->last_subj[(size_t)(r1 - 1)].V;
        const size_t nv = v1 - v0;
        vbrt::stream_t si = c->stream_in;
        ok = ok && vbrt::h2d((uint8_t*)c->m_mask.p + v0, h->mask + v0, nv, si);
        if (h->image) ok = ok && vbrt::h2d((float*)c->m_image.p + v0, h->image + v0, 4 * nv, si);
        if (need_n4 && !(flags & VB_DO_N4)) ok = ok && vbrt::h2d((float*)c->m_n4.p + v0, h->n4 + v0, 4 * nv, si);
        if (need_def && !(flags & VB_DO_VDP)) ok = ok && vbrt::h2d((uint8_t*)c->m_defect.p + v0, h->defect + v0, nv, si);
        // consecutive chunks alternate between two compute streams (and N4 workspace sets), so the
        // stragglers of one chunk's persistent N4 kernel overlap the next chunk's CTAs
        const int ci = k % c->n_cstreams;                    // compute stream (and N4 workspace set) of this chunk
        vbrt::stream_t cs = ci == 0 ? s : ci == 1 ? c->stream2 : c->cextra[ci - 2], pre = c->pre[k & 1];
        ok = ok && vbrt::event_record(c->ev_copy[2 * (size_t)k], si) && vbrt::stream_wait(pre, c->ev_copy[2 * (size_t)k]);
        if (!ok) break;
        vbrt::stream_t ps = cs;
        rc = chunk_stages(c, k, r0, r1 - r0, params, thresh, flags, &d, pre, cs, ci, &ps, true);
        if (rc) { drain(c); return rc; }
        vbrt::stream_t so = c->stream_out;
        ok = ok && vbrt::event_record(c->ev_copy[2 * (size_t)k + 1], ps) && vbrt::stream_wait(so, c->ev_copy[2 * (size_t)k + 1]);
        if (!so_only) {
        if (flags & VB_DO_N4) ok = ok && vbrt::d2h(h->n4 + v0, d.n4 + v0, 4 * nv, so);
        if (d.logbias) ok = ok && vbrt::d2h(h->logbias + v0, d.logbias + v0, 4 * nv, so);
        if (flags & VB_DO_VDP) {
            ok = ok && vbrt::d2h(h->defect + v0, d.defect + v0, nv, so) && vbrt::d2h(h->defect_lb + v0, d.defect_lb + v0, nv, so);
            if (d.defect_border) ok = ok && vbrt::d2h(h->defect_border + v0, d.defect_border + v0, nv, so);
        }
        if (d.mask_border) ok = ok && vbrt::d2h(h->mask_border + v0, d.mask_border + v0, nv, so);
        if (flags & VB_DO_CI) ok = ok && vbrt::d2h(h->ci_shell + v0, d.ci_shell + v0, 2 * nv, so);
        }
        ok = ok && vbrt::d2h((vb_subject_stats*)c->h_stats.p + r0, d.stats + r0, sizeof(vb_subject_stats) * (size_t)(r1 - r0), so);
    }
    if (!drain(c) || !ok) return VB_ERR_CUDA;
    memcpy(h->stats, c->h_stats.p, stb);
    if (flags & VB_DO_CI)
        for (int i = 0; i < n; ++i)
            if (h->stats[i].ci_error) return fail(VB_ERR_CI_RMAX, "CI: maximum sphere radius reached (table exhausted)");
    return VB_OK;
}

// The whole path with compact host I/O (include/ventb200.h): per chunk, H2D of the uint16 image and
// the mask bits -> io_unpack_kernel -> the stages -> io_codes_kernel / io_ci_sparse_kernel -> D2H of
// N4HPvent, the code bytes and the per-subject records; the sparse CI lists follow once the host
// knows their sizes (the records of a chunk arrive while later chunks compute).
int vb_pipeline_batch_host_packed(vb_ctx* c, int32_t n, const int32_t* shapes, const int32_t* handles,
                                  const vb_n4_params* params, float thresh, uint32_t flags, const vb_host_io* io) {
    if (!c || !io || !io->stats || !shapes || n <= 0) return fail(VB_ERR_ARG, "pipeline_host_packed: null argument");
    if (!io->image_u16 == !io->image_f32) return fail(VB_ERR_ARG, "pipeline_host_packed: exactly one of image_u16 / image_f32");
    if (!io->mask_bits == !io->mask_u8) return fail(VB_ERR_ARG, "pipeline_host_packed: exactly one of mask_bits / mask_u8");
    if ((flags & VB_DO_CI) && !(flags & VB_DO_VDP)) return fail(VB_ERR_ARG, "pipeline_host_packed: CI needs VDP");
    if (io->codes && !(flags & VB_DO_VDP)) return fail(VB_ERR_ARG, "pipeline_host_packed: codes need VB_DO_VDP");
    const bool want_sparse = (flags & VB_DO_CI) && io->ci_index && io->ci_value;
    if ((flags & VB_DO_CI) && (!io->ci_index != !io->ci_value)) return fail(VB_ERR_ARG, "pipeline_host_packed: ci_index and ci_value go together");
    if (want_sparse && !io->ci_offset) return fail(VB_ERR_ARG, "pipeline_host_packed: ci_offset required");
    vbrt::set_device(c->device);
    vbrt::stream_t s = c->stream;
    size_t V = 0;
    for (int i = 0; i < n; ++i) V += (size_t)shapes[3 * i] * shapes[3 * i + 1] * shapes[3 * i + 2];
    const bool need_n4 = (flags & (VB_DO_N4 | VB_DO_VDP)) != 0;
    if ((flags & VB_DO_VDP) && !(flags & VB_DO_N4)) return fail(VB_ERR_ARG, "pipeline_host_packed: VDP needs N4 here (use vb_pipeline_batch_host to pass an N4 volume in)");
    const bool mb = (flags & VB_DO_MASK_BORDER) && io->codes;
    vb_buffers d;
    memset(&d, 0, sizeof(d));
    if (!ensure(c->m_mask, V) || !ensure(c->m_image, 4 * V) || !ensure(c->m_stats, sizeof(vb_subject_stats) * (size_t)n)) return VB_ERR_CUDA;
    d.mask = (const uint8_t*)c->m_mask.p; d.image = (const float*)c->m_image.p; d.stats = (vb_subject_stats*)c->m_stats.p;
    if (need_n4) { if (!ensure(c->m_n4, 4 * V)) return VB_ERR_CUDA; d.n4 = (float*)c->m_n4.p; }
    if (io->logbias && (flags & VB_DO_N4)) { if (!ensure(c->m_logbias, 4 * V)) return VB_ERR_CUDA; d.logbias = (float*)c->m_logbias.p; }
    if (flags & VB_DO_VDP) {
        if (!ensure(c->m_defect, V) || !ensure(c->m_lb, V)) return VB_ERR_CUDA;
        d.defect = (uint8_t*)c->m_defect.p; d.defect_lb = (uint8_t*)c->m_lb.p;
        if (io->codes) { if (!ensure(c->m_dborder, V) || !ensure(c->m_codes, V)) return VB_ERR_CUDA; d.defect_border = (uint8_t*)c->m_dborder.p; }
    }
    if (mb) { if (!ensure(c->m_mborder, V)) return VB_ERR_CUDA; d.mask_border = (uint8_t*)c->m_mborder.p; }
    if (flags & VB_DO_CI) { if (!ensure(c->m_shell, 2 * V)) return VB_ERR_CUDA; d.ci_shell = (uint16_t*)c->m_shell.p; }
    if (want_sparse && (!ensure(c->m_ci_idx, 4 * V) || !ensure(c->m_ci_val, 2 * V))) return VB_ERR_CUDA;
    if (io->image_u16 && !ensure(c->m_u16, 2 * V)) return VB_ERR_CUDA;
    int rc = pipeline_check(flags & ~VB_DO_MASK_BORDER, handles, params, &d);
    if (rc) return rc;
    rc = prepare(c, n, shapes, (flags & VB_DO_CI) ? handles : nullptr, s);
    if (rc) return rc;
    // mask words of every subject (host layout: ceil(V/32) words per subject, concatenated)
    std::vector<long long> bitoff((size_t)n + 1);
    {
        long long w = 0;
        for (int i = 0; i < n; ++i) { bitoff[(size_t)i] = w; w += (c->last_subj[(size_t)i].V + 31) / 32; }
        bitoff[(size_t)n] = w;
    }
    if (io->mask_bits) {
        if (!ensure(c->m_bits, 4 * (size_t)bitoff[(size_t)n])) return VB_ERR_CUDA;
        if (bitoff != c->last_bitoff) {
            if (!ensure(c->d_bitoff, 8 * bitoff.size())) return VB_ERR_CUDA;
            if (!drain(c)) return VB_ERR_CUDA;
            c->last_bitoff = bitoff;                        // (the upload reads the vector kept in the context)
            if (!vbrt::h2d(c->d_bitoff.p, c->last_bitoff.data(), 8 * bitoff.size(), s) || !vbrt::stream_sync(s)) return VB_ERR_CUDA;
        }
    }
    const std::vector<int> bounds = chunk_bounds(c, n);
    const int nchunks = (int)bounds.size() - 1;
    const PartsScope parts_scope(c, nchunks);
    while (c->ev_copy.size() < 3 * (size_t)nchunks) {
        vbrt::event_t e;
        if (!vbrt::event_create(&e)) return VB_ERR_CUDA;
        c->ev_copy.push_back(e);
    }
    const size_t stb = sizeof(vb_subject_stats) * (size_t)n;
    if (c->h_stats.cap < stb) {
        vbrt::host_free(c->h_stats.p);
        c->h_stats.p = nullptr; c->h_stats.cap = 0;
        if (!vbrt::host_malloc(&c->h_stats.p, stb * 2)) return VB_ERR_CUDA;
        c->h_stats.cap = stb * 2;
    }
    bool ok = true;
    for (int k = 0; k < nchunks && ok; ++k) {
        const int r0 = bounds[(size_t)k], r1 = bounds[(size_t)k + 1], cnt = r1 - r0;
        const size_t v0 = (size_t)c->last_subj[(size_t)r0].off;
        const size_t v1 = (size_t)c->last_subj[(size_t)(r1 - 1)].off + (size_t)c->last_subj[(size_t)(r1 - 1)].V;
        const size_t nv = v1 - v0;
        vbrt::stream_t si = c->stream_in;
        if (io->mask_bits) ok = ok && vbrt::h2d((uint32_t*)c->m_bits.p + bitoff[(size_t)r0], io->mask_bits + bitoff[(size_t)r0], 4 * (size_t)(bitoff[(size_t)r1] - bitoff[(size_t)r0]), si);
        else ok = ok && vbrt::h2d((uint8_t*)c->m_mask.p + v0, io->mask_u8 + v0, nv, si);
        if (io->image_u16) ok = ok && vbrt::h2d((uint16_t*)c->m_u16.p + v0, io->image_u16 + v0, 2 * nv, si);
        else ok = ok && vbrt::h2d((float*)c->m_image.p + v0, io->image_f32 + v0, 4 * nv, si);
        const int ci = k % c->n_cstreams;                    // compute stream (and N4 workspace set) of this chunk
        vbrt::stream_t cs = ci == 0 ? s : ci == 1 ? c->stream2 : c->cextra[ci - 2], pre = c->pre[k & 1];
        ok = ok && vbrt::event_record(c->ev_copy[3 * (size_t)k], si) && vbrt::stream_wait(pre, c->ev_copy[3 * (size_t)k]);
        if (!ok) break;
        vbrt::clear_stale_error();
        long long vmax = 0;
        for (int i = r0; i < r1; ++i) vmax = std::max<long long>(vmax, c->last_subj[(size_t)i].V);
        if (io->image_u16 || io->mask_bits) {
            UnpackArgs ua;
            ua.image_u16 = io->image_u16 ? (const uint16_t*)c->m_u16.p : nullptr;
            ua.mask_bits = io->mask_bits ? (const uint32_t*)c->m_bits.p : nullptr;
            ua.bit_off = (const long long*)c->d_bitoff.p + r0;
            ua.subj = (const VbSubject*)c->d_subj.p + r0;
            ua.image = (float*)c->m_image.p; ua.mask = (uint8_t*)c->m_mask.p;
            VB_LAUNCH(io_unpack_kernel, dim3(grid_x_for(c, cnt, (vmax + 31) / 32, 256), cnt), dim3(256), 0, pre, ua);
            ++c->launches;
            if (!vbrt::last_launch_ok("io_unpack_kernel")) { ok = false; break; }
        }
        rc = chunk_stages(c, k, r0, cnt, params, thresh, flags & ~(mb ? 0u : (uint32_t)VB_DO_MASK_BORDER), &d, pre, cs, ci, &cs, true);
        if (rc) { drain(c); return rc; }                    // (cs is now the stream the chunk's later stages ran on)
        if (io->codes) {
            CodesArgs ca;
            ca.lb = d.defect_lb; ca.defect = d.defect; ca.dborder = d.defect_border; ca.mborder = d.mask_border;
            ca.codes = (uint8_t*)c->m_codes.p; ca.v0 = (long long)v0; ca.v1 = (long long)v1;
            const int blocks = (int)std::max<size_t>(1, std::min<size_t>((size_t)c->sms * 8, (nv / 16 + 255) / 256));
            VB_LAUNCH(io_codes_kernel, dim3(blocks), dim3(256), 0, cs, ca);
            ++c->launches;
        }
        if (want_sparse) {
            SparseArgs sa;
            sa.subj = (const VbSubject*)c->d_subj.p + r0; sa.count = (const int*)c->ws_count.p + r0;
            sa.list = (const int*)c->ws_list.p; sa.shell = d.ci_shell;
            sa.index = (uint32_t*)c->m_ci_idx.p + v0; sa.value = (uint16_t*)c->m_ci_val.p + v0; sa.n = cnt;
            VB_LAUNCH(io_ci_sparse_kernel, dim3(cnt), dim3(256), 0, cs, sa);
            ++c->launches;
        }
        if (!vbrt::last_launch_ok("io kernels")) { ok = false; break; }
        vbrt::stream_t so = c->stream_out;
        ok = ok && vbrt::event_record(c->ev_copy[3 * (size_t)k + 1], cs) && vbrt::stream_wait(so, c->ev_copy[3 * (size_t)k + 1]);
        // records first: the host needs them to place the chunk's sparse CI list
        ok = ok && vbrt::d2h((vb_subject_stats*)c->h_stats.p + r0, d.stats + r0, sizeof(vb_subject_stats) * (size_t)cnt, so);
        ok = ok && vbrt::event_record(c->ev_copy[3 * (size_t)k + 2], so);
        if (io->n4 && (flags & VB_DO_N4)) ok = ok && vbrt::d2h(io->n4 + v0, d.n4 + v0, 4 * nv, so);
        if (io->logbias && d.logbias) ok = ok && vbrt::d2h(io->logbias + v0, d.logbias + v0, 4 * nv, so);
        if (io->codes) ok = ok && vbrt::d2h(io->codes + v0, (uint8_t*)c->m_codes.p + v0, nv, so);
    }
    // sparse CI lists: as the records of each chunk arrive (later chunks are still computing)
    int64_t total = 0;
    bool too_small = false;
    if (ok && want_sparse) {
        io->ci_offset[0] = 0;
        for (int k = 0; k < nchunks && ok; ++k) {
            const int r0 = bounds[(size_t)k], r1 = bounds[(size_t)k + 1];
            ok = ok && vbrt::event_sync(c->ev_copy[3 * (size_t)k + 2]);
            if (!ok) break;
            const vb_subject_stats* hs = (const vb_subject_stats*)c->h_stats.p;
            const int64_t base = total;
            for (int i = r0; i < r1; ++i) { total += hs[i].ci_n; io->ci_offset[i + 1] = total; }
            if (total > io->ci_capacity) { too_small = true; continue; }
            const size_t v0 = (size_t)c->last_subj[(size_t)r0].off, cntk = (size_t)(total - base);
            if (cntk) {
                ok = ok && vbrt::d2h(io->ci_index + base, (uint32_t*)c->m_ci_idx.p + v0, 4 * cntk, c->stream_out);
                ok = ok && vbrt::d2h(io->ci_value + base, (uint16_t*)c->m_ci_val.p + v0, 2 * cntk, c->stream_out);
            }
        }
    }
    if (!drain(c) || !ok) return VB_ERR_CUDA;
    memcpy(io->stats, c->h_stats.p, stb);
    if (too_small) return fail(VB_ERR_ARG, "pipeline_host_packed: ci_capacity too small for the batch's defect voxels");
    if (flags & VB_DO_CI)
        for (int i = 0; i < n; ++i)
            if (io->stats[i].ci_error) return fail(VB_ERR_CI_RMAX, "CI: maximum sphere radius reached (table exhausted)");
    return VB_OK;
}

// ------------------------------------------------------------------------------- overlays (8f3)
static int ovl_begin(vb_ctx* c, int n, vbrt::stream_t s) {
    if (!ensure(c->d_ovl, sizeof(OvlStats) * (size_t)n)) return VB_ERR_CUDA;
    vbrt::clear_stale_error();
    VB_LAUNCH(ovl_init_kernel, dim3((n + 127) / 128), dim3(128), 0, s, (OvlStats*)c->d_ovl.p, subj_ptr(c), n);
    ++c->launches;
    return VB_OK;
}

int vb_overlay_rgb_batch_dev(vb_ctx* c, int32_t n, const int32_t* shapes, const float* a, const uint8_t* b,
                             int32_t gui, int32_t mode, uint8_t* rgb, void* stream) {
    if (!c || !a || !b || !rgb) return fail(VB_ERR_ARG, "overlay_rgb: null argument");
    if (mode != VB_NORM_F32 && mode != VB_NORM_F64) return fail(VB_ERR_ARG, "overlay_rgb: bad mode");
    vbrt::set_device(c->device);
    vbrt::stream_t s = (vbrt::stream_t)stream;
    int rc = prepare(c, n, shapes, nullptr, s);
    if (rc) return rc;
    c->r0 = 0;
    if ((rc = ovl_begin(c, n, s))) return rc;
    const int gx = grid_x_for(c, n, vmax_of(c), 256);
    OvlMinMaxArgs m;
    m.a[0] = a; m.a[1] = nullptr; m.a[2] = nullptr; m.subj = subj_ptr(c); m.st = (OvlStats*)c->d_ovl.p; m.use_abs = 1;
    VB_LAUNCH(ovl_minmax_kernel, dim3(gx, n), dim3(256), 0, s, m);
    OvlRgbArgs r;
    r.a = a; r.b = b; r.subj = subj_ptr(c); r.st = (const OvlStats*)c->d_ovl.p; r.rgb = rgb; r.gui = gui ? 1 : 0; r.mode = mode;
    VB_LAUNCH(ovl_rgb_kernel, dim3(gx, n), dim3(256), 0, s, r);
    c->launches += 2;
    return vbrt::last_launch_ok("overlay kernels") ? VB_OK : VB_ERR_CUDA;
}

int vb_overlay_rgb_batch_host(vb_ctx* c, int32_t n, const int32_t* shapes, const float* a, const uint8_t* b,
                              int32_t gui, int32_t mode, uint8_t* rgb) {
    if (!c || !shapes || n <= 0 || !a || !b || !rgb) return fail(VB_ERR_ARG, "overlay_rgb: null argument");
    vbrt::set_device(c->device);
    vbrt::stream_t s = c->stream;
    size_t V = 0;
    for (int i = 0; i < n; ++i) V += (size_t)shapes[3 * i] * shapes[3 * i + 1] * shapes[3 * i + 2];
    if (!ensure(c->m_ovl_in[0], 4 * V) || !ensure(c->m_ovl_in[3], V) || !ensure(c->m_rgb, 3 * V)) return VB_ERR_CUDA;
    if (!vbrt::h2d(c->m_ovl_in[0].p, a, 4 * V, s) || !vbrt::h2d(c->m_ovl_in[3].p, b, V, s)) return VB_ERR_CUDA;
    int rc = vb_overlay_rgb_batch_dev(c, n, shapes, (const float*)c->m_ovl_in[0].p, (const uint8_t*)c->m_ovl_in[3].p, gui, mode,
                                      (uint8_t*)c->m_rgb.p, s);
    if (rc) { vbrt::stream_sync(s); return rc; }
    if (!vbrt::d2h(rgb, c->m_rgb.p, 3 * V, s) || !vbrt::stream_sync(s)) return VB_ERR_CUDA;
    return VB_OK;
}

// boxes of the prepared batch from the device mask -> host (synchronises `s`)
static int crop_boxes(vb_ctx* c, int n, const uint8_t* d_mask, int border, int border_slices, int32_t* boxes, vbrt::stream_t s) {
    int rc = ovl_begin(c, n, s);
    if (rc) return rc;
    VB_LAUNCH(ovl_cropbox_kernel, dim3(n), dim3(512), 0, s, d_mask, subj_ptr(c), border, border_slices, (OvlStats*)c->d_ovl.p);
    ++c->launches;
    if (!vbrt::last_launch_ok("ovl_cropbox_kernel")) return VB_ERR_CUDA;
    std::vector<OvlStats> st((size_t)n);
    if (!vbrt::d2h(st.data(), c->d_ovl.p, sizeof(OvlStats) * (size_t)n, s) || !vbrt::stream_sync(s)) return VB_ERR_CUDA;
    bool empty = false;
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < 6; ++k) {
            boxes[6 * i + k] = st[(size_t)i].error ? -1 : st[(size_t)i].box[k];
            empty = empty || st[(size_t)i].error;
        }
    return empty ? fail(VB_ERR_EMPTY, "cropToData: an axis has no index > 0 with signal (list index out of range)") : VB_OK;
}

int vb_crop_boxes_batch_host(vb_ctx* c, int32_t n, const int32_t* shapes, const uint8_t* mask, int32_t border,
                             int32_t border_slices, int32_t* boxes) {
    if (!c || !shapes || n <= 0 || !mask || !boxes || border < 0) return fail(VB_ERR_ARG, "crop_boxes: bad argument");
    vbrt::set_device(c->device);
    vbrt::stream_t s = c->stream;
    size_t V = 0;
    for (int i = 0; i < n; ++i) V += (size_t)shapes[3 * i] * shapes[3 * i + 1] * shapes[3 * i + 2];
    if (!ensure(c->m_ovl_in[3], V) || !vbrt::h2d(c->m_ovl_in[3].p, mask, V, s)) return VB_ERR_CUDA;
    int rc = prepare(c, n, shapes, nullptr, s);
    if (rc) { vbrt::stream_sync(s); return rc; }
    c->r0 = 0;
    return crop_boxes(c, n, (const uint8_t*)c->m_ovl_in[3].p, border, border_slices, boxes, s);
}

int vb_montage_batch_host(vb_ctx* c, int32_t n, const int32_t* shapes, const int32_t* handles, const double* vox_min,
                          const vb_montage_inputs* h, const double* parula, int32_t* boxes, int64_t* out_offset,
                          uint8_t* out, int64_t out_capacity) {
    if (!c || !shapes || n <= 0 || !h || !h->image || !h->n4 || !h->mask || !h->mask_border || !h->defect || !parula || !boxes ||
        !out_offset || !out)
        return fail(VB_ERR_ARG, "montage: null argument");
    if (h->ci_shell && (!handles || !vox_min)) return fail(VB_ERR_ARG, "montage: ci_shell needs table handles and vox_min");
    vbrt::set_device(c->device);
    vbrt::stream_t s = c->stream;
    size_t V = 0;
    for (int i = 0; i < n; ++i) V += (size_t)shapes[3 * i] * shapes[3 * i + 1] * shapes[3 * i + 2];
    // device mirrors: 0 n4, 1 image, 2 proton, 3 mask, 4 mask border, 5 defect (+ m_shell)
    bool ok = ensure(c->m_ovl_in[0], 4 * V) && ensure(c->m_ovl_in[1], 4 * V) && ensure(c->m_ovl_in[3], V) && ensure(c->m_ovl_in[4], V) &&
              ensure(c->m_ovl_in[5], V) && ensure(c->d_parula, 64 * 3 * sizeof(double)) && ensure(c->d_outoff, 8 * (size_t)n) &&
              ensure(c->d_voxmin, sizeof(double) * (size_t)n);
    if (h->proton) ok = ok && ensure(c->m_ovl_in[2], 4 * V);
    if (h->ci_shell) ok = ok && ensure(c->m_shell, 2 * V);
    if (!ok) return VB_ERR_CUDA;
    ok = vbrt::h2d(c->m_ovl_in[0].p, h->n4, 4 * V, s) && vbrt::h2d(c->m_ovl_in[1].p, h->image, 4 * V, s) &&
         vbrt::h2d(c->m_ovl_in[3].p, h->mask, V, s) && vbrt::h2d(c->m_ovl_in[4].p, h->mask_border, V, s) &&
         vbrt::h2d(c->m_ovl_in[5].p, h->defect, V, s) && vbrt::h2d(c->d_parula.p, parula, 64 * 3 * sizeof(double), s);
    if (h->proton) ok = ok && vbrt::h2d(c->m_ovl_in[2].p, h->proton, 4 * V, s);
    if (h->ci_shell) ok = ok && vbrt::h2d(c->m_shell.p, h->ci_shell, 2 * V, s) && vbrt::h2d(c->d_voxmin.p, vox_min, sizeof(double) * (size_t)n, s);
    if (!ok) { vbrt::stream_sync(s); return VB_ERR_CUDA; }
    int rc = prepare(c, n, shapes, h->ci_shell ? handles : nullptr, s);
    if (!rc && h->ci_shell) rc = upload_tables(c, s);
    if (rc) { vbrt::stream_sync(s); return rc; }
    c->r0 = 0;
    if ((rc = crop_boxes(c, n, (const uint8_t*)c->m_ovl_in[3].p, 5, 0, boxes, s))) return rc;      // cropToData(mask, border=5)
    std::vector<long long> off((size_t)n + 1);
    long long total = 0, pxmax = 0;
    for (int i = 0; i < n; ++i) {
        off[(size_t)i] = total;
        const long long px = 7LL * (boxes[6 * i + 1] - boxes[6 * i]) * (boxes[6 * i + 3] - boxes[6 * i + 2]) * (boxes[6 * i + 5] - boxes[6 * i + 4]);
        total += 3 * px;
        pxmax = std::max(pxmax, px);
    }
    off[(size_t)n] = total;
    for (int i = 0; i <= n; ++i) out_offset[i] = off[(size_t)i];
    if (total > out_capacity) return fail(VB_ERR_ARG, "montage: out_capacity too small");
    if (!ensure(c->m_rgb, (size_t)total)) return VB_ERR_CUDA;
    if (!vbrt::h2d(c->d_outoff.p, off.data(), 8 * (size_t)n, s) || !vbrt::stream_sync(s)) return VB_ERR_CUDA;   // `off` is a stack object
    vbrt::clear_stale_error();
    OvlMinMaxArgs m;
    m.a[0] = (const float*)c->m_ovl_in[0].p; m.a[1] = (const float*)c->m_ovl_in[1].p;
    m.a[2] = h->proton ? (const float*)c->m_ovl_in[2].p : nullptr;
    m.subj = subj_ptr(c); m.st = (OvlStats*)c->d_ovl.p; m.use_abs = 0;
    const int gx = grid_x_for(c, n, std::max<long long>(pxmax, 1), 256);
    VB_LAUNCH(ovl_minmax_kernel, dim3(gx, n), dim3(256), 0, s, m);
    OvlMontageArgs a;
    a.proton = h->proton ? (const float*)c->m_ovl_in[2].p : nullptr; a.image = (const float*)c->m_ovl_in[1].p;
    a.n4 = (const float*)c->m_ovl_in[0].p; a.mborder = (const uint8_t*)c->m_ovl_in[4].p; a.defect = (const uint8_t*)c->m_ovl_in[5].p;
    a.shell = h->ci_shell ? (const uint16_t*)c->m_shell.p : nullptr;
    a.subj = subj_ptr(c); a.st = (OvlStats*)c->d_ovl.p; a.tables = (const CiTableDev*)c->d_tables.p;
    a.vox_min = (const double*)c->d_voxmin.p; a.parula = (const double*)c->d_parula.p;
    a.out_off = (const long long*)c->d_outoff.p; a.out = (uint8_t*)c->m_rgb.p;
    a.image_mode = h->image_mode; a.proton_mode = h->proton_mode;
    VB_LAUNCH(ovl_montage_kernel, dim3(gx, n), dim3(256), 0, s, a);
    c->launches += 2;
    if (!vbrt::last_launch_ok("montage kernels")) return VB_ERR_CUDA;
    std::vector<OvlStats> st((size_t)n);
    if (!vbrt::d2h(out, c->m_rgb.p, (size_t)total, s) || !vbrt::d2h(st.data(), c->d_ovl.p, sizeof(OvlStats) * (size_t)n, s) ||
        !vbrt::stream_sync(s))
        return VB_ERR_CUDA;
    for (int i = 0; i < n; ++i)
        if (st[(size_t)i].error == 2)
            return fail(VB_ERR_CI_COLOR, "screenShot: a CI value >= 40 mm indexes past the 64-entry colour table (index out of bounds)");
    return VB_OK;
}

// ------------------------------------------------------------------- k-space reconstruction (8f4)
int vb_recon_batch_dev(vb_ctx* c, int32_t n, const int32_t* shapes, const float* kspace, float* image, void* stream) {
    if (!c || !kspace || !image) return fail(VB_ERR_ARG, "recon: null argument");
    vbrt::set_device(c->device);
    vbrt::stream_t s = (vbrt::stream_t)stream;
    int rc = prepare(c, n, shapes, nullptr, s);
    if (rc) return rc;
    c->r0 = 0;
    int nmax = 1, n0max = 1, n1max = 1, n2max = 1;
    for (const VbSubject& S : c->last_subj) {
        n0max = std::max(n0max, S.n0); n1max = std::max(n1max, S.n1); n2max = std::max(n2max, S.n2);
    }
    nmax = std::max(n0max, n1max);
    (void)n2max;
    bool need_y = false;
    for (const VbSubject& S : c->last_subj) need_y = need_y || (S.n0 & (S.n0 - 1)) || (S.n1 & (S.n1 - 1));
    // transforms per CTA: as many as keep three CTAs per SM (at least 8)
    const long long budget = std::min<long long>(c->max_smem - 1024, 72 * 1024);
    int nl = (int)std::min<long long>(VB_RECON_NL_MAX, (budget / 8 - nmax) / ((need_y ? 2 : 1) * (long long)nmax));
    if (nl < 8) nl = (int)std::min<long long>(8, (((long long)c->max_smem - 1024) / 8 - nmax) / ((need_y ? 2 : 1) * (long long)nmax));
    if (nl < 1) return fail(VB_ERR_UNSUPPORTED, "recon: matrix size too large for the shared-memory transform");
    const size_t smem = sizeof(float) * 2 * ((size_t)(need_y ? 2 : 1) * nmax * nl + nmax);
    if (!ensure(c->d_ktmp, 8 * (size_t)c->v_total)) return VB_ERR_CUDA;
    if (!vbrt::allow_smem(recon_pass_kernel, (int)smem)) return VB_ERR_CUDA;
    vbrt::clear_stale_error();
    long long inner_max = 1, t2max = 1;
    for (const VbSubject& S : c->last_subj) {
        inner_max = std::max<long long>(inner_max, (long long)S.n1 * S.n2);
        const int KS = std::min(S.n2, nl), LU = std::max(1, nl / KS);
        t2max = std::max<long long>(t2max, (long long)((S.n0 + LU - 1) / LU) * ((S.n2 + KS - 1) / KS));
    }
    ReconArgs a;
    a.subj = subj_ptr(c); a.nl = nl; a.need_y = need_y ? 1 : 0;
    a.src = (const float2*)kspace; a.dst = (float2*)c->d_ktmp.p; a.pass = 0;
    VB_LAUNCH(recon_pass_kernel, dim3((unsigned)((inner_max + nl - 1) / nl), 1, n), dim3(VB_RECON_THREADS), smem, s, a, nmax);
    a.src = (const float2*)c->d_ktmp.p; a.dst = (float2*)image; a.pass = 1;
    VB_LAUNCH(recon_pass_kernel, dim3((unsigned)t2max, 1, n), dim3(VB_RECON_THREADS), smem, s, a, nmax);
    c->launches += 2;
    return vbrt::last_launch_ok("recon_pass_kernel") ? VB_OK : VB_ERR_CUDA;
}

int vb_recon_batch_host(vb_ctx* c, int32_t n, const int32_t* shapes, const float* kspace, float* image) {
    if (!c || !shapes || n <= 0 || !kspace || !image) return fail(VB_ERR_ARG, "recon: null argument");
    vbrt::set_device(c->device);
    vbrt::stream_t s = c->stream;
    size_t V = 0;
    for (int i = 0; i < n; ++i) V += (size_t)shapes[3 * i] * shapes[3 * i + 1] * shapes[3 * i + 2];
    if (!ensure(c->m_k[0], 8 * V) || !ensure(c->m_k[1], 8 * V)) return VB_ERR_CUDA;
    if (!vbrt::h2d(c->m_k[0].p, kspace, 8 * V, s)) return VB_ERR_CUDA;
    int rc = vb_recon_batch_dev(c, n, shapes, (const float*)c->m_k[0].p, (float*)c->m_k[1].p, s);
    if (rc) { vbrt::stream_sync(s); return rc; }
    if (!vbrt::d2h(image, c->m_k[1].p, 8 * V, s) || !vbrt::stream_sync(s)) return VB_ERR_CUDA;
    return VB_OK;
}

int vb_host_alloc(vb_ctx* c, size_t bytes, void** ptr) {
    if (!c || !ptr) return fail(VB_ERR_ARG, "host_alloc: null argument");
    vbrt::set_device(c->device);
    *ptr = nullptr;
    if (!vbrt::host_malloc(ptr, bytes)) return VB_ERR_CUDA;
    return VB_OK;
}

int vb_host_free(vb_ctx* c, void* ptr) {
    if (!c) return fail(VB_ERR_ARG, "host_free: null ctx");
    vbrt::set_device(c->device);
    vbrt::host_free(ptr);
    return VB_OK;
}

int vb_set_profiling(vb_ctx* c, int32_t on) {
    if (!c) return fail(VB_ERR_ARG, "null ctx");
    c->prof = on != 0;
    c->ev_used = 0;
    return VB_OK;
}

int vb_profile_fetch(vb_ctx* c, double* stage_ms, int32_t* n_calls) {
    if (!c || !stage_ms) return fail(VB_ERR_ARG, "profile_fetch: null argument");
    vbrt::set_device(c->device);
    for (int k = 0; k < VB_N_STAGES; ++k) stage_ms[k] = 0.0;
    const size_t calls = c->ev_used / (VB_N_STAGES + 1);
    for (size_t i = 0; i < calls; ++i) {
        vbrt::event_t* ev = c->ev.data() + i * (VB_N_STAGES + 1);
        if (!vbrt::event_sync(ev[VB_N_STAGES])) return VB_ERR_CUDA;
        for (int k = 0; k < VB_N_STAGES; ++k) {
            float ms = 0.0f;
            if (!vbrt::event_elapsed(&ms, ev[k], ev[k + 1])) return VB_ERR_CUDA;
            stage_ms[k] += (double)ms;
        }
    }
    if (n_calls) *n_calls = (int32_t)calls;
    c->ev_used = 0;
    return VB_OK;
}

int vb_set_n4_cluster(vb_ctx* c, int32_t max_k) {
    if (!c) return fail(VB_ERR_ARG, "null ctx");
    if (max_k < 0 || max_k > 8) return fail(VB_ERR_ARG, "N4 cluster size: 0 (automatic) or 1..8");
    c->n4_cluster_max = max_k;
    return VB_OK;
}

int32_t vb_n4_cluster_size(const vb_ctx* c) { return c ? c->n4_cluster : 1; }

int vb_set_debug(vb_ctx* c, int32_t on) {
    if (!c) return fail(VB_ERR_ARG, "null ctx");
    c->debug = on != 0;
    return VB_OK;
}

int vb_n4_debug_fetch(vb_ctx* c, int32_t subject, float* hist, float* emap, float* lattice, int32_t lattice_capacity,
                      int32_t* ncp, float* conv_trace, int32_t trace_capacity) {
    if (!c || !c->dbg.p || subject < 0 || subject >= c->dbg_n) return fail(VB_ERR_ARG, "no debug capture");
    vbrt::set_device(c->device);
    std::vector<float> rec((size_t)c->dbg_stride);
    if (!vbrt::device_sync()) return VB_ERR_CUDA;
    if (!vbrt::d2h(rec.data(), (const float*)c->dbg.p + (size_t)subject * c->dbg_stride, sizeof(float) * rec.size(), c->stream) ||
        !vbrt::stream_sync(c->stream))
        return VB_ERR_CUDA;
    const int nb = c->dbg_nbins;
    if (hist) memcpy(hist, rec.data() + VB_DBG_HDR, sizeof(float) * (size_t)nb);
    if (emap) memcpy(emap, rec.data() + VB_DBG_HDR + nb, sizeof(float) * (size_t)nb);
    const int n0 = (int)rec[2], n1 = (int)rec[3], n2 = (int)rec[4];
    if (ncp) { ncp[0] = n0; ncp[1] = n1; ncp[2] = n2; }
    if (lattice) memcpy(lattice, rec.data() + VB_DBG_HDR + 2 * nb, sizeof(float) * (size_t)std::min(lattice_capacity, n0 * n1 * n2));
    const int nt = (int)rec[5];
    if (conv_trace && trace_capacity >= 12 + nt) {         // phase microseconds ride behind the trace
        for (int i = 0; i < trace_capacity; ++i) conv_trace[i] = -1.0f;
        memcpy(conv_trace, rec.data() + VB_DBG_HDR + 2 * nb + c->dbg_Lmax, sizeof(float) * (size_t)nt);
        memcpy(conv_trace + trace_capacity - 12, rec.data() + 8, sizeof(float) * 12);
    } else if (conv_trace) {
        for (int i = 0; i < trace_capacity; ++i) conv_trace[i] = -1.0f;
        memcpy(conv_trace, rec.data() + VB_DBG_HDR + 2 * nb + c->dbg_Lmax, sizeof(float) * (size_t)std::min(trace_capacity, nt));
    }
    return VB_OK;
}

}  // extern "C"
