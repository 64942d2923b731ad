// rve_abi.cu — C-ABI of librve_b200.so (declared in include/rve_b200.h).
// Host side: batch set-up, uploads, the PCG driver loop, epilogue downloads.  No torch types, no
// CPU fallback: every entry point fails with RVE_E_CUDA if the device is unusable.
#include <cuda_runtime.h>
#include <malloc.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "rve_kernels.cuh"
#include "rve_symbolic.hpp"
#include "rve_devsym.cuh"
#include "rve_matfree.cuh"
#include "rve_stream.cuh"

namespace {

thread_local std::string g_create_err;

template <class T>
struct DBuf {
    T* p = nullptr;
    size_t n = 0, cap = 0;
    // grow-only: repeated passes over same-sized batches never touch cudaMalloc/cudaFree again
    cudaError_t alloc(size_t count) {
        if (count <= cap && p) { n = count; return cudaSuccess; }
        release();
        if (count == 0) return cudaSuccess;
        cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
        if (e == cudaSuccess) { n = count; cap = count; }
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; cap = 0; }
    ~DBuf() { release(); }
};

// typed view into the pinned staging arena (same surface as the std::vector it replaces)
template <class T>
struct Span {
    T* p = nullptr;
    size_t n = 0;
    T& operator[](size_t i) const { return p[i]; }
    T* data() const { return p; }
    size_t size() const { return n; }
    bool empty() const { return n == 0; }
    T* begin() const { return p; }
};

// grow-only page-locked host buffer: the flattened batch is built directly in it so that every
// cudaMemcpyAsync is a true asynchronous DMA at PCIe speed
struct PinnedArena {
    char* base = nullptr;
    size_t cap = 0, off = 0;
    cudaError_t reserve(size_t bytes) {
        off = 0;
        if (bytes <= cap) return cudaSuccess;
        if (base) cudaFreeHost(base);
        base = nullptr; cap = 0;
        cudaError_t e = cudaHostAlloc((void**)&base, bytes, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    template <class T>
    Span<T> take(size_t count) {
        off = (off + 255) & ~(size_t)255;
        Span<T> sp; sp.p = (T*)(base + off); sp.n = count;
        off += count * sizeof(T);
        return sp;
    }
    static size_t need(size_t count, size_t elem) { return ((count * elem + 255) & ~(size_t)255) + 256; }
    void release() { if (base) cudaFreeHost(base); base = nullptr; cap = 0; }
};

}  // namespace

// sizes of one system (device symbolic phase, rve_devsym.cuh)
struct SysInfo { int nv = 0, nt = 0, ne = 0, nn = 0, na = 0, nb = 0, nbe = 0, nwin = 0, nsl = 0; long long nblk = 0, ngrp = 0, nadj = 0; };

struct rve_batch_s {
    int device = 0;
    std::string err;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_block = nullptr;   // cudaEventBlockingSync
    // host-side batch description
    int n_sys = 0, model = 0, vref_nb = 0, nvref = 0;
    bool have_batch = false, assembled = false, solved = false;
    bool have_samples = false;          // the Vref samples / epilogue buffers belong to the last solve (rve_project needs them)
    int K = 0, NL = 0;  // solver rhs count, load count of the last solve
    Mat8 mat;
    double vref_box[4];
    std::vector<SysInfo> info;          // per-system sizes found by the device symbolic phase
    std::vector<int> voff, toff, cellrank, h_cnt;
    std::vector<double> h_ext, h_box;
    std::vector<unsigned> u_blk, u_grp, u_adjf;
    std::vector<int> node_base, elem_base, row_base, bnd_base, bedge_base, win_base, slice_base, halo_base, cn_base;
    std::vector<long long> blk_base, grp_base, ent_base;
    long long N = 0, E = 0, R = 0, NB = 0, NBN = 0, NBE = 0, NG = 0;
    int n_win = 0, max_tile = 0, max_cn = 0;
    int ncx = 0, ncy = 0;
    Levels LV;
    // device arrays
    DBuf<double> node_xy, box, bval, wmean, bnd_s, bedge_nl, samp_lam;
    DBuf<float4> dinv;
    DBuf<int> node_row, node_bnd, node_sys, elem_node, elem_sys, d_row_base, bnd_sys, bnd_node,
        bedge, d_bedge_base, samp_elem, win_sys, win_row0, win_n, win_slice0, win_halo0, win_cn0, sys_win0, halo,
        cn_node, active, cnt, d_iters, n_active;
    DBuf<unsigned char> elem_reg;
    DBuf<unsigned> slice_ptr, cn_beg;
    DBuf<uint16_t> scol, cn_ent;
    DBuf<unsigned> d_row_ptr, adjf_ptr;
    DBuf<int> adjf, csr_col, win_nhalo, win_ncn, row_node, sym_cnt;
    DBuf<double> d_xy, sys_ext;
    DBuf<int> d_tri, d_region, d_voff, d_toff, scrA, scrB, sys_cnt, node_va, node_vb, d_node_base, d_bnd_base, d_cellrank, d_slice_base;
    DBuf<unsigned> d_blk_base, d_grp_base, d_adjf_base;
    DBuf<unsigned char> emap, rlen, win_active;
    DBuf<int> sys_blocks;
    std::vector<long long> true_blocks;   // per system, found by the device pattern build
    std::vector<long long> adjf_base;
    int max_words = 0;
    DBuf<RowInfo> rowinfo;
    DBuf<int> m_node, m_incl;          // rve_set_batch_morphed: the template's moving nodes
    DBuf<double> m_d, m_dir, m_centre, m_prm;
    DBuf<unsigned char> rep_scratch;   // rve_set_batch_shared: system-0 copies of the index arrays while they are replicated
    DBuf<WinDesc> win_desc;
    DBuf<double> win_scal;             // [n_win][8] alpha / beta of the window's system (written by the k_fin_* kernels)
    DBuf<double> lvA[RVE_MAXLEV], lvD[RVE_MAXLEV];
    DBuf<float> lvR[RVE_MAXLEV], lvX[RVE_MAXLEV], lvT[RVE_MAXLEV];   // V-cycle vectors (fp32)
    DBuf<stv> lvA4[RVE_MAXLEV], lvH4[RVE_MAXLEV];
    DBuf<double> p, q, x, r, b, part, sc, cpart;
    DBuf<double> eps, eps_eff, Bmat, uD, gval, d_vbox, U, acc, sigL, sigT, aout, Bout, sol, Mx, Gaff, aaff;
    DBuf<double> MWt, transl, Yd, gradT;
    int opt_sym_B = 0;
    int opt_resid_replace = 0;         // > 0: every so many iterations r is recomputed as b - A x (RVE_OPT_RESIDUAL_REPLACE)
    int n_sm = 148;
    int updp_blocks[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    int applyp_blocks[4] = {0, 0, 0, 0};
    int dirp_blocks[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};   // resident CTAs per SM of k_update_p<K> (by precond), 0 = not asked yet
    // matrix-free operator (rve_matfree.cuh): window-element tables; the SELL values (bval) are built on demand only
    int opt_operator = RVE_OPERATOR_MATFREE;
    bool have_bval = false;
    bool have_emap = false;                    // element -> SELL slot map built (assembled operator only)
    int max_slots = 0;
    long long n_we = 0, n_halo = 0;
    DBuf<int> win_we0, win_nwe, we_elem;
    DBuf<unsigned char> win_sw, row_nadj;
    DBuf<uint32_t> we_idx;
    DBuf<double2> we_geo;
    int* h_nactive = nullptr;  // pinned
    PinnedArena arena;
    std::vector<cudaEvent_t> evpool;  // per-iteration SpMV timing
    std::vector<cudaEvent_t> evdbg;   // RVE_KTIME diagnostics
    double launches = 0;              // kernels launched since rve_create
    double h2d_bytes = 0, d2h_bytes = 0;  // host<->device traffic since rve_create
    double t_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double c_cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

#define H_CHECK(h)                                        \
    if (!(h)) return RVE_E_ARG;                           \
    if (cudaSetDevice((h)->device) != cudaSuccess) { (h)->err = "cudaSetDevice failed"; return RVE_E_CUDA; }

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                           \
            return RVE_E_CUDA;                                                                     \
        }                                                                                          \
    } while (0)

#define FAIL(code, msg) do { h->err = (msg); return (code); } while (0)

// wait for the handle's stream
static cudaError_t stream_wait(rve_batch_s* h) {
    // RVE_BLOCKING_SYNC=1: sleep instead of spinning (a few % slower per wait, but it frees a core per rank when
    // several ranks share few host cores); default: spin
    static const bool blocking = [] { const char* e = std::getenv("RVE_BLOCKING_SYNC"); return e && e[0] == '1'; }();
    if (!blocking) return cudaStreamSynchronize(h->stream);
    cudaError_t e = cudaEventRecord(h->ev_block, h->stream);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(h->ev_block);
}

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// RVE_DEBUG_SYNC=1: synchronise after the named kernel and name it in the error (bisecting a faulting launch)
#define DBG_SYNC(name)                                                                                   \
    do {                                                                                                 \
        static const bool dbg_ = [] { const char* e = std::getenv("RVE_DEBUG_SYNC"); return e && e[0] == '1'; }(); \
        if (dbg_) {                                                                                      \
            cudaError_t e_ = cudaStreamSynchronize(h->stream);                                           \
            if (e_ == cudaSuccess) e_ = cudaGetLastError();                                              \
            if (e_ != cudaSuccess) { h->err = std::string(name) + ": " + cudaGetErrorString(e_); return RVE_E_CUDA; } \
        }                                                                                                \
    } while (0)

#define UPL(d, v) do { CU(upload((d), (v), st)); h->h2d_bytes += (double)(v).size() * sizeof((v)[0]); } while (0)

template <class T>
static cudaError_t upload(DBuf<T>& d, const Span<T>& v, cudaStream_t s) {
    cudaError_t e = d.alloc(v.size());
    if (e != cudaSuccess || v.empty()) return e;
    return cudaMemcpyAsync(d.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
}

template <class T>
static cudaError_t upload(DBuf<T>& d, const std::vector<T>& v, cudaStream_t s) {
    cudaError_t e = d.alloc(v.size());
    if (e != cudaSuccess || v.empty()) return e;
    return cudaMemcpyAsync(d.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
}

static WinTab win_tab(rve_batch_s* h) {
    WinTab wt;
    wt.win_sys = h->win_sys.p; wt.win_row0 = h->win_row0.p; wt.win_n = h->win_n.p; wt.win_slice0 = h->win_slice0.p;
    wt.win_halo0 = h->win_halo0.p; wt.win_nhalo = h->win_nhalo.p; wt.win_cn0 = h->win_cn0.p; wt.win_ncn = h->win_ncn.p; wt.win_active = h->win_active.p; wt.sys_win0 = h->sys_win0.p; wt.row_base = h->d_row_base.p;
    return wt;
}

static Sell sell_tab(rve_batch_s* h) {
    Sell sl;
    sl.slice_ptr = h->slice_ptr.p; sl.scol = h->scol.p; sl.halo = h->halo.p;
    return sl;
}

// dynamic shared memory of the PCG kernels
static size_t smem_spmv(rve_batch_s* h, int K) { return (size_t)h->max_tile * K * sizeof(double2); }
static size_t smem_update(int K) { return (size_t)RVE_WIN * K * sizeof(double2) + RVE_WIN * sizeof(float2) + 4 * RVE_WIN * sizeof(uint16_t); }
static size_t smem_direction(rve_batch_s* h, int K) { return (size_t)h->max_cn * K * sizeof(double2); }
static size_t smem_apply(rve_batch_s* h, int K) { return ((size_t)apply_tile_pitch(h->max_tile) + h->max_slots) * K * sizeof(double2); }

// Opt-in dynamic shared memory of the kernels that size their tiles at run time.  The attribute belongs to the
// FUNCTION (shared by every handle and host thread of the process), so it is raised once, to the device limit
// minus the kernel's static shared memory; the launches then ask for what their batch needs.  (Setting it per
// batch would let one handle shrink the limit under a concurrent solve of another handle.)
template <class F>
static cudaError_t raise_smem_limit(F* kernel, int optin) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, (const void*)kernel);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);
}

static cudaError_t raise_smem_limits(int device) {
    int optin = 0;
    cudaError_t e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (e != cudaSuccess) return e;
#define RAISE_(f) if ((e = raise_smem_limit(f, optin)) != cudaSuccess) return e
    RAISE_(k_spmv<1>); RAISE_(k_spmv<2>); RAISE_(k_spmv<3>);
    RAISE_(k_update<1>); RAISE_(k_update<2>); RAISE_(k_update<3>);
    RAISE_(k_direction<1>); RAISE_(k_direction<2>); RAISE_(k_direction<3>);
    RAISE_(k_sym_window); RAISE_(k_sym_coarse); RAISE_(k_sym_elems);
    RAISE_(k_apply<1>); RAISE_(k_apply<2>); RAISE_(k_apply<3>);
    RAISE_(k_update_p<1>); RAISE_(k_update_p<2>); RAISE_(k_update_p<3>);
    RAISE_(k_direction_p<1>); RAISE_(k_direction_p<2>); RAISE_(k_direction_p<3>);
#if RVE_APPLY_PERSISTENT
    RAISE_(k_apply_p<1>); RAISE_(k_apply_p<2>); RAISE_(k_apply_p<3>);
#endif
    RAISE_(k_galerkin1_t);
    RAISE_(k_vc_mid_s<1>); RAISE_(k_vc_mid_s<2>); RAISE_(k_vc_mid_s<3>);
#undef RAISE_
    return cudaSuccess;
}

struct PcgArgs {
    WinTab wt; Sell sl; Levels LV;
    int n_win, n_sys, K, precond;
    double rtol2;
    size_t sm_spmv, sm_upd, sm_dir, sm_apply;
    int op;
};

static MfTab mf_tab(rve_batch_s* h) {
    MfTab mf;
    mf.win_we0 = h->win_we0.p; mf.win_nwe = h->win_nwe.p; mf.win_sw = h->win_sw.p; mf.row_nadj = h->row_nadj.p;
    mf.we_idx = h->we_idx.p; mf.we_geo = h->we_geo.p;
    return mf;
}

// persistent bulk-copy-pipelined vector kernels (rve_stream.cuh); RVE_PIPELINE=0 selects the one-window-per-CTA kernels
static bool use_pipeline() {
    static const bool on = [] { const char* e = std::getenv("RVE_PIPELINE"); return !(e && e[0] == '0'); }();
    return on;
}

template <int K>
static void launch_spmv(rve_batch_s* h, const PcgArgs& a, const double* p, double* q, int* active, double* part, double* sc, cudaStream_t st) {
#if RVE_APPLY_PERSISTENT
    if (a.op == RVE_OPERATOR_MATFREE && use_pipeline()) {
        int& per_sm = h->applyp_blocks[K];
        if (per_sm == 0) {
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_apply_p<K>, RVE_APPLY_THREADS, a.sm_apply);
            if (per_sm < 1) per_sm = 1;
            if (std::getenv("RVE_DEBUG_TIMING")) fprintf(stderr, "[k_apply_p<%d>] %d CTA(s) per SM, %zu bytes of dynamic shared memory\n", K, per_sm, a.sm_apply);
        }
        k_apply_p<K><<<std::min(a.n_win, per_sm * h->n_sm), RVE_APPLY_THREADS, a.sm_apply, st>>>(a.wt, a.sl, mf_tab(h), (const double2*)p, (double2*)q,
                                                                                             h->max_tile, h->max_slots, part, a.n_win);
    } else
#endif
    if (a.op == RVE_OPERATOR_MATFREE)
        k_apply<K><<<a.n_win, RVE_APPLY_THREADS, a.sm_apply, st>>>(a.wt, a.sl, mf_tab(h), (const double2*)p, (double2*)q, h->max_tile,
                                                                   h->max_slots, part);
    else
        k_spmv<K><<<a.n_win, RVE_SPMV_THREADS, a.sm_spmv, st>>>(a.wt, a.sl, (const double2*)h->bval.p, (const double2*)p, (double2*)q, part);
    k_fin_alpha<<<cdiv(a.n_sys, 8), 256, 0, st>>>(a.n_sys, K, a.wt.sys_win0, part, active, sc, h->win_scal.p);
}

template <int K>
static void launch_update(rve_batch_s* h, const PcgArgs& a, int init, int iter, cudaStream_t st) {
    if (use_pipeline()) {
        const int cn_pad = ((h->max_cn + 3) & ~3) + 4;
        const size_t sm = RVE_PIPE_STAGES(K) * updp_stage_bytes(K, cn_pad, a.precond != 0);
        int& per_sm = h->updp_blocks[K][a.precond ? 1 : 0];
        if (per_sm == 0) {
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_update_p<K>, RVE_UPDP_THREADS(K) + 32, sm);
            if (per_sm < 1) per_sm = 1;
            if (std::getenv("RVE_DEBUG_TIMING")) fprintf(stderr, "[k_update_p<%d>] %d CTA(s) per SM, %zu bytes of dynamic shared memory\n", K, per_sm, sm);
        }
        const int grid = std::min(a.n_win, per_sm * h->n_sm);
        k_update_p<K><<<grid, RVE_UPDP_THREADS(K) + 32, sm, st>>>(a.wt, h->win_desc.p, a.n_win, (const double2*)h->p.p, (const double2*)h->q.p,
                                                              (double2*)h->x.p, (double2*)h->r.p, h->dinv.p, h->rowinfo.p, h->cn_node.p,
                                                              h->cn_beg.p, h->cn_ent.p, a.LV.rc[0], a.LV.g[0].G, h->part.p, h->win_scal.p, init,
                                                              a.precond, cn_pad);
    } else {
        k_update<K><<<a.n_win, RVE_UPD_THREADS(K), a.sm_upd, st>>>(a.wt, (const double2*)h->p.p, (const double2*)h->q.p, (double2*)h->x.p,
                                                                   (double2*)h->r.p, h->dinv.p, h->rowinfo.p, h->cn_node.p, h->cn_beg.p,
                                                                   h->cn_ent.p, a.LV.rc[0], a.LV.g[0].G, h->part.p, h->sc.p, init, a.precond);
    }
    k_fin_update<<<cdiv(a.n_sys, 8), 256, 0, st>>>(a.n_sys, K, a.wt.sys_win0, h->part.p, h->active.p, a.wt.win_active, h->sc.p, init, iter,
                                                   a.rtol2, a.precond, h->d_iters.p, h->n_active.p, h->win_scal.p);
}

// coarse V-cycle: the large levels (always level 1) as grid-wide kernels (thread per coarse node), the small
// ones in one CTA per system (k_vc_mid)
template <int K>
static void launch_vcycle(rve_batch_s* h, const PcgArgs& a, int first, cudaStream_t st) {
    const Levels& LV = a.LV;
    // thread per (system, node); small grids take 128-thread blocks so that the last block of a system is not mostly idle
    auto bs = [&](int l) { return LV.g[l].G < 2048 ? 128 : 256; };
    auto grid = [&](int l) { return dim3((unsigned)((LV.g[l].G + bs(l) - 1) / bs(l)), (unsigned)a.n_sys); };
    int W = 1;                                         // number of wide levels; the coarsest level stays in k_vc_mid
    while (W < LV.L - 1 && LV.g[W].G >= RVE_WIDE_NODES) ++W;
    k_vc_down1<K><<<grid(0), bs(0), 0, st>>>(LV, 0, h->active.p);
    for (int l = 1; l < W; ++l) {
        k_vc_restrict<K><<<grid(l), bs(l), 0, st>>>(LV, l, h->active.p);
        k_vc_down1<K><<<grid(l), bs(l), 0, st>>>(LV, l, h->active.p);
    }
    int nk = 2 + 2 * (W - 1);
    if (LV.L > 1) {
        const int Gm = LV.g[W].G;
        // one CTA per system, as few idle threads as the largest level allows: more systems resident per SM (the kernel is latency-bound)
#if RVE_VC_MID_SMEM
        k_vc_mid_s<K><<<a.n_sys, std::min(320, std::max(64, (Gm + 31) & ~31)), vc_mid_smem(LV, W, K), st>>>(LV, W, h->active.p);
#else
        k_vc_mid<K><<<a.n_sys, std::min(320, std::max(64, (Gm + 31) & ~31)), 0, st>>>(LV, W, h->active.p);
#endif
        for (int l = W - 1; l >= 0; --l) {
            const float* src = (l + 1 == W) ? LV.xc[W] : LV.tc[l + 1];
            k_vc_prolong<K><<<grid(l), bs(l), 0, st>>>(LV, l, src, h->active.p);
            if (l > 0) k_vc_upl<K><<<grid(l), bs(l), 0, st>>>(LV, l, h->active.p);
        }
        nk += 1 + W + (W - 1);
    }
    k_vc_up1<K><<<grid(0), bs(0), 0, st>>>(LV, h->active.p, h->cpart.p);
    const int nb = (int)grid(0).x * (bs(0) / 32);      // per-warp partials of rc.z1
    k_fin_beta<<<cdiv(a.n_sys, 8), 256, 0, st>>>(a.n_sys, K, nb, a.wt.sys_win0, h->cpart.p, h->active.p, h->sc.p, first, h->win_scal.p);
    h->launches += nk + 1;
}

template <int K>
static void launch_direction(rve_batch_s* h, const PcgArgs& a, cudaStream_t st) {
    if (use_pipeline()) {
        const int cn_pad = ((h->max_cn + 3) & ~3) + 4;
        const size_t sm = RVE_DIRP_STAGES * dirp_stage_bytes(K, cn_pad, a.precond != 0);
        int& per_sm = h->dirp_blocks[K][a.precond ? 1 : 0];
        if (per_sm == 0) {
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_direction_p<K>, RVE_UPDP_THREADS(K) + 32, sm);
            if (per_sm < 1) per_sm = 1;
            if (std::getenv("RVE_DEBUG_TIMING")) fprintf(stderr, "[k_direction_p<%d>] %d CTA(s) per SM, %zu bytes of dynamic shared memory\n", K, per_sm, sm);
        }
        const int grid = std::min(a.n_win, per_sm * h->n_sm);
        k_direction_p<K><<<grid, RVE_UPDP_THREADS(K) + 32, sm, st>>>(a.wt, h->win_desc.p, a.n_win, (double2*)h->p.p, (const double2*)h->r.p,
                                                                      h->dinv.p, h->rowinfo.p, h->cn_node.p, (const cvec*)a.LV.tc[0],
                                                                      a.LV.g[0].G, h->win_scal.p, a.precond, cn_pad);
        return;
    }
    k_direction<K><<<a.n_win, RVE_DIR_THREADS(K), a.sm_dir, st>>>(a.wt, (double2*)h->p.p, (const double2*)h->r.p, h->dinv.p, h->rowinfo.p,
                                                      h->cn_node.p, (const cvec*)a.LV.tc[0], a.LV.g[0].G, h->active.p,
                                                      h->sc.p, a.precond);
}

#define K_DISPATCH(K, call)              \
    do {                                 \
        if ((K) == 1) { call(1); }       \
        else if ((K) == 2) { call(2); }  \
        else { call(3); }                \
    } while (0)

#ifndef RVE_COARSEST_CELLS
#define RVE_COARSEST_CELLS 2
#endif
// per-batch solver state (sizes known after the numbering): vectors of the PCG scalars, the coarse hierarchy
static int alloc_solver_state(rve_batch_s* h, int n_sys, int model, int ncx, int ncy) {
    h->have_bval = false;                       // SELL values: rve_assemble (assembled operator) or on demand
    CU(h->dinv.alloc((size_t)h->R)); CU(h->wmean.alloc(h->R));
    CU(h->active.alloc(n_sys)); CU(h->cnt.alloc(n_sys)); CU(h->d_iters.alloc(n_sys)); CU(h->n_active.alloc(1));
    CU(h->part.alloc((size_t)h->n_win * 8));  CU(h->sc.alloc((size_t)n_sys * 4 * SC_N));
    CU(h->Mx.alloc((size_t)n_sys * 16)); CU(h->Gaff.alloc((size_t)n_sys * 16)); CU(h->aaff.alloc((size_t)n_sys * 8));
    CU(h->acc.alloc((size_t)n_sys * ACC_PER_SYS));
    // coarse hierarchy
    Levels& LV = h->LV;
    memset(&LV, 0, sizeof(Levels));
    int lcx = ncx, lcy = ncy, L = 0;
    static const int maxlev = [] { const char* e = std::getenv("RVE_MAXLEVELS"); const int v = e ? std::atoi(e) : RVE_MAXLEV; return v < 1 ? 1 : (v > RVE_MAXLEV ? RVE_MAXLEV : v); }();
    while (L < maxlev) {
        Grid& g = LV.g[L];
        g.ncx = lcx; g.ncy = lcy; g.nx = lcx + 1; g.ny = lcy + 1; g.G = g.nx * g.ny;
        g.wrap = (model == RVE_MODEL_PER);
        if (model == RVE_MODEL_LIN || model == RVE_MODEL_DNN) { g.lox = 1; g.loy = 1; g.hix = lcx - 1; g.hiy = lcy - 1; }
        else if (model == RVE_MODEL_PER) { g.lox = 0; g.loy = 0; g.hix = lcx - 1; g.hiy = lcy - 1; }
        else { g.lox = 0; g.loy = 0; g.hix = lcx; g.hiy = lcy; }
        CU(h->lvA[L].alloc((size_t)n_sys * 100 * g.G)); CU(h->lvD[L].alloc((size_t)n_sys * 4 * g.G));
        CU(h->lvR[L].alloc((size_t)n_sys * g.G * RVE_MAX_RHS * 2)); CU(h->lvX[L].alloc((size_t)n_sys * g.G * RVE_MAX_RHS * 2));
        CU(h->lvT[L].alloc((size_t)n_sys * g.G * RVE_MAX_RHS * 2));
        CU(h->lvA4[L].alloc((size_t)n_sys * 25 * g.G)); CU(h->lvH4[L].alloc((size_t)n_sys * 25 * g.G));
        LV.A[L] = h->lvA[L].p; LV.dinv[L] = h->lvD[L].p; LV.rc[L] = h->lvR[L].p; LV.xc[L] = h->lvX[L].p; LV.tc[L] = h->lvT[L].p;
        LV.A4[L] = h->lvA4[L].p; LV.H4[L] = h->lvH4[L].p;
        ++L;
        // (stopping the hierarchy at >= 6 cells per side leaves the iteration counts of per / lin / dnn unchanged and saves
        // a few latency-bound phases per V-cycle, but the MR model -- unconstrained stiffness, rigid-body modes -- needs the deep
        // levels: 69.8 -> 85.0 iterations on the full meshes, r02i; the hierarchy therefore goes down to 2 x 2 cells)
        if ((lcx % 2) || (lcy % 2) || lcx / 2 < RVE_COARSEST_CELLS || lcy / 2 < RVE_COARSEST_CELLS) break;
        lcx /= 2; lcy /= 2;
    }
    LV.L = L;
    CU(h->cpart.alloc((size_t)n_sys * ((LV.g[0].G + 127) / 128 + 1) * 8 * 4));      // per-warp partials of k_vc_up1
    return 0;
}

extern "C" {

int rve_create(int device, rve_handle_t* out) {
    if (!out) return RVE_E_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_err = std::string("no CUDA device: ") + cudaGetErrorString(e);
        return RVE_E_CUDA;
    }
    if (device < 0 || device >= ndev) { g_create_err = "device index out of range"; return RVE_E_ARG; }
    if ((e = cudaSetDevice(device)) != cudaSuccess) { g_create_err = cudaGetErrorString(e); return RVE_E_CUDA; }
    // the symbolic phase allocates and frees several MB of work vectors per RVE on every call: keep freed
    // memory in the process heap (no mmap/munmap round trips, no page faults on re-use)
    static bool heap_tuned = false;
    if (!heap_tuned) {
        mallopt(M_MMAP_THRESHOLD, 32 << 20);   // glibc's upper limit for the threshold
        mallopt(M_TRIM_THRESHOLD, 1 << 30);
        mallopt(M_TOP_PAD, 64 << 20);
        heap_tuned = true;
    }
    if ((e = raise_smem_limits(device)) != cudaSuccess) { g_create_err = std::string("shared-memory opt-in: ") + cudaGetErrorString(e); return RVE_E_CUDA; }
    rve_batch_s* h = new rve_batch_s();
    h->device = device;
    cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, device);
    if (h->n_sm < 1) h->n_sm = 148;
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        g_create_err = cudaGetErrorString(e); delete h; return RVE_E_CUDA;
    }
    for (int i = 0; i < 4; ++i) cudaEventCreate(&h->ev[i]);
    cudaEventCreateWithFlags(&h->ev_block, cudaEventBlockingSync | cudaEventDisableTiming);
    cudaMallocHost((void**)&h->h_nactive, sizeof(int));
    memset(&h->LV, 0, sizeof(Levels));
    *out = h;
    return 0;
}

int rve_destroy(rve_handle_t h) {
    if (!h) return RVE_E_ARG;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (int i = 0; i < 4; ++i) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    if (h->ev_block) cudaEventDestroy(h->ev_block);
    for (cudaEvent_t e : h->evpool) cudaEventDestroy(e);
    for (cudaEvent_t e : h->evdbg) cudaEventDestroy(e);
    if (h->h_nactive) cudaFreeHost(h->h_nactive);
    h->arena.release();
    cudaStream_t s = h->stream;
    delete h;
    if (s) cudaStreamDestroy(s);
    return 0;
}

const char* rve_last_error(rve_handle_t h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int rve_set_batch(rve_handle_t h, int32_t n_sys, const int64_t* vert_off, const int64_t* tri_off,
                  const double* xy, const int32_t* tri, const int32_t* region, const double* mat,
                  int32_t model, int32_t vref_nb, const double* vref_box) {
    H_CHECK(h);
    if (n_sys <= 0 || !vert_off || !tri_off || !xy || !tri || !region || !mat) FAIL(RVE_E_ARG, "null/empty argument");
    if (model < RVE_MODEL_LIN || model > RVE_MODEL_MR) FAIL(RVE_E_ARG, "unknown model");
    if (vref_nb != 0 && vref_nb < 2) FAIL(RVE_E_ARG, "vref_nb must be 0 or >= 2");
    auto t0 = std::chrono::steady_clock::now();
    h->have_batch = h->assembled = h->solved = h->have_samples = false;
    h->n_sys = n_sys; h->model = model; h->vref_nb = vref_nb;
    h->nvref = vref_nb > 0 ? 4 * (vref_nb - 1) + 1 : 0;
    for (int k = 0; k < 8; ++k) h->mat.v[k] = mat[k];
    if (n_sys > 65535) FAIL(RVE_E_ARG, "more than 65535 systems in one batch: split it");
    if (vert_off[0] != 0 || tri_off[0] != 0) FAIL(RVE_E_ARG, "offset arrays must start at 0");
    for (int s = 0; s < n_sys; ++s)
        if (vert_off[s + 1] <= vert_off[s] || tri_off[s + 1] <= tri_off[s]) FAIL(RVE_E_MESH, "system " + std::to_string(s) + ": empty mesh");
    const long long NV = vert_off[n_sys], NT = tri_off[n_sys];
    if (NV + 3 * NT + n_sys > 2000000000LL || 6 * NT > 2000000000LL) FAIL(RVE_E_ARG, "batch too large for 32-bit indexing: split it");
    cudaStream_t st = h->stream;
    std::vector<SysInfo>& I = h->info;
    I.assign(n_sys, SysInfo());
    h->voff.resize(n_sys + 1); h->toff.resize(n_sys + 1);
    for (int s = 0; s <= n_sys; ++s) { h->voff[s] = (int)vert_off[s]; h->toff[s] = (int)tri_off[s]; }
    for (int s = 0; s < n_sys; ++s) { I[s].nv = h->voff[s + 1] - h->voff[s]; I[s].nt = h->toff[s + 1] - h->toff[s]; }

    // ---- the raw meshes go to the GPU as they are (through the pinned staging arena)
    {
        size_t bytes = PinnedArena::need(2 * NV, 8) + PinnedArena::need(3 * NT, 4) + PinnedArena::need(NT, 4);
        CU(stream_wait(h));          // the previous batch's copies out of the arena are done
        CU(h->arena.reserve(bytes));
    }
    PinnedArena& AR = h->arena;
    Span<double> a_xy = AR.take<double>(2 * NV);
    Span<int> a_tri = AR.take<int>(3 * NT), a_reg = AR.take<int>(NT);
    parallel_for(n_sys, [&](int s) {
        memcpy(a_xy.p + 2 * vert_off[s], xy + 2 * vert_off[s], 2 * sizeof(double) * (size_t)I[s].nv);
        memcpy(a_tri.p + 3 * tri_off[s], tri + 3 * tri_off[s], 3 * sizeof(int) * (size_t)I[s].nt);
        memcpy(a_reg.p + tri_off[s], region + tri_off[s], sizeof(int) * (size_t)I[s].nt);
    });
    UPL(h->d_xy, a_xy); UPL(h->d_tri, a_tri); UPL(h->d_region, a_reg); UPL(h->d_voff, h->voff); UPL(h->d_toff, h->toff);
    h->E = NT;
    DsMesh DM;
    DM.xy = h->d_xy.p; DM.tri = h->d_tri.p; DM.region = h->d_region.p; DM.voff = h->d_voff.p; DM.toff = h->d_toff.p;

    // ---- (A) orientation, bounding boxes, edges
    CU(h->scrA.alloc((size_t)(2 * NV + n_sys + 9 * NT + 64)));
    int* vptr = h->scrA.p; int* vfill = vptr + NV + n_sys; int* he_vb = vfill + NV; int* he_slot = he_vb + 3 * NT; int* emid = he_slot + 3 * NT;
    CU(h->elem_reg.alloc((size_t)NT)); CU(h->box.alloc(4 * (size_t)n_sys)); CU(h->sys_cnt.alloc(8 * (size_t)n_sys)); CU(h->sys_ext.alloc(n_sys));
    CU(cudaMemsetAsync(h->sys_cnt.p, 0, 8 * (size_t)n_sys * sizeof(int), st));
    k_ds_edges<<<n_sys, DS_THREADS, 0, st>>>(DM, vptr, vfill, he_vb, he_slot, emid, h->elem_reg.p, h->box.p, h->sys_cnt.p, h->sys_ext.p);
    h->launches += 1;
    std::vector<int>& cnt = h->h_cnt; std::vector<double>& hext = h->h_ext; std::vector<double>& hbox = h->h_box;
    cnt.resize(8 * (size_t)n_sys); hext.resize(n_sys); hbox.resize(4 * (size_t)n_sys);
    CU(cudaMemcpyAsync(cnt.data(), h->sys_cnt.p, cnt.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(hext.data(), h->sys_ext.p, hext.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(hbox.data(), h->box.p, hbox.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(stream_wait(h));
    CU(cudaGetLastError());
    auto check_sys = [&]() -> int {
        for (int s = 0; s < n_sys; ++s)
            if (cnt[8 * (size_t)s + DC_ERR]) { h->err = "system " + std::to_string(s) + ": " + ds_error_text(cnt[8 * (size_t)s + DC_ERR]); return RVE_E_MESH; }
        return 0;
    };
    if (check_sys()) return RVE_E_MESH;
    double ext = 0;
    const double Lx = hbox[2], Ly = hbox[3];
    h->node_base.assign(n_sys + 1, 0); h->bnd_base.assign(n_sys + 1, 0); h->bedge_base.assign(n_sys + 1, 0);
    for (int s = 0; s < n_sys; ++s) {
        I[s].ne = cnt[8 * (size_t)s + DC_NE]; I[s].nbe = cnt[8 * (size_t)s + DC_NBE]; I[s].nb = cnt[8 * (size_t)s + DC_NB];
        I[s].nn = I[s].nv + I[s].ne;
        ext += hext[s] / n_sys;        // batch mean of the mean element extent
        if (std::fabs(hbox[4 * (size_t)s + 2] - Lx) > 1e-9 * Lx || std::fabs(hbox[4 * (size_t)s + 3] - Ly) > 1e-9 * Ly)
            FAIL(RVE_E_MESH, "all systems of a batch must share the box size");
        h->node_base[s + 1] = h->node_base[s] + I[s].nn;
        h->bnd_base[s + 1] = h->bnd_base[s] + I[s].nb; h->bedge_base[s + 1] = h->bedge_base[s] + I[s].nbe;
    }
    h->N = h->node_base[n_sys]; h->NBN = h->bnd_base[n_sys]; h->NBE = h->bedge_base[n_sys];
    const long long N = h->N;
    const int ncx = pick_nc(Lx, ext), ncy = pick_nc(Ly, ext);
    h->ncx = ncx; h->ncy = ncy;
    for (int k = 0; k < 4; ++k) h->vref_box[k] = vref_box ? vref_box[k] : hbox[k];
    // Morton rank of every level-1 cell
    {
        auto spread = [](unsigned v) -> uint64_t {
            uint64_t x = v & 0xffffu;
            x = (x | (x << 8)) & 0x00ff00ffull; x = (x | (x << 4)) & 0x0f0f0f0full;
            x = (x | (x << 2)) & 0x33333333ull; x = (x | (x << 1)) & 0x55555555ull;
            return x;
        };
        std::vector<std::pair<uint64_t, int>> ck((size_t)ncx * ncy);
        for (int cj = 0; cj < ncy; ++cj)
            for (int ci = 0; ci < ncx; ++ci) ck[(size_t)cj * ncx + ci] = {spread((unsigned)ci) | (spread((unsigned)cj) << 1), cj * ncx + ci};
        std::sort(ck.begin(), ck.end());
        h->cellrank.resize(ck.size());
        for (size_t k = 0; k < ck.size(); ++k) h->cellrank[ck[k].second] = (int)k;
    }
    const size_t ncells = (size_t)ncx * ncy;
    auto tp1 = std::chrono::steady_clock::now();

    // ---- (B) nodes, boundary lists, constraint map, row order, elements around rows, capacities, windows; Vref samples
    const size_t nvs = (size_t)n_sys * h->nvref;
    UPL(h->d_node_base, h->node_base); UPL(h->d_bnd_base, h->bnd_base); UPL(h->d_bedge_base, h->bedge_base); UPL(h->d_cellrank, h->cellrank);
    std::vector<double> vb(h->vref_box, h->vref_box + 4);
    UPL(h->d_vbox, vb);
    CU(h->node_xy.alloc(2 * (size_t)N)); CU(h->node_va.alloc((size_t)N)); CU(h->node_vb.alloc((size_t)N)); CU(h->node_sys.alloc((size_t)N));
    CU(h->node_bnd.alloc((size_t)N)); CU(h->node_row.alloc((size_t)N)); CU(h->elem_node.alloc(6 * (size_t)NT)); CU(h->elem_sys.alloc((size_t)NT));
    CU(h->bnd_sys.alloc((size_t)h->NBN)); CU(h->bnd_node.alloc((size_t)h->NBN)); CU(h->bnd_s.alloc(model == RVE_MODEL_DNN ? 2 * (size_t)h->NBN : 0));
    CU(h->bedge.alloc(3 * (size_t)h->NBE)); CU(h->bedge_nl.alloc(2 * (size_t)h->NBE));
    CU(h->samp_elem.alloc(nvs)); CU(h->samp_lam.alloc(2 * nvs));
    {
        const size_t words = 14 * (size_t)N + 8 * (size_t)n_sys + 6 * (size_t)NT + (size_t)n_sys * (2 * ncells + 1) + (size_t)N / 4 + (size_t)N / 32 + 256;
        CU(h->scrB.alloc(words));
    }
    DsRows DR;
    {
        int* q = h->scrB.p;
        auto take = [&](size_t n) { int* r = q; q += n; return r; };
        DR.master = take(N); DR.ncell = take(N); DR.pnode = take(N); DR.prow = take(N); DR.nrow0 = take(N); DR.cur = take(N);
        DR.len0 = take(N); DR.newid = take(N); DR.oldid = take(N); DR.row_node_loc = take(N); DR.node_row_loc = take(N);
        DR.adjc = take(N + n_sys); DR.loc_rowptr = take(N + n_sys); DR.loc_adjptr = take(N + n_sys);
        DR.adj = take(6 * NT); DR.ccnt = take((size_t)n_sys * (ncells + 1)); DR.cfill = take((size_t)n_sys * ncells);
        DR.loc_slw = take((size_t)N / 32 + 2 * (size_t)n_sys + 8);
        DR.rowb = (unsigned char*)take((size_t)N / 4 + 8);
        DR.scan3 = he_slot;              // phase-A scratch that is free again
    }
    DR.emid = emid; DR.node_base = h->d_node_base.p; DR.bnd_base = h->d_bnd_base.p; DR.bedge_base = h->d_bedge_base.p;
    DR.cellrank = h->d_cellrank.p; DR.box = h->box.p; DR.vbox = vref_box ? h->d_vbox.p : nullptr;
    DR.model = model; DR.ncx = ncx; DR.ncy = ncy; DR.nside = vref_nb > 0 ? vref_nb - 1 : 0;
    DR.node_xy = h->node_xy.p; DR.node_va = h->node_va.p; DR.node_vb = h->node_vb.p; DR.node_sys = h->node_sys.p; DR.node_bnd = h->node_bnd.p;
    DR.elem_node = h->elem_node.p; DR.elem_sys = h->elem_sys.p; DR.bnd_sys = h->bnd_sys.p; DR.bnd_node = h->bnd_node.p; DR.bnd_s = h->bnd_s.p;
    DR.bedge = h->bedge.p; DR.bedge_nl = h->bedge_nl.p; DR.sys_cnt = h->sys_cnt.p;
    if (model == RVE_MODEL_DNN && DR.nside == 0) FAIL(RVE_E_MESH, "dnn model needs vref_nb > 0");
    k_ds_rows<<<n_sys, DS_THREADS, 0, st>>>(DM, DR);
    h->launches += 1;
    if (h->nvref > 0) {
        k_ds_vref<<<n_sys, 256, 0, st>>>(DM, h->box.p, DR.vbox, DR.nside, he_vb /* free phase-A scratch */, h->samp_elem.p, h->samp_lam.p, h->sys_cnt.p);
        h->launches += 1;
    }
    CU(cudaMemcpyAsync(cnt.data(), h->sys_cnt.p, cnt.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(stream_wait(h));
    CU(cudaGetLastError());
    if (check_sys()) return RVE_E_MESH;

    // prefix sums over the systems
    h->elem_base = h->toff;
    h->row_base.assign(n_sys + 1, 0); h->blk_base.assign(n_sys + 1, 0); h->win_base.assign(n_sys + 1, 0);
    h->slice_base.assign(n_sys + 1, 0); h->adjf_base.assign(n_sys + 1, 0); h->grp_base.assign(n_sys + 1, 0);
    h->max_tile = 0; h->max_cn = 0; h->max_words = 0;
    for (int s = 0; s < n_sys; ++s) {
        I[s].na = cnt[8 * (size_t)s + DC_NA]; I[s].nblk = cnt[8 * (size_t)s + DC_NBLK]; I[s].ngrp = cnt[8 * (size_t)s + DC_NGRP];
        I[s].nadj = cnt[8 * (size_t)s + DC_NADJ];
        I[s].nwin = (I[s].na + RVE_WIN - 1) / RVE_WIN; I[s].nsl = (I[s].na + 31) / 32;
        const long long nr = (long long)h->row_base[s] + I[s].na, nb = h->blk_base[s] + I[s].nblk, ng = h->grp_base[s] + I[s].ngrp,
                        nh = h->adjf_base[s] + I[s].nadj;
        if (nr * 4 > 2000000000LL || nb > 4000000000LL || ng > 4000000000LL || nh > 2000000000LL)
            FAIL(RVE_E_ARG, "batch too large for 32-bit indexing: split it");
        h->row_base[s + 1] = (int)nr; h->blk_base[s + 1] = nb; h->grp_base[s + 1] = ng; h->adjf_base[s + 1] = nh;
        h->win_base[s + 1] = h->win_base[s] + I[s].nwin; h->slice_base[s + 1] = h->slice_base[s] + I[s].nsl;
        h->max_words = std::max(h->max_words, (I[s].na + 31) / 32);
    }
    h->R = h->row_base[n_sys]; h->NB = h->blk_base[n_sys]; h->NG = h->grp_base[n_sys]; h->n_win = h->win_base[n_sys];
    const int n_win = h->n_win, n_slice = h->slice_base[n_sys];
    h->u_blk.resize(n_sys + 1); h->u_grp.resize(n_sys + 1); h->u_adjf.resize(n_sys + 1);
    for (int s = 0; s <= n_sys; ++s) { h->u_blk[s] = (unsigned)h->blk_base[s]; h->u_grp[s] = (unsigned)h->grp_base[s]; h->u_adjf[s] = (unsigned)h->adjf_base[s]; }

    // ---- (C) local numbering -> global row / block / slice index spaces
    UPL(h->d_row_base, h->row_base); UPL(h->sys_win0, h->win_base); UPL(h->d_slice_base, h->slice_base);
    UPL(h->d_blk_base, h->u_blk); UPL(h->d_grp_base, h->u_grp); UPL(h->d_adjf_base, h->u_adjf);
    CU(h->row_node.alloc((size_t)h->R)); CU(h->d_row_ptr.alloc((size_t)h->R + 1)); CU(h->adjf_ptr.alloc((size_t)h->R + 1));
    CU(h->adjf.alloc((size_t)h->adjf_base[n_sys])); CU(h->win_sys.alloc(n_win)); CU(h->win_row0.alloc(n_win)); CU(h->win_n.alloc(n_win));
    CU(h->win_slice0.alloc(n_win)); CU(h->slice_ptr.alloc((size_t)n_slice + 1));
    DsFinish DF;
    DF.node_base = h->d_node_base.p; DF.row_base = h->d_row_base.p; DF.win_base = h->sys_win0.p; DF.slice_base = h->d_slice_base.p;
    DF.blk_base = h->d_blk_base.p; DF.grp_base = h->d_grp_base.p; DF.adjf_base = h->d_adjf_base.p;
    DF.row_node_loc = DR.row_node_loc; DF.node_row_loc = DR.node_row_loc; DF.loc_rowptr = DR.loc_rowptr; DF.loc_adjptr = DR.loc_adjptr;
    DF.loc_slw = DR.loc_slw; DF.oldid = DR.oldid; DF.adjc = DR.adjc; DF.adj = DR.adj;
    DF.node_row = h->node_row.p; DF.row_node = h->row_node.p; DF.row_ptr = h->d_row_ptr.p; DF.adjf_ptr = h->adjf_ptr.p; DF.adjf = h->adjf.p;
    DF.win_sys = h->win_sys.p; DF.win_row0 = h->win_row0.p; DF.win_n = h->win_n.p; DF.win_slice0 = h->win_slice0.p; DF.slice_ptr = h->slice_ptr.p;
    DF.n_sys = n_sys;
    k_ds_finish<<<n_sys, DS_THREADS, 0, st>>>(DM, DF);
    DBG_SYNC("k_ds_finish");
    h->launches += 1;
    auto t1 = std::chrono::steady_clock::now();
    h->t_ms[0] = std::chrono::duration<double, std::milli>(t1 - t0).count();
    if (std::getenv("RVE_DEBUG_TIMING"))
        fprintf(stderr, "[rve_set_batch] n_sys %d: copy + edges %.1f ms, rows + finish (enqueue) %.1f ms\n", n_sys,
                std::chrono::duration<double, std::milli>(tp1 - t0).count(), std::chrono::duration<double, std::milli>(t1 - tp1).count());
    if (int rc = alloc_solver_state(h, n_sys, model, ncx, ncy)) return rc;
    Levels& LV = h->LV;
    // block columns, halo lists, SELL columns, restriction lists
    const long long halo_cap = h->R + h->R / 2 + 4096;
    if (halo_cap > 2000000000LL) FAIL(RVE_E_ARG, "batch too large for 32-bit indexing: split it");
    CU(h->csr_col.alloc((size_t)h->NB)); CU(h->rlen.alloc((size_t)h->R));
    h->have_emap = h->opt_operator == RVE_OPERATOR_ASSEMBLED;     // element -> slot map: only k_assemble reads it (ensure_emap otherwise)
    if (h->have_emap) CU(h->emap.alloc(36 * (size_t)h->E));
    CU(h->sys_blocks.alloc(n_sys)); CU(cudaMemsetAsync(h->sys_blocks.p, 0, n_sys * sizeof(int), st)); CU(h->halo.alloc((size_t)halo_cap)); CU(h->scol.alloc((size_t)h->NG * 32));
    CU(h->win_halo0.alloc(n_win)); CU(h->win_nhalo.alloc(n_win)); CU(h->sym_cnt.alloc(8));
    CU(cudaMemsetAsync(h->sym_cnt.p, 0, 8 * sizeof(int), st));
    const long long cn_cap = h->R + 5LL * n_win + 4096;      // level-1 nodes touched per window are far fewer than its rows
    CU(h->win_cn0.alloc(n_win)); CU(h->win_ncn.alloc(n_win)); CU(h->win_active.alloc((size_t)n_win + 4)); CU(h->cn_node.alloc((size_t)cn_cap)); CU(h->cn_beg.alloc((size_t)cn_cap));
    CU(h->cn_ent.alloc(4 * (size_t)h->R + 8 * (size_t)n_win + 64)); CU(h->rowinfo.alloc((size_t)h->R)); CU(h->win_desc.alloc((size_t)n_win)); CU(h->win_scal.alloc(8 * (size_t)n_win));
    {
        WinTab wt = win_tab(h);
        k_sym_cols<<<n_win * RVE_SYMCOLS_SPLIT, RVE_WIN / RVE_SYMCOLS_SPLIT, 0, st>>>(wt, h->d_row_ptr.p, h->adjf_ptr.p, h->adjf.p, h->elem_node.p, h->node_row.p,
                                               h->csr_col.p, h->rlen.p, h->have_emap ? h->emap.p : nullptr, h->sys_blocks.p, h->sym_cnt.p);
        DBG_SYNC("k_sym_cols");
        const size_t sm = 2 * (size_t)h->max_words * sizeof(unsigned);
        if (sm > 200 * 1024) FAIL(RVE_E_MESH, "system too large for the halo bitmap in shared memory");
        k_sym_window<<<n_win, RVE_WIN, sm, st>>>(wt, h->win_halo0.p, h->win_nhalo.p, h->d_row_ptr.p, h->rlen.p, h->csr_col.p, h->slice_ptr.p,
                                                  h->scol.p, h->halo.p, (int)halo_cap, h->sym_cnt.p);
        DBG_SYNC("k_sym_window");
        const size_t smc = 2 * (size_t)((LV.g[0].G + 31) / 32) * sizeof(unsigned);
        if (smc > 150 * 1024) FAIL(RVE_E_MESH, "level-1 grid too large for the coarse-node bitmap in shared memory");
        k_sym_coarse<<<n_win, RVE_WIN, smc, st>>>(wt, h->win_cn0.p, h->win_ncn.p, h->row_node.p, h->node_xy.p, h->box.p, LV.g[0],
                                                  model, h->rowinfo.p, h->cn_node.p, h->cn_beg.p, h->cn_ent.p, (int)cn_cap, h->sym_cnt.p, h->win_desc.p);
        DBG_SYNC("k_sym_coarse");
        h->launches += 3;
    }
    int scnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // window-element tables of the matrix-free operator (after the halo lists: columns are tile-local)
    CU(h->win_we0.alloc(n_win)); CU(h->win_nwe.alloc(n_win)); CU(h->win_sw.alloc(8 * (size_t)n_win)); CU(h->row_nadj.alloc((size_t)h->R));
    {
        int max_ne = 0;
        for (int s = 0; s < n_sys; ++s) max_ne = std::max(max_ne, I[s].nt);
        const size_t sme = 2 * (size_t)((max_ne + 31) / 32) * sizeof(unsigned);
        if (sme > 150 * 1024) FAIL(RVE_E_MESH, "system too large for the element bitmap in shared memory");
        long long we_cap = 2 * h->E + 32LL * n_win + 1024;
        static const bool skip_mf = [] { const char* e = std::getenv("RVE_SKIP_MATFREE"); return e && e[0] == '1'; }();
        for (int attempt = 0; attempt < 2 && !skip_mf; ++attempt) {
            if (6 * we_cap > 2000000000LL) FAIL(RVE_E_ARG, "batch too large for 32-bit indexing: split it");
            CU(h->we_elem.alloc((size_t)we_cap)); CU(h->we_idx.alloc(6 * (size_t)we_cap));
            CU(cudaMemsetAsync(h->sym_cnt.p + 6, 0, 2 * sizeof(int), st));
            k_sym_elems<<<n_win, RVE_WIN, sme, st>>>(win_tab(h), sell_tab(h), h->d_toff.p, h->adjf_ptr.p, h->adjf.p, h->elem_node.p,
                                                      h->node_row.p, h->win_we0.p, h->win_nwe.p, h->win_sw.p, h->row_nadj.p,
                                                      h->we_elem.p, h->we_idx.p, we_cap, h->sym_cnt.p);
            DBG_SYNC("k_sym_elems");
            h->launches += 1;
            CU(cudaMemcpyAsync(scnt, h->sym_cnt.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
            CU(stream_wait(h));
            if (scnt[2] != 5 || scnt[6] <= we_cap) break;       // capacity was enough (or another error)
            we_cap = scnt[6];                                   // the counter kept running: exact need
            CU(cudaMemsetAsync(h->sym_cnt.p + 2, 0, sizeof(int), st));
        }
        h->n_we = scnt[6]; h->max_slots = scnt[7];
        CU(h->we_geo.alloc(3 * (size_t)std::max<long long>(h->n_we, 1)));
    }
    CU(cudaMemcpyAsync(scnt, h->sym_cnt.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
    std::vector<int> sb(n_sys);
    CU(cudaMemcpyAsync(sb.data(), h->sys_blocks.p, n_sys * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(stream_wait(h));
    CU(cudaGetLastError());
    if (scnt[2] == 2) FAIL(RVE_E_MESH, "device pattern build: a row has more blocks than its element fan allows (non-manifold vertex?)");
    h->true_blocks.assign(sb.begin(), sb.end());
    if (scnt[2] == 3) FAIL(RVE_E_MESH, "device pattern build: halo list overflow");
    if (scnt[2] == 4) FAIL(RVE_E_MESH, "device pattern build: coarse-node list overflow");
    if (scnt[2] == 5) FAIL(RVE_E_MESH, "device pattern build: a window touches more than 1024 elements");
    h->max_tile = scnt[1]; h->max_cn = scnt[5]; h->n_halo = scnt[0];
    if (((size_t)h->max_tile + 8 + h->max_slots) * 3 * sizeof(double2) > 200 * 1024) FAIL(RVE_E_MESH, "window halo too large for shared memory");
    auto t2 = std::chrono::steady_clock::now();
    h->t_ms[1] = std::chrono::duration<double, std::milli>(t2 - t1).count();
    h->have_batch = true;
    return 0;
}

// One topology, n_sys coordinate sets (meshes morphed from one template, deepbnd_b200/mesh/rve_morph.py): the symbolic
// phase runs for system 0 only; its index arrays are replicated with shifted indices, the coordinate-dependent pieces
// (node coordinates, level-1 cells of the rows, restriction lists) are rebuilt for every system.
struct MorphHost {                 // rve_set_batch_morphed: template description + per-system inclusion parameters (host pointers)
    int n_mov, n_incl;
    const int32_t* node; const int32_t* incl; const double* d; const double* dir; const double* centre; const double* prm;
    double r0, R;
};

static int set_batch_shared_impl(rve_handle_t h, int32_t n_sys, int32_t nv, int32_t nt, const double* xy, const int32_t* tri,
                                 const int32_t* region, const double* mat, int32_t model, int32_t vref_nb, const double* vref_box,
                                 const MorphHost* morph) {
    H_CHECK(h);
    if (n_sys <= 0 || nv <= 0 || nt <= 0 || !xy || !tri || !region || !mat) FAIL(RVE_E_ARG, "null/empty argument");
    if (n_sys > 65535) FAIL(RVE_E_ARG, "more than 65535 systems in one batch: split it");
    auto t0 = std::chrono::steady_clock::now();
    const int64_t voff1[2] = {0, nv}, toff1[2] = {0, nt};
    if (int rc = rve_set_batch(h, 1, voff1, toff1, xy, tri, region, mat, model, vref_nb, vref_box)) return rc;
    if (n_sys == 1 && !morph) return 0;
    h->have_batch = false;
    cudaStream_t st = h->stream;
    const SysInfo Y = h->info[0];
    const long long nn = Y.nn, na = Y.na, nb = Y.nb, nbe = Y.nbe, nwin = Y.nwin, nsl = Y.nsl, nblk = Y.nblk, ngrp = Y.ngrp,
                    nadj = Y.nadj, H1 = h->n_halo, WE1 = h->n_we, nvref = h->nvref;
    if (nn * n_sys > 2000000000LL || nblk * n_sys > 4000000000LL || ngrp * n_sys > 4000000000LL || 6 * WE1 * n_sys > 2000000000LL ||
        6LL * nt * n_sys > 2000000000LL)
        FAIL(RVE_E_ARG, "batch too large for 32-bit indexing: split it");
    // ---- the pieces copied from system 0 are only valid if the mesh boundary and the Vref interface do not move
    //      (morphing on the device never moves them: the template's moving set excludes those nodes)
    if (!morph) {
        std::vector<int> nbnd(nv), se(nvref);
        CU(cudaMemcpyAsync(nbnd.data(), h->node_bnd.p, nv * sizeof(int), cudaMemcpyDeviceToHost, st));
        if (nvref) CU(cudaMemcpyAsync(se.data(), h->samp_elem.p, nvref * sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(stream_wait(h));
        std::vector<char> fixed(nv, 0);
        for (int v = 0; v < nv; ++v) fixed[v] = nbnd[v] >= 0;
        for (int k = 0; k < nvref; ++k)
            for (int c = 0; c < 3; ++c) fixed[tri[3 * (size_t)se[k] + c]] |= 2;      // candidates: on the interface if they stay put
        const double tol = 1e-12 * std::max(h->h_box[2], h->h_box[3]);
        const double* vb = h->vref_box;
        for (int v = 0; v < nv; ++v)
            if ((fixed[v] & 2) && !(fixed[v] & 1)) {          // vertex of a sampled element: fixed if it lies on the Vref outline
                const double x = xy[2 * (size_t)v], y = xy[2 * (size_t)v + 1];
                const bool on = ((std::fabs(x - vb[0]) < 1e-9 * vb[2] || std::fabs(x - vb[0] - vb[2]) < 1e-9 * vb[2]) && y > vb[1] - 1e-9 * vb[3] && y < vb[1] + vb[3] * (1 + 1e-9)) ||
                                ((std::fabs(y - vb[1]) < 1e-9 * vb[3] || std::fabs(y - vb[1] - vb[3]) < 1e-9 * vb[3]) && x > vb[0] - 1e-9 * vb[2] && x < vb[0] + vb[2] * (1 + 1e-9));
                fixed[v] = on ? 1 : 0;
            }
        for (int s = 1; s < n_sys; ++s)
            for (int v = 0; v < nv; ++v)
                if (fixed[v] & 1) {
                    const double dx = xy[2 * ((size_t)s * nv + v)] - xy[2 * (size_t)v], dy = xy[2 * ((size_t)s * nv + v) + 1] - xy[2 * (size_t)v + 1];
                    if (std::fabs(dx) > tol || std::fabs(dy) > tol)
                        FAIL(RVE_E_MESH, "system " + std::to_string(s) + ": a vertex of the mesh boundary / Vref outline differs from system 0 (shared topology needs them fixed)");
                }
    }
    // ---- replicate the index arrays.  The system-0 copy of every array goes to one grow-only scratch buffer first, so
    //      that the arrays themselves can grow in place: no cudaMalloc / cudaFree (device-wide synchronisation) once the
    //      handle has seen a batch of this size
    {
        size_t need = 0;
        auto acc = [&](long long cnt, size_t elem) { need += (((size_t)cnt * elem + 255) & ~(size_t)255); };
        acc(2 * nn, 4); acc(2 * nn, 4); acc(6LL * nt, 4); acc(nb, 4); acc(3 * nbe, 4); acc(nvref, 4); acc(na, 4); acc(nadj, 4); acc(nblk, 4);
        acc(6 * nwin, 4); acc(2 * nwin, 4); acc(H1, 4); acc(WE1, 4); acc(3 * (na + 1) + nsl + 1, 4);
        acc(nt, 1); acc(36LL * nt, 1); acc(2 * na, 1); acc(8 * nwin, 1); acc(32 * ngrp, 2); acc(6 * WE1, 4);
        acc(2 * nbe + 2 * nvref + 4 + 2 * nb, 8); acc(2LL * nv, 8);
        CU(h->rep_scratch.alloc(need + 64 * 256));
    }
    size_t rep_off = 0;
    auto stash = [&](const void* src, size_t bytes) -> void* {       // system-0 data -> scratch (device to device)
        void* dst = h->rep_scratch.p + rep_off;
        rep_off += (bytes + 255) & ~(size_t)255;
        if (bytes) cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st);
        return dst;
    };
    auto rep_i = [&](DBuf<int>& d, long long n1, long long add, int keep_neg) -> cudaError_t {
        const int* src = (const int*)stash(d.p, (size_t)n1 * sizeof(int));
        cudaError_t e = d.alloc((size_t)(n1 * n_sys));
        if (e != cudaSuccess || n1 == 0) return e;
        k_rep_add<int><<<cdiv(n1 * n_sys, 256), 256, 0, st>>>(d.p, src, n1, n_sys, add, keep_neg);
        return cudaSuccess;
    };
    auto rep_p = [&](DBuf<unsigned>& d, long long n1, long long add) -> cudaError_t {
        const unsigned* src = (const unsigned*)stash(d.p, (size_t)(n1 + 1) * sizeof(unsigned));
        cudaError_t e = d.alloc((size_t)(n1 * n_sys + 1));
        if (e != cudaSuccess) return e;
        k_rep_ptr<<<cdiv(n1 * n_sys + 1, 256), 256, 0, st>>>(d.p, src, n1, n_sys, add);
        return cudaSuccess;
    };
    auto rep_s = [&](DBuf<int>& d, long long n1) -> cudaError_t {
        cudaError_t e = d.alloc((size_t)(n1 * n_sys));
        if (e != cudaSuccess || n1 == 0) return e;
        k_rep_sys<<<cdiv(n1 * n_sys, 256), 256, 0, st>>>(d.p, n1, n_sys);
        return cudaSuccess;
    };
    // plain copies (no index shift), any element type: byte-wise
#define REP_COPY(d, n1)                                                                                        \
    do {                                                                                                       \
        const long long bytes_ = (long long)(n1) * (long long)sizeof(*(d).p);                                  \
        const unsigned char* src_ = (const unsigned char*)stash((d).p, (size_t)bytes_);                        \
        CU((d).alloc((size_t)((n1) * n_sys)));                                                                 \
        if (bytes_ > 0) k_rep_add<unsigned char><<<cdiv(bytes_ * n_sys, 256), 256, 0, st>>>((unsigned char*)(d).p, src_, bytes_, n_sys, 0, 0); \
    } while (0)
    CU(rep_i(h->node_va, nn, 0, 0)); CU(rep_i(h->node_vb, nn, 0, 0));
    CU(rep_i(h->node_bnd, nn, nb, 1)); CU(rep_i(h->node_row, nn, na, 1)); CU(rep_i(h->elem_node, 6LL * nt, nn, 0));
    CU(rep_i(h->bnd_node, nb, nn, 0)); CU(rep_i(h->bedge, 3 * nbe, nn, 0)); CU(rep_i(h->samp_elem, nvref, nt, 0));
    CU(rep_i(h->row_node, na, nn, 0)); CU(rep_i(h->adjf, nadj, nt, 0)); CU(rep_i(h->csr_col, nblk, na, 0));
    CU(rep_i(h->win_row0, nwin, na, 0)); CU(rep_i(h->win_n, nwin, 0, 0)); CU(rep_i(h->win_slice0, nwin, nsl, 0));
    CU(rep_i(h->win_halo0, nwin, H1, 0)); CU(rep_i(h->win_nhalo, nwin, 0, 0)); CU(rep_i(h->win_we0, nwin, WE1, 0));
    CU(rep_i(h->win_nwe, nwin, 0, 0)); CU(rep_i(h->halo, H1, na, 0)); CU(rep_i(h->we_elem, WE1, nt, 1));
    CU(rep_s(h->node_sys, nn)); CU(rep_s(h->elem_sys, nt)); CU(rep_s(h->bnd_sys, nb)); CU(rep_s(h->win_sys, nwin));
    CU(rep_p(h->d_row_ptr, na, nblk)); CU(rep_p(h->adjf_ptr, na, nadj)); CU(rep_p(h->slice_ptr, nsl, ngrp));
    REP_COPY(h->elem_reg, (long long)nt); REP_COPY(h->rlen, na); REP_COPY(h->row_nadj, na);
    if (h->have_emap) REP_COPY(h->emap, 36LL * nt);
    REP_COPY(h->win_sw, 8 * nwin); REP_COPY(h->scol, 32 * ngrp); REP_COPY(h->we_idx, 6 * WE1);
    REP_COPY(h->bedge_nl, 2 * nbe); REP_COPY(h->samp_lam, 2 * nvref); REP_COPY(h->box, 4LL);
    if (model == RVE_MODEL_DNN) REP_COPY(h->bnd_s, 2 * nb);
#undef REP_COPY
    // ---- host bookkeeping
    h->n_sys = n_sys;
    h->info.assign(n_sys, Y);
    h->true_blocks.assign(n_sys, h->true_blocks[0]);
    auto mult = [&](std::vector<int>& v, long long n1) { v.resize(n_sys + 1); for (int s = 0; s <= n_sys; ++s) v[s] = (int)(s * n1); };
    auto multl = [&](std::vector<long long>& v, long long n1) { v.resize(n_sys + 1); for (int s = 0; s <= n_sys; ++s) v[s] = s * n1; };
    mult(h->voff, nv); mult(h->toff, nt); mult(h->node_base, nn); mult(h->bnd_base, nb); mult(h->bedge_base, nbe);
    mult(h->row_base, na); mult(h->win_base, nwin); mult(h->slice_base, nsl);
    multl(h->blk_base, nblk); multl(h->grp_base, ngrp); multl(h->adjf_base, nadj);
    h->elem_base = h->toff;
    h->h_box.resize(4 * (size_t)n_sys);
    for (int s = 1; s < n_sys; ++s) for (int k = 0; k < 4; ++k) h->h_box[4 * (size_t)s + k] = h->h_box[k];
    h->N = nn * n_sys; h->E = (long long)nt * n_sys; h->R = na * n_sys; h->NB = nblk * n_sys; h->NG = ngrp * n_sys;
    h->NBN = nb * n_sys; h->NBE = nbe * n_sys; h->n_win = (int)(nwin * n_sys); h->n_halo = H1 * n_sys; h->n_we = WE1 * n_sys;
    UPL(h->sys_win0, h->win_base); UPL(h->d_row_base, h->row_base); UPL(h->d_bedge_base, h->bedge_base);
    UPL(h->d_node_base, h->node_base); UPL(h->d_bnd_base, h->bnd_base); UPL(h->d_toff, h->toff);
    // ---- coordinates of every system
    {
        const size_t cnt = 2 * (size_t)nv * n_sys;
        if (morph) {
            // template coordinates (already on the device as system 0) replicated, moving nodes recomputed per system
            const size_t nm = (size_t)morph->n_mov, ni = (size_t)morph->n_incl;
            CU(h->arena.reserve(PinnedArena::need(nm, 4) * 2 + PinnedArena::need(nm, 8) + PinnedArena::need(2 * nm, 8) +
                                PinnedArena::need(2 * ni, 8) + PinnedArena::need(3 * ni * n_sys, 8)));
            Span<int> a_node = h->arena.take<int>(nm), a_incl = h->arena.take<int>(nm);
            Span<double> a_d = h->arena.take<double>(nm), a_dir = h->arena.take<double>(2 * nm), a_c = h->arena.take<double>(2 * ni),
                         a_prm = h->arena.take<double>(3 * ni * n_sys);
            memcpy(a_node.p, morph->node, nm * sizeof(int)); memcpy(a_incl.p, morph->incl, nm * sizeof(int));
            memcpy(a_d.p, morph->d, nm * sizeof(double)); memcpy(a_dir.p, morph->dir, 2 * nm * sizeof(double));
            memcpy(a_c.p, morph->centre, 2 * ni * sizeof(double)); memcpy(a_prm.p, morph->prm, 3 * ni * n_sys * sizeof(double));
            UPL(h->m_node, a_node); UPL(h->m_incl, a_incl); UPL(h->m_d, a_d); UPL(h->m_dir, a_dir); UPL(h->m_centre, a_c); UPL(h->m_prm, a_prm);
            const unsigned char* src = (const unsigned char*)stash(h->d_xy.p, 2 * (size_t)nv * sizeof(double));
            CU(h->d_xy.alloc(cnt));
            const long long bytes = 2LL * nv * (long long)sizeof(double);
            k_rep_add<unsigned char><<<cdiv(bytes * n_sys, 256), 256, 0, st>>>((unsigned char*)h->d_xy.p, src, bytes, n_sys, 0, 0);
            MorphDev MD;
            MD.node = h->m_node.p; MD.incl = h->m_incl.p; MD.d = h->m_d.p; MD.dir = h->m_dir.p; MD.centre = h->m_centre.p; MD.prm = h->m_prm.p;
            MD.n_mov = morph->n_mov; MD.n_incl = morph->n_incl; MD.r0 = morph->r0; MD.R = morph->R;
            if (nm) k_morph<<<cdiv((long long)nm * n_sys, 256), 256, 0, st>>>(n_sys, nv, MD, h->d_xy.p);
        } else {
        CU(h->arena.reserve(PinnedArena::need(cnt, 8)));
        Span<double> a_xy = h->arena.take<double>(cnt);
        parallel_for(n_sys, [&](int s) { memcpy(a_xy.p + 2 * (size_t)s * nv, xy + 2 * (size_t)s * nv, 2 * sizeof(double) * (size_t)nv); });
        UPL(h->d_xy, a_xy);
        }
        CU(h->node_xy.alloc(2 * (size_t)h->N));
        k_shared_node_xy<<<cdiv(h->N, 256), 256, 0, st>>>(n_sys, (int)nn, nv, h->d_xy.p, h->node_va.p, h->node_vb.p, h->node_xy.p);
    }
    // ---- solver state and the coordinate-dependent symbolic pieces for all systems
    if (int rc = alloc_solver_state(h, n_sys, model, h->ncx, h->ncy)) return rc;
    CU(h->we_geo.alloc(3 * (size_t)std::max<long long>(h->n_we, 1)));
    const int n_win = h->n_win;
    const long long cn_cap = h->R + 5LL * n_win + 4096;
    CU(h->win_cn0.alloc(n_win)); CU(h->win_ncn.alloc(n_win)); CU(h->win_active.alloc((size_t)n_win + 4)); CU(h->cn_node.alloc((size_t)cn_cap));
    CU(h->cn_beg.alloc((size_t)cn_cap)); CU(h->cn_ent.alloc(4 * (size_t)h->R + 8 * (size_t)n_win + 64)); CU(h->rowinfo.alloc((size_t)h->R));
    CU(h->win_desc.alloc((size_t)n_win)); CU(h->win_scal.alloc(8 * (size_t)n_win)); CU(h->sys_blocks.alloc(n_sys));
    CU(cudaMemsetAsync(h->sym_cnt.p + 2, 0, 4 * sizeof(int), st));
    const size_t smc = 2 * (size_t)((h->LV.g[0].G + 31) / 32) * sizeof(unsigned);
    k_sym_coarse<<<n_win, RVE_WIN, smc, st>>>(win_tab(h), h->win_cn0.p, h->win_ncn.p, h->row_node.p, h->node_xy.p, h->box.p, h->LV.g[0],
                                              model, h->rowinfo.p, h->cn_node.p, h->cn_beg.p, h->cn_ent.p, (int)cn_cap, h->sym_cnt.p, h->win_desc.p);
    h->launches += 40;
    int scnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    CU(cudaMemcpyAsync(scnt, h->sym_cnt.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(stream_wait(h));
    CU(cudaGetLastError());
    if (scnt[2] == 4) FAIL(RVE_E_MESH, "device pattern build: coarse-node list overflow");
    h->max_cn = scnt[5];
    h->t_ms[0] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    h->t_ms[1] = 0.0;
    h->have_batch = true;
    return 0;
}

int rve_set_batch_shared(rve_handle_t h, int32_t n_sys, int32_t nv, int32_t nt, const double* xy, const int32_t* tri,
                         const int32_t* region, const double* mat, int32_t model, int32_t vref_nb, const double* vref_box) {
    return set_batch_shared_impl(h, n_sys, nv, nt, xy, tri, region, mat, model, vref_nb, vref_box, nullptr);
}

int rve_set_batch_morphed(rve_handle_t h, int32_t n_sys, int32_t nv, int32_t nt, const double* xy_template, const int32_t* tri,
                          const int32_t* region, const double* mat, int32_t model, int32_t vref_nb, const double* vref_box,
                          int32_t n_mov, const int32_t* mov_node, const int32_t* mov_incl, const double* mov_d, const double* mov_dir,
                          int32_t n_incl, const double* centre, double r0, double R, const double* prm) {
    if (!h) return RVE_E_ARG;
    if (n_mov < 0 || n_incl <= 0 || !centre || !prm || (n_mov > 0 && (!mov_node || !mov_incl || !mov_d || !mov_dir)) || !(r0 > 0) || !(R > r0))
        FAIL(RVE_E_ARG, "bad morphing description");
    for (long long i = 0; i < (long long)n_sys * n_incl; ++i) {
        const double l = prm[3 * i], e = prm[3 * i + 1];
        if (!(l > 0) || !(e > 0) || !(std::max(l, e * l) < R)) FAIL(RVE_E_MESH, "inclusion " + std::to_string(i % n_incl) + " of system " + std::to_string(i / n_incl) + ": radius outside (0, R)");
    }
    MorphHost M;
    M.n_mov = n_mov; M.n_incl = n_incl; M.node = mov_node; M.incl = mov_incl; M.d = mov_d; M.dir = mov_dir; M.centre = centre; M.prm = prm;
    M.r0 = r0; M.R = R;
    return set_batch_shared_impl(h, n_sys, nv, nt, xy_template, tri, region, mat, model, vref_nb, vref_box, &M);
}

static int ensure_emap(rve_handle_t h);

int rve_assemble(rve_handle_t h) {
    H_CHECK(h);
    if (!h->have_batch) FAIL(RVE_E_STATE, "rve_assemble before rve_set_batch");
    cudaStream_t st = h->stream;
    CU(cudaEventRecord(h->ev[0], st));
    if (h->opt_operator == RVE_OPERATOR_MATFREE) {
        // matrix-free operator: geometry x material of the window elements, block-Jacobi diagonal from the element fans
        k_we_geo<<<h->n_win, RVE_WIN, 0, st>>>(h->n_win, h->win_we0.p, h->win_nwe.p, h->we_elem.p, h->elem_node.p, h->elem_reg.p,
                                                h->node_xy.p, h->we_geo.p, h->mat);
        k_diag<<<h->n_win, RVE_WIN, 0, st>>>(win_tab(h), h->adjf_ptr.p, h->adjf.p, h->elem_node.p, h->elem_reg.p, h->node_xy.p,
                                              h->node_row.p, h->dinv.p, h->wmean.p, h->mat);
        h->have_bval = false;
    } else {
        if (int rc = ensure_emap(h)) return rc;
        CU(h->bval.alloc(128 * (size_t)h->NG));
        k_assemble<<<h->n_win, RVE_WIN, 0, st>>>(win_tab(h), sell_tab(h), h->adjf_ptr.p, h->adjf.p, h->elem_node.p, h->elem_reg.p,
                                                  h->node_xy.p, h->node_row.p, h->emap.p, h->bval.p, h->wmean.p, h->mat);
        k_dinv<<<h->n_win, RVE_WIN, 0, st>>>(win_tab(h), sell_tab(h), h->bval.p, h->dinv.p);
        h->have_bval = true;
    }
    CU(cudaEventRecord(h->ev[1], st));
    // two-level preconditioner: Galerkin hierarchy
    Levels& LV = h->LV;
    for (int l = 0; l < LV.L; ++l) CU(cudaMemsetAsync(LV.A[l], 0, h->lvA[l].n * sizeof(double), st));
#if RVE_GALERKIN_THREAD
    k_galerkin1_t<<<cdiv(h->E, RVE_GAL_THREADS), RVE_GAL_THREADS, 54 * RVE_GAL_THREADS * sizeof(double), st>>>((int)h->E, h->elem_node.p, h->elem_reg.p, h->elem_sys.p, h->node_xy.p,
                                                h->node_row.p, h->box.p, LV.g[0], LV.A[0], h->mat);
#else
    k_galerkin1<<<cdiv(h->E, 8), 256, 0, st>>>((int)h->E, h->elem_node.p, h->elem_reg.p, h->elem_sys.p, h->node_xy.p,
                                                h->node_row.p, h->box.p, LV.g[0], LV.A[0], h->mat);
#endif
    k_galerkin_mirror<<<cdiv((long long)h->n_sys * 12 * LV.g[0].G, 256), 256, 0, st>>>(h->n_sys, LV.g[0], LV.A[0]);
    for (int l = 0; l + 1 < LV.L; ++l) {
        long long n = (long long)h->n_sys * LV.g[l + 1].G * 25;
        k_galerkin_coarse<<<cdiv(n, 256), 256, 0, st>>>(h->n_sys, LV.g[l], LV.g[l + 1], LV.A[l], LV.A[l + 1]);
    }
    for (int l = 0; l < LV.L; ++l) {
        long long n = (long long)h->n_sys * LV.g[l].G;
        k_coarse_dinv<<<cdiv(n, 256), 256, 0, st>>>(h->n_sys, LV.g[l], LV.A[l], LV.dinv[l]);
        k_coarse_pack<<<cdiv(25 * n, 256), 256, 0, st>>>(h->n_sys, LV.g[l], LV.A[l], LV.dinv[l], LV.A4[l], LV.H4[l]);
    }
    h->launches += 4 + (LV.L - 1) + 2 * LV.L;
    CU(cudaEventRecord(h->ev[2], st));
    CU(stream_wait(h));
    CU(cudaGetLastError());
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]); h->t_ms[2] = ms;
    cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]); h->t_ms[3] = ms;
    h->assembled = true;
    return 0;
}

// SELL values of the assembled operator, built on demand when the batch was set up for the matrix-free one
// element -> slot map of the assembled operator, built on demand when the batch was set for the matrix-free one:
// k_sym_cols once more (same columns and lengths, block counts left alone)
static int ensure_emap(rve_handle_t h) {
    if (h->have_emap) return 0;
    cudaStream_t st = h->stream;
    CU(h->emap.alloc(36 * (size_t)h->E));
    k_sym_cols<<<h->n_win * RVE_SYMCOLS_SPLIT, RVE_WIN / RVE_SYMCOLS_SPLIT, 0, st>>>(win_tab(h), h->d_row_ptr.p, h->adjf_ptr.p, h->adjf.p,
                                                                                    h->elem_node.p, h->node_row.p, h->csr_col.p, h->rlen.p,
                                                                                    h->emap.p, nullptr, h->sym_cnt.p);
    h->launches += 1;
    h->have_emap = true;
    return 0;
}

static int ensure_bval(rve_handle_t h) {
    if (h->have_bval) return 0;
    if (int rc = ensure_emap(h)) return rc;
    cudaStream_t st = h->stream;
    CU(h->bval.alloc(128 * (size_t)h->NG));
    k_assemble<<<h->n_win, RVE_WIN, 0, st>>>(win_tab(h), sell_tab(h), h->adjf_ptr.p, h->adjf.p, h->elem_node.p, h->elem_reg.p,
                                              h->node_xy.p, h->node_row.p, h->emap.p, h->bval.p, h->wmean.p, h->mat);
    h->launches += 1;
    CU(stream_wait(h));
    CU(cudaGetLastError());
    h->have_bval = true;
    return 0;
}

int rve_get_sizes(rve_handle_t h, int32_t sys, int64_t* n_nodes, int64_t* n_rows, int64_t* n_blocks, int64_t* n_elems) {
    H_CHECK(h);
    if (!h->have_batch || sys < 0 || sys >= h->n_sys) FAIL(RVE_E_ARG, "bad system index");
    if (n_nodes) *n_nodes = h->info[sys].nn;
    if (n_rows) *n_rows = h->info[sys].na;
    if (n_blocks) *n_blocks = (int64_t)h->true_blocks[sys];
    if (n_elems) *n_elems = h->info[sys].nt;
    return 0;
}

int rve_get_csr(rve_handle_t h, int32_t sys, int32_t* rowptr, int32_t* col, double* val) {
    H_CHECK(h);
    if (!h->assembled || sys < 0 || sys >= h->n_sys) FAIL(RVE_E_STATE, "not assembled / bad system index");
    if (int rc = ensure_bval(h)) return rc;
    const SysInfo& Y = h->info[sys];
    const size_t ng = (size_t)Y.ngrp;
    const int rb0 = h->row_base[sys], sl0 = h->slice_base[sys];
    std::vector<double> bv(128 * ng);
    CU(cudaMemcpy(bv.data(), h->bval.p + 128 * (size_t)h->grp_base[sys], bv.size() * sizeof(double), cudaMemcpyDeviceToHost));
    std::vector<int> cc((size_t)Y.nblk);       // block columns as found on the device (global rows), capacity layout
    CU(cudaMemcpy(cc.data(), h->csr_col.p + (size_t)h->blk_base[sys], cc.size() * sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<unsigned char> rl(Y.na);       // true row lengths
    CU(cudaMemcpy(rl.data(), h->rlen.p + (size_t)rb0, rl.size(), cudaMemcpyDeviceToHost));
    std::vector<unsigned> rp((size_t)Y.na + 1), sp((size_t)Y.nsl + 1);   // capacity row pointer / slice pointer (global offsets)
    CU(cudaMemcpy(rp.data(), h->d_row_ptr.p + (size_t)rb0, rp.size() * sizeof(unsigned), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(sp.data(), h->slice_ptr.p + (size_t)sl0, sp.size() * sizeof(unsigned), cudaMemcpyDeviceToHost));
    const unsigned kb0 = (unsigned)h->blk_base[sys], gb0 = (unsigned)h->grp_base[sys];
    // SELL slot j of a row holds the j-th column the device found for that row
    int64_t pos = 0;
    for (int r = 0; r < Y.na; ++r) {
        const int w = r / RVE_WIN, rlw = r % RVE_WIN, lane = rlw & 31;
        const size_t g0 = sp[(size_t)w * (RVE_WIN / 32) + (rlw >> 5)] - gb0;
        const unsigned k0 = rp[r] - kb0;
        for (int c = 0; c < 2; ++c) {
            rowptr[2 * r + c] = (int32_t)pos;
            for (unsigned k = k0; k < k0 + rl[r]; ++k) {
                const size_t g = g0 + (k - k0);
                for (int d = 0; d < 2; ++d) { col[pos] = 2 * (cc[k] - rb0) + d; val[pos] = bv[128 * g + 64 * c + 2 * lane + d]; ++pos; }
            }
        }
    }
    rowptr[2 * Y.na] = (int32_t)pos;
    return 0;
}

// (va, vb) = the mesh vertices a node sits on (va == vb) or between: numbering-independent node keys
static int node_keys(rve_handle_t h, int32_t sys, std::vector<int>& va, std::vector<int>& vb) {
    const SysInfo& Y = h->info[sys];
    va.resize(Y.nn); vb.resize(Y.nn);
    CU(cudaMemcpy(va.data(), h->node_va.p + (size_t)h->node_base[sys], va.size() * sizeof(int), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(vb.data(), h->node_vb.p + (size_t)h->node_base[sys], vb.size() * sizeof(int), cudaMemcpyDeviceToHost));
    return 0;
}

int rve_get_rowmap(rve_handle_t h, int32_t sys, int32_t* va, int32_t* vb) {
    H_CHECK(h);
    if (!h->have_batch || sys < 0 || sys >= h->n_sys) FAIL(RVE_E_ARG, "bad system index");
    const SysInfo& Y = h->info[sys];
    std::vector<int> a, b, rn(Y.na);
    if (int rc = node_keys(h, sys, a, b)) return rc;
    CU(cudaMemcpy(rn.data(), h->row_node.p + (size_t)h->row_base[sys], rn.size() * sizeof(int), cudaMemcpyDeviceToHost));
    for (int r = 0; r < Y.na; ++r) { va[r] = a[rn[r] - h->node_base[sys]]; vb[r] = b[rn[r] - h->node_base[sys]]; }
    return 0;
}

int rve_get_nodemap(rve_handle_t h, int32_t sys, int32_t* va, int32_t* vb) {
    H_CHECK(h);
    if (!h->have_batch || sys < 0 || sys >= h->n_sys) FAIL(RVE_E_ARG, "bad system index");
    std::vector<int> a, b;
    if (int rc = node_keys(h, sys, a, b)) return rc;
    for (size_t n = 0; n < a.size(); ++n) { va[n] = a[n]; vb[n] = b[n]; }
    return 0;
}

int rve_get_rhs(rve_handle_t h, int32_t sys, double* rhs) {
    H_CHECK(h);
    if (!h->solved || sys < 0 || sys >= h->n_sys) FAIL(RVE_E_STATE, "no solve yet / bad system index");
    CU(cudaMemcpy(rhs, h->b.p + (size_t)h->row_base[sys] * h->K * 2, (size_t)h->info[sys].na * h->K * 2 * sizeof(double),
                  cudaMemcpyDeviceToHost));
    return 0;
}

int rve_get_solution(rve_handle_t h, int32_t sys, double* u) {
    H_CHECK(h);
    if (!h->solved || sys < 0 || sys >= h->n_sys) FAIL(RVE_E_STATE, "no solve yet / bad system index");
    CU(cudaMemcpy(u, h->U.p + (size_t)h->node_base[sys] * h->NL * 2, (size_t)h->info[sys].nn * h->NL * 2 * sizeof(double),
                  cudaMemcpyDeviceToHost));
    return 0;
}

int rve_get_coarse_dim(rve_handle_t h, int32_t level, int32_t* G, int32_t* n_levels) {
    H_CHECK(h);
    if (!h->have_batch) FAIL(RVE_E_STATE, "no batch");
    if (n_levels) *n_levels = h->LV.L;
    if (G) { if (level < 0 || level >= h->LV.L) FAIL(RVE_E_ARG, "bad level"); *G = h->LV.g[level].G; }
    return 0;
}

int rve_get_coarse_dense(rve_handle_t h, int32_t sys, int32_t level, double* dense) {
    H_CHECK(h);
    if (!h->assembled || sys < 0 || sys >= h->n_sys || level < 0 || level >= h->LV.L) FAIL(RVE_E_STATE, "bad state/arguments");
    const Grid g = h->LV.g[level];
    std::vector<double> A(100 * (size_t)g.G);
    CU(cudaMemcpy(A.data(), h->LV.A[level] + (size_t)sys * 100 * g.G, A.size() * sizeof(double), cudaMemcpyDeviceToHost));
    const size_t n = 2 * (size_t)g.G;
    std::fill(dense, dense + n * n, 0.0);
    for (int j = 0; j < g.ny; ++j)
        for (int i = 0; i < g.nx; ++i) {
            if (i < g.lox || i > g.hix || j < g.loy || j > g.hiy) continue;
            const int I = j * g.nx + i;
            for (int dj = -2; dj <= 2; ++dj)
                for (int di = -2; di <= 2; ++di) {
                    int ii = i + di, jj = j + dj;
                    if (g.wrap) { ii = ((ii % g.ncx) + g.ncx) % g.ncx; jj = ((jj % g.ncy) + g.ncy) % g.ncy; }
                    else if (ii < g.lox || ii > g.hix || jj < g.loy || jj > g.hiy) continue;
                    const int J = jj * g.nx + ii;
                    const size_t slot = (size_t)(dj + 2) * 5 + (di + 2);
                    for (int c = 0; c < 2; ++c)
                        for (int d = 0; d < 2; ++d)
                            dense[(2 * (size_t)I + c) * n + 2 * J + d] += A[(slot * 4 + 2 * c + d) * g.G + I];
                }
        }
    return 0;
}

int rve_solve(rve_handle_t h, int32_t n_rhs, const double* eps, const double* uD, double rtol, int32_t maxit,
              int32_t precond, int32_t* status, int32_t* iters) {
    H_CHECK(h);
    if (!h->assembled) FAIL(RVE_E_STATE, "rve_solve before rve_assemble");
    if (n_rhs < 1 || n_rhs > 3 || !eps || maxit < 1 || rtol <= 0) FAIL(RVE_E_ARG, "bad solve arguments");
    if (uD && h->model != RVE_MODEL_DNN) FAIL(RVE_E_ARG, "uD given but the batch model is not 'dnn'");
    if (!uD && h->model == RVE_MODEL_DNN) FAIL(RVE_E_ARG, "'dnn' needs uD (use 'lin' for zero Dirichlet data)");
    const int NL = n_rhs;
    const int K = (h->model == RVE_MODEL_MR) ? 3 : n_rhs;
    const int n_sys = h->n_sys;
    if (h->opt_operator == RVE_OPERATOR_ASSEMBLED) { if (int rc = ensure_bval(h)) return rc; }
    h->K = K; h->NL = NL; h->solved = false; h->have_samples = false;
    cudaStream_t st = h->stream;
    const size_t RK = (size_t)h->R * K * 2;
    CU(h->p.alloc(RK)); CU(h->q.alloc(RK)); CU(h->x.alloc(RK)); CU(h->r.alloc(RK)); CU(h->b.alloc(RK));
    CU(h->eps.alloc((size_t)n_sys * NL * 3)); CU(h->eps_eff.alloc((size_t)n_sys * NL * 3)); CU(h->Bmat.alloc((size_t)n_sys * NL * 4));
    CU(h->U.alloc((size_t)h->N * NL * 2));
    CU(cudaMemcpyAsync(h->eps.p, eps, (size_t)n_sys * NL * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
    h->h2d_bytes += (double)n_sys * NL * 3 * sizeof(double) + (uD ? (double)n_sys * NL * 2 * h->nvref * sizeof(double) : 0.0);
    h->d2h_bytes += 2.0 * n_sys * sizeof(int);
    const int nside = h->vref_nb > 0 ? h->vref_nb - 1 : 0;
    if (uD) {
        CU(h->uD.alloc((size_t)n_sys * NL * 2 * h->nvref));
        CU(cudaMemcpyAsync(h->uD.p, uD, h->uD.n * sizeof(double), cudaMemcpyHostToDevice, st));
        CU(h->gval.alloc((size_t)h->NBN * NL * 2));
    }
    CU(cudaEventRecord(h->ev[0], st));
    k_dnn_prep<<<cdiv((long long)n_sys * NL, 128), 128, 0, st>>>(n_sys, NL, nside, h->d_vbox.p, uD ? h->uD.p : nullptr,
                                                                  h->eps.p, h->eps_eff.p, h->Bmat.p, h->opt_sym_B);
    if (uD)
        k_bnd_values<<<cdiv(h->NBN * NL, 256), 256, 0, st>>>((int)h->NBN, NL, nside, h->bnd_sys.p, h->bnd_node.p, h->bnd_s.p,
                                                              h->node_xy.p, h->uD.p, h->Bmat.p, h->gval.p);
    CU(cudaMemsetAsync(h->sc.p, 0, h->sc.n * sizeof(double), st));
    if (h->model == RVE_MODEL_MR) {
        // (each boundary node receives at most two edge terms: the sum is order-independent)
        CU(cudaMemsetAsync(h->b.p, 0, RK * sizeof(double), st));
        k_rhs_mr<<<cdiv(h->NBE, 256), 256, 0, st>>>((int)h->NBE, h->bedge.p, h->bedge_nl.p, h->node_row.p, h->b.p);
    } else {
        k_rhs_rows<<<h->n_win, RVE_WIN, 0, st>>>(win_tab(h), K, h->adjf_ptr.p, h->adjf.p, h->elem_node.p, h->elem_reg.p, h->node_xy.p,
                                                  h->node_row.p, h->node_bnd.p, h->eps_eff.p, uD ? h->gval.p : nullptr, h->b.p, h->part.p, h->mat);
        k_fin_bref<<<cdiv(n_sys, 8), 256, 0, st>>>(n_sys, K, h->sys_win0.p, h->part.p, h->sc.p);
    }
    CU(cudaEventRecord(h->ev[1], st));
    // ---- PCG
    CU(cudaMemsetAsync(h->p.p, 0, h->p.n * sizeof(double), st));
    CU(cudaMemsetAsync(h->q.p, 0, RK * sizeof(double), st));
    CU(cudaMemsetAsync(h->x.p, 0, RK * sizeof(double), st));
    CU(cudaMemcpyAsync(h->r.p, h->b.p, RK * sizeof(double), cudaMemcpyDeviceToDevice, st));
    CU(cudaMemsetAsync(h->cnt.p, 0, n_sys * sizeof(int), st));
    CU(cudaMemsetAsync(h->d_iters.p, 0, n_sys * sizeof(int), st));
    std::vector<int> ones(n_sys, 1);
    CU(cudaMemcpyAsync(h->active.p, ones.data(), n_sys * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(h->win_active.p, 1, (size_t)h->n_win, st));
    CU(cudaMemcpyAsync(h->n_active.p, &n_sys, sizeof(int), cudaMemcpyHostToDevice, st));
    PcgArgs pa;
    pa.wt = win_tab(h); pa.sl = sell_tab(h); pa.LV = h->LV; pa.LV.K = K;
    pa.n_win = h->n_win; pa.n_sys = n_sys; pa.K = K; pa.precond = precond; pa.rtol2 = rtol * rtol;
    pa.sm_spmv = smem_spmv(h, K); pa.sm_upd = smem_update(K); pa.sm_dir = smem_direction(h, K);
    pa.sm_apply = smem_apply(h, K); pa.op = h->opt_operator;
    const Levels& LV = pa.LV;
    CU(cudaMemsetAsync(LV.rc[0], 0, h->lvR[0].n * sizeof(float), st));
#define UPD0_(KK) launch_update<KK>(h, pa, 1, -1, st)
    K_DISPATCH(K, UPD0_);
#undef UPD0_
    int it = 0, launched = 0;
    int CHECK = 8;                     // iterations between two reads of the active-systems counter; 2 once systems start to converge
    h->launches += 3 + (uD ? 1 : 0);
    auto pool_event = [&](size_t i) -> cudaEvent_t {
        while (h->evpool.size() <= i) { cudaEvent_t e; cudaEventCreate(&e); h->evpool.push_back(e); }
        return h->evpool[i];
    };
    *h->h_nactive = n_sys;
    // RVE_KTIME=1: per-iteration time of the V-cycle, direction, SpMV and update kernels (stderr; diagnostics only)
    static const bool ktime = [] { const char* e = std::getenv("RVE_KTIME"); return e && e[0] == '1'; }();
    auto dbg_event = [&](size_t i) -> cudaEvent_t {
        while (h->evdbg.size() <= i) { cudaEvent_t e; cudaEventCreate(&e); h->evdbg.push_back(e); }
        return h->evdbg[i];
    };
    while (it < maxit) {
        for (int k = 0; k < CHECK && it < maxit; ++k, ++it) {
            if (ktime) cudaEventRecord(dbg_event(3 * (size_t)it), st);
            if (precond) {
#define VC_(KK) launch_vcycle<KK>(h, pa, it == 0, st)
                K_DISPATCH(K, VC_);
#undef VC_
            }
            if (ktime) cudaEventRecord(dbg_event(3 * (size_t)it + 1), st);
#define DIR_(KK) launch_direction<KK>(h, pa, st)
            K_DISPATCH(K, DIR_);
#undef DIR_
            cudaEventRecord(pool_event(2 * (size_t)it), st);
#define SPMV_(KK) launch_spmv<KK>(h, pa, h->p.p, h->q.p, h->active.p, h->part.p, h->sc.p, st)
            K_DISPATCH(K, SPMV_);
#undef SPMV_
            cudaEventRecord(pool_event(2 * (size_t)it + 1), st);
#define UPD_(KK) launch_update<KK>(h, pa, 0, it, st)
            K_DISPATCH(K, UPD_);
#undef UPD_
            if (h->opt_resid_replace > 0 && (it + 1) % h->opt_resid_replace == 0) {
                // residual replacement: the recurrence residual is replaced by the true one, r = b - A x, and its norms and
                // restriction are recomputed (the alpha k_fin_alpha derives from this application is never used)
#define RSP_(KK) launch_spmv<KK>(h, pa, h->x.p, h->q.p, h->active.p, h->part.p, h->sc.p, st)
                K_DISPATCH(K, RSP_);
#undef RSP_
                k_true_residual<<<pa.n_win, 256, 0, st>>>(pa.wt, K, (const double2*)h->b.p, (const double2*)h->q.p, (double2*)h->r.p);
                if (precond) CU(cudaMemsetAsync(LV.rc[0], 0, h->lvR[0].n * sizeof(float), st));
#define UPD2_(KK) launch_update<KK>(h, pa, 2, it, st)
                K_DISPATCH(K, UPD2_);
#undef UPD2_
                h->launches += 4;
            }
            if (ktime) cudaEventRecord(dbg_event(3 * (size_t)it + 2), st);
            ++launched;
            h->launches += 5;
        }
        CU(cudaMemcpyAsync(h->h_nactive, h->n_active.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(stream_wait(h));
        if (*h->h_nactive <= 0) break;
        if (*h->h_nactive < n_sys) CHECK = 2;      // the tail: stop within two iterations of the last system
    }
    CU(cudaEventRecord(h->ev[2], st));
    // ---- constraint post-processing + nodal field
    MixPtrs M; M.Mx = h->Mx.p; M.Gaff = h->Gaff.p; M.aaff = h->aaff.p;
    k_mix_coeffs<<<n_sys, 256, 0, st>>>(h->model, K, NL, h->x.p, h->wmean.p, h->d_row_base.p, h->d_bedge_base.p, h->bedge.p,
                                        h->bedge_nl.p, h->node_row.p, h->box.p, h->eps_eff.p, M);
    k_nodal_field<<<cdiv(h->N * NL, 256), 256, 0, st>>>(h->N, K, NL, h->node_sys.p, h->node_row.p, h->node_bnd.p, h->node_xy.p,
                                                        h->x.p, uD ? h->gval.p : nullptr, M, h->model != RVE_MODEL_MR, h->U.p);
    CU(cudaEventRecord(h->ev[3], st));
    std::vector<int> hit(n_sys), hact(n_sys);
    CU(cudaMemcpyAsync(hit.data(), h->d_iters.p, n_sys * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(hact.data(), h->active.p, n_sys * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(stream_wait(h));
    CU(cudaGetLastError());
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]); h->t_ms[4] = ms;
    cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]); h->t_ms[5] = ms;
    double spmv_ms = 0;
    for (int i = 0; i < launched; ++i) { cudaEventElapsedTime(&ms, h->evpool[2 * i], h->evpool[2 * i + 1]); spmv_ms += ms; }
    h->t_ms[6] = spmv_ms;
    if (ktime && launched > 0) {
        double tv = 0, td = 0, tu = 0;
        for (int i = 0; i < launched; ++i) {
            cudaEventElapsedTime(&ms, h->evdbg[3 * i], h->evdbg[3 * i + 1]); tv += ms;
            cudaEventElapsedTime(&ms, h->evdbg[3 * i + 1], h->evpool[2 * i]); td += ms;
            cudaEventElapsedTime(&ms, h->evpool[2 * i + 1], h->evdbg[3 * i + 2]); tu += ms;
        }
        if (const char* e = std::getenv("RVE_KTIME_SERIES")) {      // time of every launched iteration (us): tail cost of the converged systems
            if (e[0] == '1') {
                fprintf(stderr, "[rve_solve] iteration series (us):");
                for (int i = 0; i < launched; ++i) {
                    cudaEventElapsedTime(&ms, h->evdbg[3 * i], h->evdbg[3 * i + 2]);
                    fprintf(stderr, " %.0f", 1e3 * ms);
                }
                fprintf(stderr, "\n");
            }
        }
        fprintf(stderr, "[rve_solve] n_sys %d K %d its %d: per iteration (us) vcycle %.1f direction %.1f spmv %.1f update %.1f\n", n_sys, K,
                launched, 1e3 * tv / launched, 1e3 * td / launched, 1e3 * spmv_ms / launched, 1e3 * tu / launched);
    }
    cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]); h->c_cnt[5] = ms;  // constraint post-processing + nodal field
    h->launches += 2;
    bool allconv = true;
    double sumit = 0;
    for (int s = 0; s < n_sys; ++s) {
        const int stt = hact[s] ? 1 : 0;
        if (stt) { allconv = false; hit[s] = launched; }
        if (status) status[s] = stt;
        if (iters) iters[s] = hit[s];
        sumit += hit[s];
    }
    h->c_cnt[0] = launched; h->c_cnt[2] = launched; h->c_cnt[3] = sumit;
    // algorithmic SpMV bytes of the solve: sum over systems of iters * B_spmv(K)
    double bytes = 0;
    for (int s = 0; s < n_sys; ++s) {
        const double n = 2.0 * h->info[s].na, nnz = 4.0 * (double)h->true_blocks[s];
        bytes += hit[s] * (12.0 * nnz + 4.0 * (n + 1) + 16.0 * n * K);
    }
    h->c_cnt[1] = bytes;
    h->solved = true;
    if (!allconv) FAIL(RVE_E_NOCONV, "PCG did not converge for at least one system");
    return 0;
}

static int run_epilogue(rve_handle_t h) {
    cudaStream_t st = h->stream;
    const int NL = h->NL, n_sys = h->n_sys;
    CU(h->sigL.alloc((size_t)n_sys * NL * 3)); CU(h->sigT.alloc((size_t)n_sys * NL * 3));
    CU(h->aout.alloc((size_t)n_sys * NL * 2)); CU(h->Bout.alloc((size_t)n_sys * NL * 4)); CU(h->gradT.alloc((size_t)n_sys * NL * 4));
    CU(cudaEventRecord(h->ev[0], st));
    CU(cudaMemsetAsync(h->acc.p, 0, h->acc.n * sizeof(double), st));
    k_epi_elements<<<cdiv(h->E, 128), 128, 0, st>>>((int)h->E, NL, h->elem_node.p, h->elem_reg.p, h->elem_sys.p, h->node_xy.p,
                                                    h->U.p, h->acc.p, h->mat);
    k_epi_finalize<<<cdiv((long long)n_sys * NL, 128), 128, 0, st>>>(n_sys, NL, h->acc.p, h->eps_eff.p, h->mat, h->sigL.p,
                                                                     h->sigT.p, h->aout.p, h->Bout.p, h->gradT.p);
    if (h->nvref) {
        CU(h->sol.alloc((size_t)n_sys * NL * 2 * h->nvref));
        const long long n = (long long)n_sys * h->nvref * NL;
        k_epi_sample<<<cdiv(n, 256), 256, 0, st>>>(n, h->nvref, NL, h->samp_elem.p, h->samp_lam.p, h->elem_node.p, h->U.p, h->sol.p);
    }
    h->launches += 2 + (h->nvref ? 1 : 0);
    h->have_samples = h->nvref > 0;
    CU(cudaEventRecord(h->ev[1], st));
    return 0;
}

int rve_epilogue_tangent(rve_handle_t h, double* C) {
    H_CHECK(h);
    if (!h->solved) FAIL(RVE_E_STATE, "rve_epilogue_tangent before rve_solve");
    if (h->NL != 3 || !C) FAIL(RVE_E_ARG, "tangent needs the 3 unit load cases");
    int rc = run_epilogue(h);
    if (rc) return rc;
    std::vector<double> s((size_t)h->n_sys * 9);
    CU(cudaMemcpyAsync(s.data(), h->sigL.p, s.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    h->d2h_bytes += (double)s.size() * sizeof(double);
    CU(stream_wait(h));
    CU(cudaGetLastError());
    float ms = 0; cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]); h->t_ms[7] = ms;
    for (int sys = 0; sys < h->n_sys; ++sys)
        for (int l = 0; l < 3; ++l)
            for (int k = 0; k < 3; ++k) C[(size_t)sys * 9 + 3 * k + l] = s[((size_t)sys * 3 + l) * 3 + k];
    return 0;
}

int rve_epilogue_snapshot(rve_handle_t h, double* sol, double* sigma, double* a, double* B, double* sigmaT) {
    H_CHECK(h);
    if (!h->solved) FAIL(RVE_E_STATE, "rve_epilogue_snapshot before rve_solve");
    if (sol && !h->nvref) FAIL(RVE_E_ARG, "no Vref was given to rve_set_batch");
    int rc = run_epilogue(h);
    if (rc) return rc;
    cudaStream_t st = h->stream;
    const size_t n = (size_t)h->n_sys * h->NL;
    h->d2h_bytes += (double)n * sizeof(double) * ((sol ? 2 * h->nvref : 0) + (sigma ? 3 : 0) + (sigmaT ? 3 : 0) + (a ? 2 : 0) + (B ? 4 : 0));
    if (sol) CU(cudaMemcpyAsync(sol, h->sol.p, n * 2 * h->nvref * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (sigma) CU(cudaMemcpyAsync(sigma, h->sigL.p, n * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (sigmaT) CU(cudaMemcpyAsync(sigmaT, h->sigT.p, n * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (a) CU(cudaMemcpyAsync(a, h->aout.p, n * 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (B) CU(cudaMemcpyAsync(B, h->Bout.p, n * 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(stream_wait(h));
    CU(cudaGetLastError());
    float ms = 0; cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]); h->t_ms[7] = ms;
    return 0;
}

int rve_set_option(rve_handle_t h, int32_t option, int32_t value) {
    H_CHECK(h);
    if (option == RVE_OPT_SYM_B) { h->opt_sym_B = value ? 1 : 0; return 0; }
    if (option == RVE_OPT_RESIDUAL_REPLACE) {
        if (value < 0) FAIL(RVE_E_ARG, "residual replacement interval must be >= 0");
        h->opt_resid_replace = value;
        return 0;
    }
    if (option == RVE_OPT_OPERATOR) {
        if (value != RVE_OPERATOR_ASSEMBLED && value != RVE_OPERATOR_MATFREE) FAIL(RVE_E_ARG, "unknown operator");
        if (value == RVE_OPERATOR_MATFREE && h->assembled && h->opt_operator != RVE_OPERATOR_MATFREE) h->assembled = false;   // needs k_we_geo
        h->opt_operator = value;
        return 0;
    }
    FAIL(RVE_E_ARG, "unknown option");
}

int rve_epilogue_homogenisation(rve_handle_t h, double* sigmaL, double* sigmaT, double* gradL, double* gradT, double* eps_eff) {
    H_CHECK(h);
    if (!h->solved) FAIL(RVE_E_STATE, "rve_epilogue_homogenisation before rve_solve");
    int rc = run_epilogue(h);
    if (rc) return rc;
    cudaStream_t st = h->stream;
    const size_t n = (size_t)h->n_sys * h->NL;
    std::vector<double> bl(gradL ? 4 * n : 0);
    if (sigmaL) CU(cudaMemcpyAsync(sigmaL, h->sigL.p, n * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (sigmaT) CU(cudaMemcpyAsync(sigmaT, h->sigT.p, n * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (gradL) CU(cudaMemcpyAsync(bl.data(), h->Bout.p, n * 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (gradT) CU(cudaMemcpyAsync(gradT, h->gradT.p, n * 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (eps_eff) CU(cudaMemcpyAsync(eps_eff, h->eps_eff.p, n * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    h->d2h_bytes += (double)n * sizeof(double) * ((sigmaL ? 3 : 0) + (sigmaT ? 3 : 0) + (gradL ? 4 : 0) + (gradT ? 4 : 0) + (eps_eff ? 3 : 0));
    CU(stream_wait(h));
    CU(cudaGetLastError());
    if (gradL) for (size_t i = 0; i < 4 * n; ++i) gradL[i] = -bl[i];      // B = -average gradient
    float ms = 0; cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]); h->t_ms[7] = ms;
    return 0;
}

int rve_project(rve_handle_t h, int32_t nmax, int32_t nh, const double* W, const double* M, const double* translate, double* Y) {
    H_CHECK(h);
    if (!h->solved || !h->have_samples || !h->nvref) FAIL(RVE_E_STATE, "rve_project needs rve_epilogue_snapshot of the current solve first");
    if (nh != 2 * h->nvref || nmax < 1 || !W || !M || !Y) FAIL(RVE_E_ARG, "bad projection arguments");
    const int NL = h->NL;
    // MWt[l][q][j] = sum_k M[q][k] W[l][j][k]
    std::vector<double> mwt((size_t)NL * nh * nmax);
    for (int l = 0; l < NL; ++l)
        for (int q = 0; q < nh; ++q)
            for (int j = 0; j < nmax; ++j) {
                double acc = 0;
                const double* Mq = M + (size_t)q * nh;
                const double* Wj = W + ((size_t)l * nmax + j) * nh;
                for (int k = 0; k < nh; ++k) acc += Mq[k] * Wj[k];
                mwt[((size_t)l * nh + q) * nmax + j] = acc;
            }
    cudaStream_t st = h->stream;
    UPL(h->MWt, mwt);
    const size_t n = (size_t)h->n_sys * NL;
    if (translate) {
        CU(h->transl.alloc(n * 2));
        CU(cudaMemcpyAsync(h->transl.p, translate, n * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    CU(h->Yd.alloc(n * nmax));
    h->launches += 1;
    k_project<<<(unsigned)n, 128, nh * sizeof(double), st>>>(NL, nh, nmax, h->sol.p, translate ? h->transl.p : nullptr, h->MWt.p, h->Yd.p);
    CU(cudaMemcpyAsync(Y, h->Yd.p, n * nmax * sizeof(double), cudaMemcpyDeviceToHost, st));
    h->d2h_bytes += (double)n * nmax * sizeof(double);
    h->h2d_bytes += (translate ? (double)n * 2 * sizeof(double) : 0.0);
    CU(stream_wait(h));
    CU(cudaGetLastError());
    return 0;
}

int rve_timings(rve_handle_t h, double* t, double* c) {
    H_CHECK(h);
    if (t) for (int k = 0; k < 8; ++k) t[k] = h->t_ms[k];
    h->c_cnt[4] = h->launches; h->c_cnt[6] = h->h2d_bytes; h->c_cnt[7] = h->d2h_bytes;
    if (c) for (int k = 0; k < 8; ++k) c[k] = h->c_cnt[k];
    return 0;
}

int rve_bench_spmv(rve_handle_t h, int32_t n_rhs, int32_t reps, double* ms_per_launch, double* alg_bytes_per_launch) {
    H_CHECK(h);
    if (!h->assembled) FAIL(RVE_E_STATE, "rve_bench_spmv before rve_assemble");
    if (n_rhs < 1 || n_rhs > 3 || reps < 1) FAIL(RVE_E_ARG, "bad arguments");
    if (h->opt_operator == RVE_OPERATOR_ASSEMBLED) { if (int rc = ensure_bval(h)) return rc; }
    cudaStream_t st = h->stream;
    const int K = n_rhs, n_sys = h->n_sys;
    DBuf<double> p, q, part, sc;
    DBuf<int> active, cnt;
    CU(p.alloc((size_t)h->R * K * 2)); CU(q.alloc((size_t)h->R * K * 2)); CU(part.alloc((size_t)h->n_win * 8));
    CU(sc.alloc((size_t)n_sys * 4 * SC_N)); CU(active.alloc(n_sys)); CU(cnt.alloc(n_sys));
    std::vector<double> hp((size_t)h->R * K * 2);
    for (size_t i = 0; i < hp.size(); ++i) hp[i] = 1.0 + 1e-3 * (double)(i % 97);
    CU(cudaMemcpyAsync(p.p, hp.data(), hp.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    std::vector<int> ones(n_sys, 1);
    CU(cudaMemcpyAsync(active.p, ones.data(), n_sys * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(h->win_active.p, 1, (size_t)h->n_win, st));   // every window on (a finished solve leaves them off)
    CU(cudaMemsetAsync(cnt.p, 0, n_sys * sizeof(int), st));
    CU(cudaMemsetAsync(sc.p, 0, sc.n * sizeof(double), st));
    PcgArgs pa;
    pa.wt = win_tab(h); pa.sl = sell_tab(h); pa.LV = h->LV; pa.n_win = h->n_win; pa.n_sys = n_sys; pa.K = K;
    pa.precond = 0; pa.rtol2 = 0; pa.sm_spmv = smem_spmv(h, K); pa.sm_upd = smem_update(K); pa.sm_dir = smem_direction(h, K);
    pa.sm_apply = smem_apply(h, K); pa.op = h->opt_operator;
#define SPMV_(KK) launch_spmv<KK>(h, pa, p.p, q.p, active.p, part.p, sc.p, st)
    for (int wu = 0; wu < 3; ++wu) K_DISPATCH(K, SPMV_);
    CU(cudaEventRecord(h->ev[0], st));
    for (int i = 0; i < reps; ++i) K_DISPATCH(K, SPMV_);
#undef SPMV_
    CU(cudaEventRecord(h->ev[1], st));
    CU(stream_wait(h));
    CU(cudaGetLastError());
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
    if (ms_per_launch) *ms_per_launch = ms / reps;
    if (alg_bytes_per_launch) {
        double blocks = 0;
        for (long long tb : h->true_blocks) blocks += (double)tb;          // true blocks (NB counts row capacities)
        const double n = 2.0 * (double)h->R, nnz = 4.0 * blocks;
        *alg_bytes_per_launch = 12.0 * nnz + 4.0 * (n + n_sys) + 16.0 * n * K;
    }
    return 0;
}

// parity-test accessor: q = A p with the operator RVE_OPT_OPERATOR selects, whole batch
int rve_apply_operator(rve_handle_t h, int32_t n_rhs, const double* p_host, double* q_host) {
    H_CHECK(h);
    if (!h->assembled) FAIL(RVE_E_STATE, "rve_apply_operator before rve_assemble");
    if (n_rhs < 1 || n_rhs > 3 || !p_host || !q_host) FAIL(RVE_E_ARG, "bad arguments");
    if (h->opt_operator == RVE_OPERATOR_ASSEMBLED) { if (int rc = ensure_bval(h)) return rc; }
    cudaStream_t st = h->stream;
    const int K = n_rhs, n_sys = h->n_sys;
    const size_t RK = (size_t)h->R * K * 2;
    DBuf<double> p, q, part, sc;
    DBuf<int> active, cnt;
    CU(p.alloc(RK)); CU(q.alloc(RK)); CU(part.alloc((size_t)h->n_win * 8));
    CU(sc.alloc((size_t)n_sys * 4 * SC_N)); CU(active.alloc(n_sys)); CU(cnt.alloc(n_sys));
    CU(cudaMemcpyAsync(p.p, p_host, RK * sizeof(double), cudaMemcpyHostToDevice, st));
    std::vector<int> ones(n_sys, 1);
    CU(cudaMemcpyAsync(active.p, ones.data(), n_sys * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(h->win_active.p, 1, (size_t)h->n_win, st));
    CU(cudaMemsetAsync(cnt.p, 0, n_sys * sizeof(int), st));
    CU(cudaMemsetAsync(sc.p, 0, sc.n * sizeof(double), st));
    CU(cudaMemsetAsync(q.p, 0, RK * sizeof(double), st));
    PcgArgs pa;
    pa.wt = win_tab(h); pa.sl = sell_tab(h); pa.LV = h->LV; pa.n_win = h->n_win; pa.n_sys = n_sys; pa.K = K;
    pa.precond = 0; pa.rtol2 = 0; pa.sm_spmv = smem_spmv(h, K); pa.sm_upd = smem_update(K); pa.sm_dir = smem_direction(h, K);
    pa.sm_apply = smem_apply(h, K); pa.op = h->opt_operator;
#define SPMV_(KK) launch_spmv<KK>(h, pa, p.p, q.p, active.p, part.p, sc.p, st)
    K_DISPATCH(K, SPMV_);
#undef SPMV_
    CU(cudaMemcpyAsync(q_host, q.p, RK * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(stream_wait(h));
    CU(cudaGetLastError());
    return 0;
}

}  // extern "C"

// Host-only test hook (no CUDA call): symbolic phase of ONE system, so that the CPU test-suite can
// check numbering, constraint maps and the block pattern against the oracle without a GPU.
extern "C" int rve_host_symbolic(const double* xy, int32_t nv, const int32_t* tri, const int32_t* region, int32_t nt,
                                 int32_t model, int32_t vref_nb, int32_t cap_rows, int32_t cap_blocks, int32_t* n_nodes,
                                 int32_t* n_rows, int32_t* n_blocks, int32_t* nc, int32_t* rowptr, int32_t* bcol,
                                 int32_t* row_va, int32_t* row_vb, int32_t* node_row, int32_t* samp_elem, double* samp_lam) {
    SysSym S;
    sym_pass1(S, xy, tri, region, nv, nt);
    if (!S.err.empty()) return RVE_E_MESH;
    SymParams P;
    P.model = model; P.vref_nb = vref_nb; P.vref_box_given = false;
    P.ncx = pick_nc(S.box[2], S.mean_ext); P.ncy = pick_nc(S.box[3], S.mean_ext);
    sym_pass2(S, P);
    if (!S.err.empty()) return RVE_E_MESH;
    if (!sym_selfcheck(S).empty()) return RVE_E_STATE;
    *n_nodes = S.nn; *n_rows = S.na; *n_blocks = (int32_t)S.bcol.size(); *nc = P.ncx;
    if (S.na > cap_rows || (int)S.bcol.size() > cap_blocks) return RVE_E_ARG;
    for (int r = 0; r <= S.na; ++r) rowptr[r] = (int32_t)S.row_ptr[r];
    for (size_t k = 0; k < S.bcol.size(); ++k) bcol[k] = S.bcol[k];
    for (int r = 0; r < S.na; ++r) { row_va[r] = S.node_va[S.row_node[r]]; row_vb[r] = S.node_vb[S.row_node[r]]; }
    for (int n = 0; n < S.nn; ++n) node_row[n] = S.node_row[n];
    if (vref_nb > 0 && samp_elem && samp_lam) {
        const int nvref = 4 * (vref_nb - 1) + 1;
        for (int k = 0; k < nvref; ++k) { samp_elem[k] = S.samp_elem[k]; samp_lam[2 * k] = S.samp_lam[2 * k]; samp_lam[2 * k + 1] = S.samp_lam[2 * k + 1]; }
    }
    return 0;
}
