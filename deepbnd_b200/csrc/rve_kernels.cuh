// rve_kernels.cuh — sm_100a kernels of the batched RVE solver.
//
// Data layout (all systems of a batch concatenated; "row" = active P2 node = 2x2 block row):
//   node_xy[N][2], node_row[N] (global row or -1), node_bnd[N]; elem_node[E][6] (global node ids),
//   elem_reg[E], elem_sys[E].
//   Rows of a system are ordered along the Morton curve of the level-1 coarse cells and cut into
//   WINDOWS of RVE_WIN (=256) rows (one CTA of the PCG kernels each); inside a window rows are
//   sorted by length so that the 32-row SELL slices are uniform.  Matrix storage (SELL-32):
//     slice_ptr[S+1]  offsets in 32-entry groups;  group g = slot j of one slice
//     scol[g][32]     uint16 column, LOCAL to the window's p-tile (own rows, then the halo rows)
//     bval[g][2][32]  double2: (K00,K01) and (K10,K11) of the block of lane's row
//     halo[]          global row ids each window gathers into its shared-memory p-tile
//   PCG vectors p,x,r,q,b[R][K][2] (K = number of right-hand sides, compile-time in the kernels),
//   dinv32[R] float4 (block-Jacobi, preconditioner only), wmean[R], rowinfo[R] (bilinear weights and
//   window-local slots of the 4 level-1 nodes around the row), cn_node/cn_beg/cn_ent (per window: the
//   level-1 nodes it touches and, per node, the (row, corner) pairs restricting to it).
//   Coarse level l (structured (ncx+1)x(ncy+1) grid per system, 25-point block stencil):
//                A[sys][slot 25][4][G] fp64 (set-up), A4/H4[sys][25][G] fp16 blocks (V-cycle), dinv[sys][G][4],
//                V-cycle vectors rc/xc/tc[sys][G][K] float2.
// Per-window partial dot products are combined in a fixed order by the k_fin_* kernels (warp per system).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "../../include/rve_b200.h"
#include "rve_math.cuh"
#include "rve_symbolic.hpp"   // RVE_WIN, RowInfo

#define RVE_MAXLEV 8
#ifndef RVE_OMEGA
#define RVE_OMEGA 0.8      // damped block-Jacobi smoother on the coarse levels
#endif

struct Mat8 { double v[8]; };  // [4][2] lambda*, mu per region

struct Grid {
    int ncx, ncy, nx, ny, G;
    int lox, hix, loy, hiy;  // valid node index range (inclusive)
    int wrap;                // periodic wrap
};

typedef float2 cvec;   // one rhs of a coarse node

// Storage of the coarse 2x2 stencil blocks read by the V-cycle (preconditioner arithmetic only): fp16 by default
// (8 bytes per block: the stencil arrays are the V-cycle's dominant traffic), fp32 with -DRVE_STENCIL_HALF=0.
#ifndef RVE_STENCIL_HALF
#define RVE_STENCIL_HALF 1
#endif
#if RVE_STENCIL_HALF
typedef uint2 stv;
__device__ __forceinline__ stv st_pack(float a, float b, float c, float d) {
    const __half2 lo = __floats2half2_rn(a, b), hi = __floats2half2_rn(c, d);
    return make_uint2(*(const unsigned*)&lo, *(const unsigned*)&hi);
}
__device__ __forceinline__ float4 st_unpack(stv v) {
    const float2 lo = __half22float2(*(const __half2*)&v.x), hi = __half22float2(*(const __half2*)&v.y);
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}
#else
typedef float4 stv;
__device__ __forceinline__ stv st_pack(float a, float b, float c, float d) { return make_float4(a, b, c, d); }
__device__ __forceinline__ float4 st_unpack(stv v) { return v; }
#endif

struct Levels {
    int L;          // number of coarse levels
    int K;          // rhs count
    Grid g[RVE_MAXLEV];
    double* A[RVE_MAXLEV];      // [n_sys][25][4][G]
    double* dinv[RVE_MAXLEV];   // [n_sys][G][4]
    float* rc[RVE_MAXLEV];      // [n_sys][G][K][2]  the V-cycle runs in fp32 (preconditioner arithmetic only)
    float* xc[RVE_MAXLEV];
    float* tc[RVE_MAXLEV];
    stv* A4[RVE_MAXLEV];        // [n_sys][25][G] reduced-precision copy of A (a00,a01,a10,a11): preconditioner arithmetic only
    stv* H4[RVE_MAXLEV];        // [n_sys][25][G] A_ij * (omega Dinv_j): residual after the first Jacobi step
};

struct WinTab {
    const int* win_sys;      // [n_win]
    const int* win_row0;     // global first row
    const int* win_n;        // rows in the window (<= RVE_WIN)
    const int* win_slice0;   // global first slice (ceil(n/32) slices)
    const int* win_halo0;    // [n_win] offset of the window's halo list (segments are claimed on the device)
    const int* win_nhalo;    // [n_win] halo rows of the window
    const int* win_cn0;      // [n_win] offset of the window's coarse-node list (claimed on the device; ncn + 1 slots)
    const int* win_ncn;      // [n_win] level-1 nodes the window touches
    unsigned char* win_active;   // [n_win] copy of active[win_sys[w]]: the PCG kernels test it without the hop over win_sys
    const int* sys_win0;     // [n_sys+1]
    const int* row_base;     // [n_sys+1] first global row of a system
};

struct Sell {
    const unsigned* slice_ptr;
    const uint16_t* scol;
    const int* halo;
};

// everything the streaming PCG kernels need to know about a window, in one 32-byte record (written by k_sym_coarse):
// the producer thread of the persistent kernels turns it into bulk copies without a second dependent load
struct WinDesc { int row0, n, cn0, ncn; unsigned ent0, nent; int sys, pad; };

// per (system, rhs) scalars
#define SC_RZ 0
#define SC_PQ 1
#define SC_ALPHA 2
#define SC_BETA 3
#define SC_BB 4
#define SC_RR 5
#define SC_RDR 6
#define SC_RZOLD 7
#define SC_BREF 8   // sum of squared UNASSEMBLED element load contributions (absolute floor)
#define SC_N 10

__device__ __forceinline__ bool grid_node(const Grid& g, int i, int j, int& idx) {
    if (g.wrap) {
        i = i % g.ncx; if (i < 0) i += g.ncx;
        j = j % g.ncy; if (j < 0) j += g.ncy;
    } else if (i < g.lox || i > g.hix || j < g.loy || j > g.hiy) {
        return false;
    }
    idx = j * g.nx + i;
    return true;
}

// same as grid_node for |offset| <= ncx (one conditional wrap instead of two integer modulos):
// every index the V-cycle kernels form is within [-2, ncx+2] and ncx >= 2
__device__ __forceinline__ bool grid_node_fast(const Grid& g, int i, int j, int& idx) {
    if (g.wrap) {
        i += (i < 0) ? g.ncx : 0; i -= (i >= g.ncx) ? g.ncx : 0;
        j += (j < 0) ? g.ncy : 0; j -= (j >= g.ncy) ? g.ncy : 0;
    } else if (i < g.lox || i > g.hix || j < g.loy || j > g.hiy) {
        return false;
    }
    idx = j * g.nx + i;
    return true;
}

__device__ __forceinline__ bool grid_valid(const Grid& g, int i, int j) {
    return i >= g.lox && i <= g.hix && j >= g.loy && j <= g.hiy;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------
// (0) second half of the symbolic phase (the first half — numbering, row order, element fans — is rve_devsym.cuh).
//     From the capacity of every row and the elements around it the GPU finds the block columns (k_sym_cols) and,
//     per window, the halo list and the window-local 16-bit SELL columns (k_sym_window: shared-memory bitmap over
//     the rows of the system, popcount prefix = rank of a halo row; halo segments are claimed with one atomic per
//     window), then the level-1 nodes a window touches and its restriction lists (k_sym_coarse).
// ---------------------------------------------------------------------------------------------
// counters: [0] halo entries used, [1] max rows of a p-tile (own + halo), [2] error code
// Launched with RVE_SYMCOLS_SPLIT CTAs of RVE_WIN / RVE_SYMCOLS_SPLIT threads per window: the rows of a window are sorted by
// length, so its two vertex-row warps work 5-10x longer than its six mid-edge warps; as one 256-thread CTA the six wait
// at the final barrier and hold the SM's warp slots (62 % of the stall samples, r02h), as 64-thread CTAs heavy and light
// pieces of different windows mix freely.
#define RVE_SYMCOLS_SPLIT 4
__global__ void __launch_bounds__(RVE_WIN / RVE_SYMCOLS_SPLIT)
k_sym_cols(WinTab wt, const unsigned* __restrict__ row_ptr, const unsigned* __restrict__ adjf_ptr,
           const int* __restrict__ adjf, const int* __restrict__ elem_node, const int* __restrict__ node_row,
           int* __restrict__ csr_col, unsigned char* __restrict__ rlen, unsigned char* __restrict__ emap,
           int* __restrict__ sys_blocks, int* counters) {
    // emap == nullptr: columns, lengths and block counts only (the element -> slot map serves the ASSEMBLED operator; it is
    // 36 single-byte stores per element and is built on demand, with sys_blocks == nullptr, by ensure_emap)
    constexpr int TPB = RVE_WIN / RVE_SYMCOLS_SPLIT;
    constexpr int SC = 32;
    __shared__ int s_cols[SC * TPB];
    const int w = blockIdx.x / RVE_SYMCOLS_SPLIT, t = (blockIdx.x % RVE_SYMCOLS_SPLIT) * TPB + threadIdx.x;
    int cnt = 0;
    if (t < wt.win_n[w]) {
        const int r = wt.win_row0[w] + t;
        const unsigned o = row_ptr[r];
        const int cap = (int)(row_ptr[r + 1] - o);   // capacity of the row (>= its true length)
        int* cols = csr_col + o;                     // unique rows in discovery order (same order as rve_symbolic.hpp)
        unsigned long long seen = 0ull;
        for (unsigned i = adjf_ptr[r]; i < adjf_ptr[r + 1]; ++i) {
            const int e = adjf[i];
            const int* en = elem_node + 6 * (size_t)e;
            int rows[6], a = 0;
#pragma unroll
            for (int k = 0; k < 6; ++k) { rows[k] = node_row[en[k]]; if (rows[k] == r) a = k; }
            // emap[e][a][k] = SELL slot of block (row of local node a, row of local node k): k_assemble scatters
            // without searching
            unsigned char* em = emap ? emap + (size_t)e * 36 + a * 6 : nullptr;
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const int c = rows[k];
                if (c < 0) { if (em) em[k] = 0xff; continue; }
                // the columns found so far are searched in a shared-memory column of the thread (searching the global
                // list re-read freshly written lines from L2: one L2 round trip per probe); rows longer than SC fall
                // back to the global list for the tail
                // (a 64-bit filter of the columns seen so far: a column whose bit is clear is new, no search; about half of
                // the look-ups of a fan are first sightings and used to scan the whole list)
                const unsigned long long bit = 1ull << (c & 63);
                int j = cnt;
                if (seen & bit) { j = 0; while (j < cnt && j < cap && (j < SC ? s_cols[j * TPB + threadIdx.x] : cols[j]) != c) ++j; }
                seen |= bit;
                if (j == cnt) {
                    if (cnt < cap) { cols[cnt] = c; if (cnt < SC) s_cols[cnt * TPB + threadIdx.x] = c; }
                    ++cnt;
                }
                if (em) em[k] = (unsigned char)j;
            }
        }
        if (cnt > cap) { atomicExch(counters + 2, 2); cnt = cap; }
        rlen[r] = (unsigned char)cnt;
    }
    // true number of blocks of the system (one integer atomic per warp, no block barrier)
    int v = cnt;
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o2);
    if (sys_blocks && (threadIdx.x & 31) == 0 && v) atomicAdd(sys_blocks + wt.win_sys[w], v);
}

__global__ void __launch_bounds__(RVE_WIN)
k_sym_window(WinTab wt, int* __restrict__ win_halo0, int* __restrict__ win_nhalo, const unsigned* __restrict__ row_ptr,
             const unsigned char* __restrict__ rlen, const int* __restrict__ csr_col, const unsigned* __restrict__ slice_ptr, uint16_t* __restrict__ scol,
             int* __restrict__ halo, int halo_cap, int* counters) {
    extern __shared__ unsigned s_bm[];           // [words] bitmap of halo rows, then [words] popcount prefix
    __shared__ int s_scan[RVE_WIN];
    __shared__ int s_h0;
    const int w = blockIdx.x, tid = threadIdx.x;
    const int sys = wt.win_sys[w];
    const int rb0 = wt.row_base[sys], na = wt.row_base[sys + 1] - rb0;
    const int words = (na + 31) >> 5;
    unsigned* s_pre = s_bm + words;
    const int r0 = wt.win_row0[w], n = wt.win_n[w], r0l = r0 - rb0;
    for (int i = tid; i < words; i += RVE_WIN) s_bm[i] = 0u;
    __syncthreads();
    unsigned o = 0; int len = 0;
    if (tid < n) {
        o = row_ptr[r0 + tid]; len = rlen[r0 + tid];
        for (int j = 0; j < len; ++j) {
            const int c = csr_col[o + j] - rb0;
            if (c < r0l || c >= r0l + n) atomicOr(&s_bm[c >> 5], 1u << (c & 31));
        }
    }
    __syncthreads();
    // exclusive popcount prefix over the words: contiguous span per thread, block scan of the span sums
    const int wpt = (words + RVE_WIN - 1) / RVE_WIN;
    const int wa = min(words, tid * wpt), wb = min(words, wa + wpt);
    int sum = 0;
    for (int i = wa; i < wb; ++i) sum += __popc(s_bm[i]);
    s_scan[tid] = sum;
    __syncthreads();
    for (int d = 1; d < RVE_WIN; d <<= 1) {
        const int v = tid >= d ? s_scan[tid - d] : 0;
        __syncthreads();
        s_scan[tid] += v;
        __syncthreads();
    }
    int run = s_scan[tid] - sum;                  // exclusive
    for (int i = wa; i < wb; ++i) { s_pre[i] = (unsigned)run; run += __popc(s_bm[i]); }
    const int nh = s_scan[RVE_WIN - 1];
    if (tid == 0) {
        int h0 = atomicAdd(counters, nh);
        atomicMax(counters + 1, n + nh);
        if (h0 + nh > halo_cap || n + nh > 65535) { atomicExch(counters + 2, 3); h0 = -1; }
        s_h0 = h0;
        win_halo0[w] = h0 < 0 ? 0 : h0;
        win_nhalo[w] = h0 < 0 ? 0 : nh;
    }
    __syncthreads();
    const int h0 = s_h0;
    if (h0 < 0) return;
    for (int i = tid; i < words; i += RVE_WIN) {
        unsigned bits = s_bm[i];
        int pos = h0 + (int)s_pre[i];
        while (bits) {
            const int b = __ffs(bits) - 1;
            halo[pos++] = rb0 + i * 32 + b;
            bits &= bits - 1;
        }
    }
    const int nsl = (n + 31) >> 5;
    if (tid < nsl * 32) {
        const int slice = wt.win_slice0[w] + (tid >> 5), lane = tid & 31;
        const unsigned g0 = slice_ptr[slice];
        const int width = (int)(slice_ptr[slice + 1] - g0);
        uint16_t* sc = scol + (size_t)g0 * 32 + lane;
        for (int j = 0; j < len; ++j) {
            const int c = csr_col[o + j] - rb0;
            const int lc = (c >= r0l && c < r0l + n) ? c - r0l
                                                     : n + (int)s_pre[c >> 5] + __popc(s_bm[c >> 5] & ((1u << (c & 31)) - 1u));
            sc[(size_t)j * 32] = (uint16_t)lc;
        }
        const uint16_t self = (uint16_t)(tid < n ? tid : 0);     // padding: zero block against a valid tile entry
        for (int j = len; j < width; ++j) sc[(size_t)j * 32] = self;
    }
}

// Restriction / prolongation data of a window (fine rows <-> level-1 grid): bilinear position (s,t) of every row in
// its coarse cell, the sorted list of level-1 nodes the window touches (bitmap + popcount ranks, like the halo), the
// window-local slot of the 4 corners of every row, and per coarse node the (row, corner) pairs restricting to it
// (counting sort in shared memory, each list then ordered so that the summation order is reproducible).
// counters: [3] coarse-node slots used, [4] list entries used, [5] max nodes of a window
__global__ void __launch_bounds__(RVE_WIN)
k_sym_coarse(WinTab wt, int* __restrict__ win_cn0, int* __restrict__ win_ncn, const int* __restrict__ row_node,
             const double* __restrict__ node_xy, const double* __restrict__ box, Grid g1, int model,
             RowInfo* __restrict__ rowinfo, int* __restrict__ cn_node, unsigned* __restrict__ cn_beg,
             uint16_t* __restrict__ cn_ent, int cn_cap, int* counters, WinDesc* __restrict__ desc) {
    extern __shared__ unsigned s_bm[];           // [words] bitmap of touched level-1 nodes, then [words] popcount prefix
    __shared__ int s_scan[RVE_WIN];
    __shared__ int s_base[2];
    __shared__ int s_cnt[4 * RVE_WIN + 1];       // per-slot counts / offsets (ncn <= 4 * rows)
    const int w = blockIdx.x, tid = threadIdx.x;
    const int sys = wt.win_sys[w];
    const int words = (g1.G + 31) >> 5;
    unsigned* s_pre = s_bm + words;
    const int r0 = wt.win_row0[w], n = wt.win_n[w];
    for (int i = tid; i < words; i += RVE_WIN) s_bm[i] = 0u;
    for (int i = tid; i <= 4 * RVE_WIN; i += RVE_WIN) s_cnt[i] = 0;
    __syncthreads();
    int nd[4] = {-1, -1, -1, -1};
    float fs = 0.f, ft = 0.f;
    if (tid < n) {
        const int node = row_node[r0 + tid];
        int ci, cj; double s, t;
        coarse_cell(node_xy[2 * (size_t)node], node_xy[2 * (size_t)node + 1], box[4 * sys], box[4 * sys + 1],
                    g1.ncx / box[4 * sys + 2], g1.ncy / box[4 * sys + 3], g1.ncx, g1.ncy, ci, cj, s, t);
        fs = (float)s; ft = (float)t;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            int idx;
            if (grid_node_fast(g1, ci + (c & 1), cj + (c >> 1), idx)) { nd[c] = idx; atomicOr(&s_bm[idx >> 5], 1u << (idx & 31)); }
        }
    }
    __syncthreads();
    const int wpt = (words + RVE_WIN - 1) / RVE_WIN;
    const int wa = min(words, tid * wpt), wb = min(words, wa + wpt);
    int sum = 0;
    for (int i = wa; i < wb; ++i) sum += __popc(s_bm[i]);
    s_scan[tid] = sum;
    __syncthreads();
    for (int d = 1; d < RVE_WIN; d <<= 1) {
        const int v = tid >= d ? s_scan[tid - d] : 0;
        __syncthreads();
        s_scan[tid] += v;
        __syncthreads();
    }
    int run = s_scan[tid] - sum;
    for (int i = wa; i < wb; ++i) { s_pre[i] = (unsigned)run; run += __popc(s_bm[i]); }
    const int ncn = s_scan[RVE_WIN - 1];
    __syncthreads();
    // slots of the corners, per-slot counts
    int slot[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        slot[c] = -1;
        if (nd[c] >= 0) {
            slot[c] = (int)s_pre[nd[c] >> 5] + __popc(s_bm[nd[c] >> 5] & ((1u << (nd[c] & 31)) - 1u));
            atomicAdd(&s_cnt[slot[c] + 1], 1);
        }
    }
    __syncthreads();
    {   // offsets: s_cnt[j + 1] = inclusive sum of the counts (4 consecutive lists per thread, warp scans, 8 warp totals)
        __shared__ int s_wsum[RVE_WIN / 32];
        const int b4 = 4 * tid;
        int a[4], tsum = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) { a[i] = (b4 + i < ncn) ? s_cnt[b4 + i + 1] : 0; tsum += a[i]; }
        int inc = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if ((tid & 31) >= o) inc += u; }
        if ((tid & 31) == 31) s_wsum[tid >> 5] = inc;
        __syncthreads();
        int run2 = inc - tsum;
        for (int i = 0; i < (tid >> 5); ++i) run2 += s_wsum[i];
#pragma unroll
        for (int i = 0; i < 4; ++i) { run2 += a[i]; if (b4 + i < ncn) s_cnt[b4 + i + 1] = run2; }
    }
    __syncthreads();
    if (tid == 0) {
        const int nent = s_cnt[ncn];
        // segments start on 16-byte boundaries (4 ints / 8 uint16): the persistent PCG kernels fetch them with bulk copies
        int c0 = atomicAdd(counters + 3, (ncn + 1 + 3) & ~3);
        int e0 = atomicAdd(counters + 4, (nent + 7) & ~7);
        atomicMax(counters + 5, ncn);
        if (c0 + ((ncn + 1 + 3) & ~3) > cn_cap) { atomicExch(counters + 2, 4); c0 = -1; }
        s_base[0] = c0; s_base[1] = e0;
        win_cn0[w] = c0 < 0 ? 0 : c0;
        win_ncn[w] = c0 < 0 ? 0 : ncn;
        WinDesc d;
        d.row0 = r0; d.n = n; d.cn0 = c0 < 0 ? 0 : c0; d.ncn = c0 < 0 ? 0 : ncn; d.ent0 = (unsigned)e0; d.nent = (unsigned)nent; d.sys = sys; d.pad = 0;
        desc[w] = d;
    }
    __syncthreads();
    const int c0 = s_base[0], e0 = s_base[1];
    if (c0 < 0) return;
    for (int i = tid; i < words; i += RVE_WIN) {
        unsigned bits = s_bm[i];
        int pos = c0 + (int)s_pre[i];
        while (bits) {
            const int b = __ffs(bits) - 1;
            cn_node[pos++] = i * 32 + b;
            bits &= bits - 1;
        }
    }
    for (int j = tid; j <= ncn; j += RVE_WIN) cn_beg[c0 + j] = (unsigned)(e0 + s_cnt[j]);
    __syncthreads();
    // fill in shared memory (position inside a list by atomics on a second counter array), order every list there
    // (reproducible summation order in the restriction), then write the window's entries out in one coalesced pass
    // (the lists used to be ordered by one insertion sort per list: a few long lists kept the CTA waiting, 67 % of the
    // stall samples at the barrier behind them, r02h; now every entry finds its rank in its list by counting the smaller
    // ones -- four entries per thread, same work for all)
    __shared__ int s_fill[4 * RVE_WIN];
    __shared__ uint16_t s_entl[4 * RVE_WIN], s_ents[4 * RVE_WIN];
    for (int j = tid; j < ncn; j += RVE_WIN) s_fill[j] = 0;
    __syncthreads();
    if (tid < n) {
        RowInfo ri;
        ri.s = fs; ri.t = ft;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            ri.slot[c] = slot[c] < 0 ? (uint16_t)0xffffu : (uint16_t)slot[c];
            if (slot[c] >= 0) {
                const int pos = atomicAdd(&s_fill[slot[c]], 1);
                s_entl[s_cnt[slot[c]] + pos] = (uint16_t)((tid << 2) | c);
            }
        }
        rowinfo[r0 + tid] = ri;
    }
    __syncthreads();
    if (tid < n) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (slot[c] < 0) continue;
            const int b = s_cnt[slot[c]], e = s_cnt[slot[c] + 1];
            const uint16_t v = (uint16_t)((tid << 2) | c);
            int rank = 0;
            for (int i = b; i < e; ++i) rank += (s_entl[i] < v) ? 1 : 0;
            s_ents[b + rank] = v;
        }
    }
    __syncthreads();
    const int nent = s_cnt[ncn];
    for (int i = tid; i < nent; i += RVE_WIN) cn_ent[(size_t)e0 + i] = s_ents[i];
}

// ---------------------------------------------------------------------------------------------
// (0b) shared topology (rve_set_batch_shared): the symbolic phase ran for system 0; its index arrays are replicated for
//      the other systems by shifting the global indices, the node coordinates are rebuilt from every system's vertices
// ---------------------------------------------------------------------------------------------
template <class T>
__global__ void k_rep_add(T* __restrict__ dst, const T* __restrict__ src, long long n1, int n_sys, long long add, int keep_neg) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n1 * n_sys) return;
    const long long s = gid / n1;
    const T v = src[gid - s * n1];
    dst[gid] = (keep_neg && v < (T)0) ? v : (T)(v + s * add);
}
// row-pointer style arrays (n1 + 1 entries per system, the last one = the next system's first)
__global__ void k_rep_ptr(unsigned* __restrict__ dst, const unsigned* __restrict__ src, long long n1, int n_sys, long long add) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid > n1 * n_sys) return;
    if (gid == n1 * n_sys) { dst[gid] = (unsigned)(src[n1] + (n_sys - 1) * add); return; }
    const long long s = gid / n1;
    dst[gid] = (unsigned)(src[gid - s * n1] + s * add);
}
__global__ void k_rep_sys(int* __restrict__ dst, long long n1, int n_sys) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < n1 * n_sys) dst[gid] = (int)(gid / n1);
}
// node coordinates of every system from its own vertices: vertex nodes first, then the mid-edge nodes (va, vb)
__global__ void k_shared_node_xy(int n_sys, int nn, int nv, const double* __restrict__ xy /*[n_sys][nv][2]*/,
                                 const int* __restrict__ node_va, const int* __restrict__ node_vb, double* __restrict__ node_xy) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)n_sys * nn) return;
    const int s = (int)(gid / nn), n = (int)(gid - (long long)s * nn);
    const double* x = xy + 2 * (size_t)s * nv;
    const int a = node_va[n], b = node_vb[n];
    node_xy[2 * gid] = 0.5 * (x[2 * a] + x[2 * b]);
    node_xy[2 * gid + 1] = 0.5 * (x[2 * a + 1] + x[2 * b + 1]);
}

// Radial morphing of a template mesh on the device (deepbnd_b200/mesh/rve_morph.py): a moving node at distance d and
// direction u from the centre c of its nearest inclusion goes to c + u d', d' = d r / r0 inside the template radius r0 and
// r + (d - r0)(R - r)/(R - r0) in the annulus up to R; r = the inclusion's radius, or the radius of its ellipse
// (semi-axes l, e l, angle theta: ellipse2_mesh.py:27-39) in the direction u.  Thread per (system, moving node).
struct MorphDev {
    const int* node;       // [n_mov] vertex id
    const int* incl;       // [n_mov] nearest inclusion
    const double* d;       // [n_mov]
    const double* dir;     // [n_mov][2]
    const double* centre;  // [n_incl][2]
    const double* prm;     // [n_sys][n_incl][3]  l, e, theta
    int n_mov, n_incl;
    double r0, R;
};
__global__ void k_morph(int n_sys, int nv, MorphDev M, double* __restrict__ xy /*[n_sys][nv][2], pre-filled with the template*/) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)n_sys * M.n_mov) return;
    const int s = (int)(gid / M.n_mov), j = (int)(gid - (long long)s * M.n_mov);
    const int k = M.incl[j];
    const double* pr = M.prm + 3 * ((size_t)s * M.n_incl + k);
    const double ux = M.dir[2 * j], uy = M.dir[2 * j + 1], d = M.d[j];
    double r = pr[0];
    if (pr[1] != 1.0) {
        const double a = pr[0], b = pr[1] * pr[0];
        const double phi = atan2(uy, ux) - pr[2];
        const double cb = b * cos(phi), sa = a * sin(phi);
        r = a * b / sqrt(cb * cb + sa * sa);
    }
    const double dn = (d <= M.r0) ? d * r / M.r0 : r + (d - M.r0) * (M.R - r) / (M.R - M.r0);
    double* o = xy + 2 * ((size_t)s * nv + M.node[j]);
    o[0] = M.centre[2 * k] + ux * dn;
    o[1] = M.centre[2 * k + 1] + uy * dn;
}

// ---------------------------------------------------------------------------------------------
// (1) numeric assembly, row-wise: thread per block row (CTA per window, warp per 32-row SELL slice).  A thread
//     zeroes the slots of its row, walks the elements around the row (adjf), computes the 6 blocks K_e[a][0..5] of
//     its own local node a (closed-form P2, branch-free gradient table) and adds them into the SELL slots named
//     by the element map of k_sym_cols.  One writer per row: no atomics, no memset of the 128-byte groups, the
//     matrix leaves the L2 once (the element-wise scatter with fp64 atomics it replaces moved 3x the bytes).
//     Also int phi_a dx per row (mean-value constraint of 'per' / 'MR').  Replaces mp.block_assemble(a)
//     (micro_model.py:60).
// ---------------------------------------------------------------------------------------------
// grad phi_a at quadrature point q = sum_i c[q][a][i] grad lambda_i  (points (1/2,1/2,0), (0,1/2,1/2), (1/2,0,1/2))
__constant__ double c_p2grad[3][6][3] = {
    {{1, 0, 0}, {0, 1, 0}, {0, 0, -1}, {2, 2, 0}, {0, 0, 2}, {0, 0, 2}},
    {{-1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {2, 0, 0}, {0, 2, 2}, {2, 0, 0}},
    {{1, 0, 0}, {0, -1, 0}, {0, 0, 1}, {0, 2, 0}, {0, 2, 0}, {2, 0, 2}}};

__global__ void __launch_bounds__(RVE_WIN, 3)
k_assemble(WinTab wt, Sell sl, const unsigned* __restrict__ adjf_ptr, const int* __restrict__ adjf,
           const int* __restrict__ elem_node, const unsigned char* __restrict__ elem_reg,
           const double* __restrict__ node_xy, const int* __restrict__ node_row,
           const unsigned char* __restrict__ emap, double* __restrict__ bval, double* __restrict__ wmean, Mat8 mat) {
    __shared__ double s_mat[8];
    __shared__ double s_c[3][6][3];                     // the table in shared memory: lanes index it with different a
    if (threadIdx.x < 8) s_mat[threadIdx.x] = mat.v[threadIdx.x];
    if (threadIdx.x < 54) (&s_c[0][0][0])[threadIdx.x] = (&c_p2grad[0][0][0])[threadIdx.x];
    __syncthreads();
    const int w = blockIdx.x, t = threadIdx.x, lane = t & 31;
    const int n = wt.win_n[w];
    if ((t >> 5) >= ((n + 31) >> 5)) return;            // no such slice in this window
    const int slice = wt.win_slice0[w] + (t >> 5);
    const size_t g0 = sl.slice_ptr[slice];
    const int width = (int)(sl.slice_ptr[slice + 1] - g0);
    double* brow = bval + g0 * 128 + 2 * lane;            // slot j of this row: brow[j*128 + {0,1}], brow[j*128 + 64 + {0,1}]
    for (int j = 0; j < width; ++j) {                     // rows beyond n (last slice) stay zero
        *(double2*)(brow + (size_t)j * 128) = make_double2(0.0, 0.0);
        *(double2*)(brow + (size_t)j * 128 + 64) = make_double2(0.0, 0.0);
    }
    if (t >= n) return;
    const int r = wt.win_row0[w] + t;
    double wm = 0.0;
    for (unsigned i = adjf_ptr[r]; i < adjf_ptr[r + 1]; ++i) {
        const int e = adjf[i];
        const int* en = elem_node + 6 * (size_t)e;
        int a = 0;
#pragma unroll
        for (int k = 0; k < 6; ++k) if (node_row[en[k]] == r) a = k;
        const double* p0 = node_xy + 2 * (size_t)en[0]; const double* p1 = node_xy + 2 * (size_t)en[1]; const double* p2 = node_xy + 2 * (size_t)en[2];
        TriGeom g;
        tri_geom(p0[0], p0[1], p1[0], p1[1], p2[0], p2[1], g);
        const int reg = elem_reg[e];
        const double lam = s_mat[2 * reg], mu = s_mat[2 * reg + 1], wq = g.area * (1.0 / 3.0);
        if (a >= 3) wm += wq;
        double ax[3], ay[3];                               // grad phi_a at the 3 quadrature points
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const double* c = s_c[q][a];
            ax[q] = c[0] * g.gl[0][0] + c[1] * g.gl[1][0] + c[2] * g.gl[2][0];
            ay[q] = c[0] * g.gl[0][1] + c[1] * g.gl[1][1] + c[2] * g.gl[2][1];
        }
        const unsigned char* em = emap + (size_t)e * 36 + a * 6;
#pragma unroll 2
        for (int k = 0; k < 6; ++k) {
            const int j = em[k];
            if (j == 0xff) continue;
            double G00 = 0, G01 = 0, G10 = 0, G11 = 0;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const double* c = s_c[q][k];
                const double bx = c[0] * g.gl[0][0] + c[1] * g.gl[1][0] + c[2] * g.gl[2][0];
                const double by = c[0] * g.gl[0][1] + c[1] * g.gl[1][1] + c[2] * g.gl[2][1];
                G00 += ax[q] * bx; G01 += ax[q] * by; G10 += ay[q] * bx; G11 += ay[q] * by;
            }
            G00 *= wq; G01 *= wq; G10 *= wq; G11 *= wq;
            const double tr = G00 + G11;
            double2* d0 = (double2*)(brow + (size_t)j * 128);
            double2* d1 = (double2*)(brow + (size_t)j * 128 + 64);
            double2 v0 = *d0, v1 = *d1;
            v0.x += lam * G00 + mu * G00 + mu * tr; v0.y += lam * G01 + mu * G10;
            v1.x += lam * G10 + mu * G01;           v1.y += lam * G11 + mu * G11 + mu * tr;
            *d0 = v0; *d1 = v1;
        }
    }
    wmean[r] = wm;
}

// 2x2 inverse of the diagonal blocks (block-Jacobi), stored in fp32: CTA per window, thread per row
__global__ void __launch_bounds__(RVE_WIN)
k_dinv(WinTab wt, Sell sl, const double* __restrict__ bval, float4* __restrict__ dinv32) {
    const int w = blockIdx.x, t = threadIdx.x;
    if (t >= wt.win_n[w]) return;
    const int lane = t & 31, slice = wt.win_slice0[w] + (t >> 5);
    const unsigned g0 = sl.slice_ptr[slice], g1 = sl.slice_ptr[slice + 1];
    double a = 0, b = 0, c = 0, e = 0;
    for (unsigned g = g0; g < g1; ++g)
        if (sl.scol[(size_t)g * 32 + lane] == t) {
            const double* d = bval + (size_t)g * 128 + 2 * lane;
            a = d[0]; b = d[1]; c = d[64]; e = d[65];
            break;
        }
    const double det = a * e - b * c;
    const double inv = det != 0.0 ? 1.0 / det : 0.0;
    dinv32[wt.win_row0[w] + t] = make_float4((float)(e * inv), (float)(-b * inv), (float)(-c * inv), (float)(a * inv));
}

// ---------------------------------------------------------------------------------------------
// (2) level-1 Galerkin operator A1 = Z^T A Z, element by element, warp per element.  The element stiffness is never
//     formed: with the bilinear weights W[a][I] of the element's 6 nodes on its 3x3 window of coarse nodes, the
//     coarse basis function of node I restricted to the element has the gradient gI(q) = sum_a W[a][I] grad phi_a(q)
//     at the 3 quadrature points, and the 2x2 block of the pair (I, J) is
//        sum_q w [ lam gI (x) gJ + mu gJ (x) gI + mu (gI . gJ) I ]
//     (p2_stiff_block with the coarse gradients in place of the P2 ones).  A1 is symmetric: only the blocks with
//     offset (dj > 0) or (dj == 0, di >= 0) are accumulated (atomics into the 25-slot stencil arrays),
//     k_galerkin_mirror fills the other 12 slots with the transposes.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_galerkin1(int n_elem, const int* __restrict__ elem_node, const unsigned char* __restrict__ elem_reg,
            const int* __restrict__ elem_sys, const double* __restrict__ node_xy,
            const int* __restrict__ node_row, const double* __restrict__ box, Grid g1,
            double* __restrict__ A1, Mat8 mat) {
    __shared__ double s_mat[8];
    __shared__ double s_W[8][6][9];
    __shared__ double s_G[8][3][9][2];          // coarse gradients gI(q)
    if (threadIdx.x < 8) s_mat[threadIdx.x] = mat.v[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int e = blockIdx.x * 8 + w;
    if (e >= n_elem) return;
    const int sys = elem_sys[e];
    int node = 0, row = -1;
    double px = 0, py = 0;
    if (lane < 6) {
        node = elem_node[(size_t)e * 6 + lane];
        row = node_row[node];
        px = node_xy[2 * (size_t)node];
        py = node_xy[2 * (size_t)node + 1];
    }
    double x0 = __shfl_sync(0xffffffffu, px, 0), y0 = __shfl_sync(0xffffffffu, py, 0);
    double x1 = __shfl_sync(0xffffffffu, px, 1), y1 = __shfl_sync(0xffffffffu, py, 1);
    double x2 = __shfl_sync(0xffffffffu, px, 2), y2 = __shfl_sync(0xffffffffu, py, 2);
    TriGeom g;
    tri_geom(x0, y0, x1, y1, x2, y2, g);
    const int reg = elem_reg[e];
    const double lam = s_mat[2 * reg], mu = s_mat[2 * reg + 1];
    // coarse cell of every node
    const double bx0 = box[4 * sys], by0 = box[4 * sys + 1];
    const double invHx = g1.ncx / box[4 * sys + 2], invHy = g1.ncy / box[4 * sys + 3];
    int ci = 1 << 20, cj = 1 << 20;
    double s = 0, t = 0;
    if (lane < 6) coarse_cell(px, py, bx0, by0, invHx, invHy, g1.ncx, g1.ncy, ci, cj, s, t);
    int imin = ci, jmin = cj;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        imin = min(imin, __shfl_xor_sync(0xffffffffu, imin, o));
        jmin = min(jmin, __shfl_xor_sync(0xffffffffu, jmin, o));
    }
    imin = __shfl_sync(0xffffffffu, imin, 0);
    jmin = __shfl_sync(0xffffffffu, jmin, 0);
    if (lane < 6) {
#pragma unroll
        for (int k = 0; k < 9; ++k) s_W[w][lane][k] = 0.0;
        // an element larger than a coarse cell can reach a third cell: its far nodes are snapped onto the
        // last coarse-node line of the element's 3x3 window (the operator stays a sum of SPSD terms)
        int oi = ci - imin, oj = cj - jmin;
        if (oi > 1) { oi = 1; s = 1.0; }
        if (oj > 1) { oj = 1; t = 1.0; }
        if (row >= 0) {
            s_W[w][lane][oj * 3 + oi] = (1 - s) * (1 - t);
            s_W[w][lane][oj * 3 + oi + 1] = s * (1 - t);
            s_W[w][lane][(oj + 1) * 3 + oi] = (1 - s) * t;
            s_W[w][lane][(oj + 1) * 3 + oi + 1] = s * t;
        }
    }
    __syncwarp();
    // gI(q) for the 9 window nodes x 3 quadrature points (27 lanes); window nodes that carry weight
    double wsum = 0.0;
    if (lane < 27) {
        const int I = lane % 9, q = lane / 9;
        double gx = 0.0, gy = 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            const double wa = s_W[w][a][I];
            double ax, ay;
            p2_grad(g, q, a, ax, ay);
            gx += wa * ax; gy += wa * ay;
            wsum += wa;
        }
        s_G[w][q][I][0] = gx; s_G[w][q][I][1] = gy;
    }
    __syncwarp();
    double* Asys = A1 + (size_t)sys * 100 * g1.G;
    const unsigned am = __ballot_sync(0xffffffffu, lane < 9 && wsum != 0.0);
    const int na = __popc(am);
    const double wq = g.area * (1.0 / 3.0);
    // one pair of weighted window nodes per lane (4 nodes for an element inside one cell, up to 9)
    for (int pr = lane; pr < na * na; pr += 32) {
        const int I = __fns(am, 0, pr / na + 1), J = __fns(am, 0, pr % na + 1);
        int Ii = imin + I % 3, Ij = jmin + I / 3, Ji = imin + J % 3, Jj = jmin + J / 3;
        if (Jj < Ij || (Jj == Ij && Ji < Ii)) continue;     // lower half: k_galerkin_mirror
        int nI, nJ;
        if (!grid_node(g1, Ii, Ij, nI) || !grid_node(g1, Ji, Jj, nJ)) continue;
        double G00 = 0, G01 = 0, G10 = 0, G11 = 0;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const double ix = s_G[w][q][I][0], iy = s_G[w][q][I][1], jx = s_G[w][q][J][0], jy = s_G[w][q][J][1];
            G00 += ix * jx; G01 += ix * jy; G10 += iy * jx; G11 += iy * jy;
        }
        G00 *= wq; G01 *= wq; G10 *= wq; G11 *= wq;
        const double tr = G00 + G11;
        const int slot = (Jj - Ij + 2) * 5 + (Ji - Ii + 2);
        double* d = Asys + (size_t)slot * 4 * g1.G + nI;
        atomicAdd(d, lam * G00 + mu * G00 + mu * tr); atomicAdd(d + g1.G, lam * G01 + mu * G10);
        atomicAdd(d + 2 * (size_t)g1.G, lam * G10 + mu * G01); atomicAdd(d + 3 * (size_t)g1.G, lam * G11 + mu * G11 + mu * tr);
    }
}

// Thread-per-element version of k_galerkin1 (same sums, default).  The warp-per-element kernel above spends 855 warp
// instructions per element with 6 to 27 lanes busy and is issue-bound (67 % issue-active, 4.9 ms per 250 full RVEs:
// profiles/r02f_ncu_galerkin1_summary.csv); here a thread owns an element, its 3 x 9 coarse gradients live in a
// shared-memory column (54 doubles, [entry][thread]: conflict-free) and the weighted window nodes are walked as bit sets.
#ifndef RVE_GALERKIN_THREAD
#define RVE_GALERKIN_THREAD 1
#endif
#define RVE_GAL_THREADS 128
__global__ void __launch_bounds__(RVE_GAL_THREADS)
k_galerkin1_t(int n_elem, const int* __restrict__ elem_node, const unsigned char* __restrict__ elem_reg,
              const int* __restrict__ elem_sys, const double* __restrict__ node_xy,
              const int* __restrict__ node_row, const double* __restrict__ box, Grid g1,
              double* __restrict__ A1, Mat8 mat) {
    extern __shared__ double s_gal[];             // [54][RVE_GAL_THREADS]: gI(q) = entry (q * 9 + I) * 2 + component
    const int tid = threadIdx.x;
    const int e = blockIdx.x * RVE_GAL_THREADS + tid;
    if (e >= n_elem) return;
    double* sg = s_gal + tid;
#define SG(q, I, c) sg[(((q) * 9 + (I)) * 2 + (c)) * RVE_GAL_THREADS]
    const int sys = elem_sys[e];
    int row[6], ci[6], cj[6];
    double cs[6], ct[6];
    double vx[3], vy[3];
    const double bx0 = box[4 * sys], by0 = box[4 * sys + 1];
    const double invHx = g1.ncx / box[4 * sys + 2], invHy = g1.ncy / box[4 * sys + 3];
    int imin = 1 << 20, jmin = 1 << 20;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        const int node = elem_node[(size_t)e * 6 + a];
        row[a] = node_row[node];
        const double px = node_xy[2 * (size_t)node], py = node_xy[2 * (size_t)node + 1];
        if (a < 3) { vx[a] = px; vy[a] = py; }
        coarse_cell(px, py, bx0, by0, invHx, invHy, g1.ncx, g1.ncy, ci[a], cj[a], cs[a], ct[a]);
        imin = min(imin, ci[a]); jmin = min(jmin, cj[a]);
    }
    TriGeom g;
    tri_geom(vx[0], vy[0], vx[1], vy[1], vx[2], vy[2], g);
    const int reg = elem_reg[e];
    const double lam = mat.v[2 * reg], mu = mat.v[2 * reg + 1];
#pragma unroll
    for (int i = 0; i < 54; ++i) sg[i * RVE_GAL_THREADS] = 0.0;
    unsigned am = 0u;                             // window nodes that carry weight
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        if (row[a] < 0) continue;
        // an element larger than a coarse cell can reach a third cell: its far nodes are snapped onto the
        // last coarse-node line of the element's 3x3 window (the operator stays a sum of SPSD terms)
        int oi = ci[a] - imin, oj = cj[a] - jmin;
        double sa = cs[a], ta = ct[a];
        if (oi > 1) { oi = 1; sa = 1.0; }
        if (oj > 1) { oj = 1; ta = 1.0; }
        const int I0 = oj * 3 + oi;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const double wgt = ((c & 1) ? sa : 1.0 - sa) * ((c & 2) ? ta : 1.0 - ta);
            if (wgt == 0.0) continue;
            const int I = I0 + (c & 1) + 3 * (c >> 1);
            am |= 1u << I;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                double ax, ay;
                p2_grad(g, q, a, ax, ay);
                SG(q, I, 0) += wgt * ax; SG(q, I, 1) += wgt * ay;
            }
        }
    }
    double* Asys = A1 + (size_t)sys * 100 * g1.G;
    const double wq = g.area * (1.0 / 3.0);
    for (unsigned mi = am; mi; mi &= mi - 1) {
        const int I = __ffs(mi) - 1;
        const int Ii = imin + I % 3, Ij = jmin + I / 3;
        int nI;
        if (!grid_node(g1, Ii, Ij, nI)) continue;
        double gi[3][2];
#pragma unroll
        for (int q = 0; q < 3; ++q) { gi[q][0] = SG(q, I, 0); gi[q][1] = SG(q, I, 1); }
        for (unsigned mj = am & ~((1u << I) - 1u); mj; mj &= mj - 1) {       // upper half: J >= I in window order = (j, i) order
            const int J = __ffs(mj) - 1;
            const int Ji = imin + J % 3, Jj = jmin + J / 3;
            int nJ;
            if (!grid_node(g1, Ji, Jj, nJ)) continue;
            double G00 = 0, G01 = 0, G10 = 0, G11 = 0;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const double jx = SG(q, J, 0), jy = SG(q, J, 1);
                G00 += gi[q][0] * jx; G01 += gi[q][0] * jy; G10 += gi[q][1] * jx; G11 += gi[q][1] * jy;
            }
            G00 *= wq; G01 *= wq; G10 *= wq; G11 *= wq;
            const double tr = G00 + G11;
            const int slot = (Jj - Ij + 2) * 5 + (Ji - Ii + 2);
            double* d = Asys + (size_t)slot * 4 * g1.G + nI;
            atomicAdd(d, lam * G00 + mu * G00 + mu * tr); atomicAdd(d + g1.G, lam * G01 + mu * G10);
            atomicAdd(d + 2 * (size_t)g1.G, lam * G10 + mu * G01); atomicAdd(d + 3 * (size_t)g1.G, lam * G11 + mu * G11 + mu * tr);
        }
    }
#undef SG
}

// lower half of the level-1 stencils from the upper half: A[n][-d] = A[n - d][d]^T.  thread per (sys, lower slot, node)
__global__ void k_galerkin_mirror(int n_sys, Grid g, double* __restrict__ A) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)n_sys * 12 * g.G) return;
    const int idx = (int)(gid % g.G), slot = (int)((gid / g.G) % 12), sys = (int)(gid / ((long long)12 * g.G));
    const int di = slot % 5 - 2, dj = slot / 5 - 2;        // slots 0..11: dj < 0, or dj == 0 and di < 0
    const int i = idx % g.nx, j = idx / g.nx;
    int m;
    if (!grid_valid(g, i, j) || !grid_node(g, i + di, j + dj, m)) return;
    double* As = A + (size_t)sys * 100 * g.G;
    const double* src = As + (size_t)(24 - slot) * 4 * g.G + m;
    double* dst = As + (size_t)slot * 4 * g.G + idx;
    dst[0] = src[0]; dst[g.G] = src[2 * (size_t)g.G]; dst[2 * (size_t)g.G] = src[g.G]; dst[3 * (size_t)g.G] = src[3 * (size_t)g.G];
}

// level l -> l+1 Galerkin (stencil collapse with bilinear P), gather form: thread per (sys, I, slot)
__global__ void k_galerkin_coarse(int n_sys, Grid gf, Grid gc, const double* __restrict__ Af,
                                  double* __restrict__ Ac) {
    const int per_sys = gc.G * 25;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)n_sys * per_sys) return;
    const int sys = (int)(gid / per_sys);
    const int rem = (int)(gid % per_sys);
    const int slot = rem / gc.G, I = rem % gc.G;
    const int Ii = I % gc.nx, Ij = I / gc.nx;
    if (!grid_valid(gc, Ii, Ij)) return;
    const int Di = slot % 5 - 2, Dj = slot / 5 - 2;
    int nJ;
    if (!grid_node(gc, Ii + Di, Ij + Dj, nJ)) return;
    const double* A = Af + (size_t)sys * 100 * gf.G;
    double acc[4] = {0, 0, 0, 0};
    for (int aj = -1; aj <= 1; ++aj)
        for (int ai = -1; ai <= 1; ++ai) {
            int nf;
            if (!grid_node(gf, 2 * Ii + ai, 2 * Ij + aj, nf)) continue;
            const double wa = (ai ? 0.5 : 1.0) * (aj ? 0.5 : 1.0);
            for (int dj = -2; dj <= 2; ++dj) {
                const int bj = aj + dj - 2 * Dj;
                if (bj < -1 || bj > 1) continue;
                for (int di = -2; di <= 2; ++di) {
                    const int bi = ai + di - 2 * Di;
                    if (bi < -1 || bi > 1) continue;
                    int nfj;
                    if (!grid_node(gf, 2 * Ii + ai + di, 2 * Ij + aj + dj, nfj)) continue;
                    const double ww = wa * (bi ? 0.5 : 1.0) * (bj ? 0.5 : 1.0);
                    const double* a4 = A + (size_t)((dj + 2) * 5 + (di + 2)) * 4 * gf.G + nf;
                    acc[0] += ww * a4[0]; acc[1] += ww * a4[gf.G];
                    acc[2] += ww * a4[2 * (size_t)gf.G]; acc[3] += ww * a4[3 * (size_t)gf.G];
                }
            }
        }
    double* d = Ac + (size_t)sys * 100 * gc.G + (size_t)slot * 4 * gc.G + I;
    d[0] = acc[0]; d[gc.G] = acc[1]; d[2 * (size_t)gc.G] = acc[2]; d[3 * (size_t)gc.G] = acc[3];
}

// inverse diagonal blocks of a coarse level (sums the slots that alias the node itself under wrap)
__global__ void k_coarse_dinv(int n_sys, Grid g, const double* __restrict__ A, double* __restrict__ dinv) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)n_sys * g.G) return;
    const int sys = (int)(gid / g.G), I = (int)(gid % g.G);
    const int Ii = I % g.nx, Ij = I / g.nx;
    double d0 = 0, d1 = 0, d2 = 0, d3 = 0;
    if (grid_valid(g, Ii, Ij)) {
        const double* As = A + (size_t)sys * 100 * g.G;
        for (int dj = -2; dj <= 2; ++dj)
            for (int di = -2; di <= 2; ++di) {
                int nb;
                if (!grid_node(g, Ii + di, Ij + dj, nb) || nb != I) continue;
                const double* a4 = As + (size_t)((dj + 2) * 5 + (di + 2)) * 4 * g.G + I;
                d0 += a4[0]; d1 += a4[g.G]; d2 += a4[2 * (size_t)g.G]; d3 += a4[3 * (size_t)g.G];
            }
    }
    double det = d0 * d3 - d1 * d2;
    double inv = (det != 0.0) ? 1.0 / det : 0.0;
    double* o = dinv + ((size_t)sys * g.G + I) * 4;
    o[0] = d3 * inv; o[1] = -d1 * inv; o[2] = -d2 * inv; o[3] = d0 * inv;
}

// ---------------------------------------------------------------------------------------------
// (3) PCG kernels (one CTA of RVE_WIN threads per window)
// ---------------------------------------------------------------------------------------------

// q = A p, per-window partial of p.q.  CTA of RVE_SPMV_WARPS (= 4) warps per window.  The CTA first stages the p entries
// its window needs (own rows: one coalesced copy; halo rows: list-driven gather) in shared memory.
// Then every lane owns one row of a SELL-32 slice and warp w walks the slices w and w+4 of the window: rows are sorted
// by length inside a window (2 long vertex slices, 6 short mid-edge slices), warps 0-1 get a long and a short slice,
// warps 2-3 two short ones.  Per slot: one coalesced 2-byte
// column load, two coalesced LDG.128 of the 2x2 block and K LDS.128 gathers from the tile; the loads of
// the next slot pair (also across a slice boundary) are in flight while the current pair is multiplied.
// No cross-lane reduction; q leaves through shared memory so the global store is coalesced.
// (K = 3: the 48-byte pitch of the tile was suspected of the 38-43 % conflict wavefronts ncu shows; a 64-byte pitch was measured
// in round 2 and is WORSE — 67 % conflict wavefronts, +20 % time: with 4 x 16 B per row only two bank groups are reachable per
// rhs, while 3 is coprime with 8.  The conflicts are birthday collisions of 8 random rows per quarter warp, not a stride artefact:
// profiles/r02e_ncu_spmv_k3_pitch64_rejected_summary.csv.)
// warps per CTA (window): measured 2 -> 4: SpMV -4 % (K = 2) / -7 % (K = 3), the p-tile is shared by more warps and the
// shared-memory-limited occupancy doubles; 8 (one slice per warp: 2 long + 6 short slices) is 15-30 % slower
#ifndef RVE_SPMV_WARPS
#define RVE_SPMV_WARPS 4
#endif
#define RVE_SPMV_THREADS (32 * RVE_SPMV_WARPS)
template <int K>
__global__ void __launch_bounds__(RVE_SPMV_THREADS)
k_spmv(WinTab wt, Sell sl, const double2* __restrict__ bval, const double2* __restrict__ p,
       double2* __restrict__ q, double* __restrict__ part) {
    extern __shared__ double2 s_tile[];           // [(n + nhalo) * K] p-tile
    __shared__ unsigned s_g0[RVE_WIN / 32 + 1];
    const int win = blockIdx.x;
    if (!wt.win_active[win]) return;
    const int tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;
    const int row0 = wt.win_row0[win], n = wt.win_n[win];
    const int h0 = wt.win_halo0[win], nh = wt.win_nhalo[win];
    const int nsl = (n + 31) >> 5;
    if (tid <= nsl) s_g0[tid] = sl.slice_ptr[wt.win_slice0[win] + tid];
    {
        const double2* src = p + (size_t)row0 * K;
        for (int e = tid; e < n * K; e += RVE_SPMV_THREADS) s_tile[e] = src[e];
        double2* dst = s_tile + n * K;
        for (int e = tid; e < nh * K; e += RVE_SPMV_THREADS) {
            const int hr = sl.halo[h0 + e / K];
            dst[e] = p[(size_t)hr * K + (e % K)];
        }
    }
    __syncthreads();
    __shared__ double2 s_wq[RVE_SPMV_WARPS][32 * K];
    const uint16_t* cp = sl.scol + lane;
    const double2* vp = bval + lane;
    double dot[K];
    double2 y[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { dot[k] = 0.0; y[k] = make_double2(0.0, 0.0); }
    // current pair
    int s = wv, j = 0;
    int c0 = 0, c1 = 0;
    double2 a0 = make_double2(0.0, 0.0), a1 = a0, b0 = a0, b1 = a0;
    if (s < nsl) {
        const size_t g = s_g0[s];
        const int ng = (int)(s_g0[s + 1] - s_g0[s]);
        c0 = cp[g * 32]; a0 = vp[g * 64]; a1 = vp[g * 64 + 32];
        if (1 < ng) { c1 = cp[(g + 1) * 32]; b0 = vp[(g + 1) * 64]; b1 = vp[(g + 1) * 64 + 32]; }
    }
    while (s < nsl) {
        // position of the next pair (possibly in the warp's next slice) and its loads
        int ns = s, nj = j + 2;
        if (nj >= (int)(s_g0[s + 1] - s_g0[s])) { ns = s + RVE_SPMV_WARPS; nj = 0; }
        int nc0 = 0, nc1 = 0;
        double2 na0 = make_double2(0.0, 0.0), na1 = na0, nb0 = na0, nb1 = na0;
        if (ns < nsl) {
            const size_t g = (size_t)s_g0[ns] + nj;
            const int rem = (int)(s_g0[ns + 1] - s_g0[ns]) - nj;
            nc0 = cp[g * 32]; na0 = vp[g * 64]; na1 = vp[g * 64 + 32];
            if (1 < rem) { nc1 = cp[(g + 1) * 32]; nb0 = vp[(g + 1) * 64]; nb1 = vp[(g + 1) * 64 + 32]; }
        }
        // a missing second slot carries zero blocks against column 0 of the tile
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const double2 xa = s_tile[c0 * K + k], xb = s_tile[c1 * K + k];
            y[k].x += a0.x * xa.x + a0.y * xa.y; y[k].y += a1.x * xa.x + a1.y * xa.y;
            y[k].x += b0.x * xb.x + b0.y * xb.y; y[k].y += b1.x * xb.x + b1.y * xb.y;
        }
        if (ns != s) {                              // slice finished: rows s*32 + lane
            const int rl = s * 32 + lane;
            if (rl < n) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const double2 pv = s_tile[rl * K + k];
                    dot[k] += y[k].x * pv.x + y[k].y * pv.y;
                }
            }
            // the 32 rows of the slice leave through a per-warp staging line: coalesced 512-byte stores
            __syncwarp();
#pragma unroll
            for (int k = 0; k < K; ++k) s_wq[wv][lane * K + k] = y[k];
            __syncwarp();
            {
                double2* dst = q + ((size_t)row0 + s * 32) * K;
                const int lim = (min(n - s * 32, 32)) * K;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int e = lane + 32 * k;
                    if (e < lim) dst[e] = s_wq[wv][e];
                }
            }
#pragma unroll
            for (int k = 0; k < K; ++k) y[k] = make_double2(0.0, 0.0);
        }
        s = ns; j = nj; c0 = nc0; c1 = nc1; a0 = na0; a1 = na1; b0 = nb0; b1 = nb1;
    }
    __shared__ double s_dot[RVE_SPMV_WARPS][4];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const double v = warp_sum(dot[k]);
        if (lane == 0) s_dot[wv][k] = v;
    }
    __syncthreads();
    if (tid < K) {                                   // per-window partial of p.q; k_fin_alpha adds them per system
        double my = 0.0;
        for (int i = 0; i < RVE_SPMV_WARPS; ++i) my += s_dot[i][tid];
        part[(size_t)win * 4 + tid] = my;
    }
}

// x += alpha p ; r -= alpha q ; partial rr and r.Dinv r ; restriction of the new residual to the
// level-1 grid (rc += Z^T r, per-window partial sums then one RED per touched coarse value);
// convergence test by the last CTA of the system.  Thread = one (row, rhs) pair, flat and coalesced.
// elements (row, rhs) per thread of the vector kernels: measured best is 2 for K = 2 (256-thread CTAs) and
// 1 for K = 1, 3 (256 / 768-thread CTAs); all of a thread's loads are issued before the first use
#ifndef RVE_EPT_UPD
#define RVE_EPT_UPD(K) ((K) == 2 ? 2 : 1)
#endif
#ifndef RVE_EPT_DIR
#define RVE_EPT_DIR(K) ((K) == 2 ? 2 : 1)
#endif
#define RVE_UPD_THREADS(K) (RVE_WIN * (K) / RVE_EPT_UPD(K))
#define RVE_DIR_THREADS(K) (RVE_WIN * (K) / RVE_EPT_DIR(K))
#ifndef RVE_UPD_MINB
#define RVE_UPD_MINB(K) ((K) == 1 ? 5 : (RVE_UPD_THREADS(K) <= 256 ? 4 : 2))
#endif
#ifndef RVE_DIR_MINB
#define RVE_DIR_MINB(K) ((K) == 1 ? 5 : (RVE_DIR_THREADS(K) <= 256 ? 4 : 2))
#endif

template <int K>
__global__ void __launch_bounds__(RVE_UPD_THREADS(K), RVE_UPD_MINB(K))
k_update(WinTab wt, const double2* __restrict__ p, const double2* __restrict__ q, double2* __restrict__ x,
         double2* __restrict__ r, const float4* __restrict__ dinv32, const RowInfo* __restrict__ rowinfo,
         const int* __restrict__ cn_node, const unsigned* __restrict__ cn_beg, const uint16_t* __restrict__ cn_ent,
         float* __restrict__ rc1, int G, double* __restrict__ part, const double* __restrict__ sc, int init, int precond) {
    extern __shared__ double2 s_r[];              // [n*K] new residual, then float2 s_st[n], uint16 s_ent[4n]
    constexpr int EPT = RVE_EPT_UPD(K), TPB = RVE_UPD_THREADS(K);   // TPB % K == 0: a thread keeps its rhs index
    const int win = blockIdx.x;
    const int sys = wt.win_sys[win];
    if (!wt.win_active[win]) return;
    const int tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;
    const int row0 = wt.win_row0[win], n = wt.win_n[win];
    const int k = tid % K;
    float2* s_st = (float2*)(s_r + (size_t)RVE_WIN * K);
    uint16_t* s_ent = (uint16_t*)(s_st + RVE_WIN);
    const int cn0 = wt.win_cn0[win], ncn = wt.win_ncn[win];
    unsigned ent0 = 0, nent = 0;
    double rr = 0.0, rdr = 0.0;
    {
        const size_t base = (size_t)row0 * K;
        double2 rv[EPT], pv[EPT], qv[EPT], xv[EPT];
        float4 dv[EPT];
#pragma unroll
        for (int it = 0; it < EPT; ++it) {
            const int e = tid + it * TPB;
            if (e < n * K) {
                rv[it] = r[base + e];
                dv[it] = dinv32[row0 + e / K];
                if (!init) { pv[it] = p[base + e]; qv[it] = q[base + e]; xv[it] = x[base + e]; }
            }
        }
        if (precond) {
            ent0 = cn_beg[cn0]; nent = cn_beg[cn0 + ncn] - ent0;
            for (unsigned t = tid; t < nent; t += TPB) s_ent[t] = cn_ent[ent0 + t];
            for (int t = tid; t < n; t += TPB) { const RowInfo ri = rowinfo[row0 + t]; s_st[t] = make_float2(ri.s, ri.t); }
        }
        const double alpha = sc[((size_t)sys * 4 + k) * SC_N + SC_ALPHA];
#pragma unroll
        for (int it = 0; it < EPT; ++it) {
            const int e = tid + it * TPB;
            if (e < n * K) {
                double2 rw = rv[it];
                if (!init) {
                    double2 xw = xv[it];
                    xw.x += alpha * pv[it].x; xw.y += alpha * pv[it].y;
                    rw.x -= alpha * qv[it].x; rw.y -= alpha * qv[it].y;
                    x[base + e] = xw; r[base + e] = rw;
                }
                const float4 d = dv[it];
                rr += rw.x * rw.x + rw.y * rw.y;
                rdr += rw.x * ((double)d.x * rw.x + (double)d.y * rw.y) + rw.y * ((double)d.z * rw.x + (double)d.w * rw.y);
                s_r[e] = rw;
            }
        }
    }
    __shared__ double s_a[TPB], s_b[TPB];
    __shared__ double s_fin[8];
    s_a[tid] = rr; s_b[tid] = rdr;
    if (tid < 8) s_fin[tid] = 0.0;
    __syncthreads();
    // class sums: warp 2k+which adds the entries of rhs k
    if (wv < 2 * K) {
        const int kk = wv >> 1;
        const double* src = (wv & 1) ? s_b : s_a;
        double acc = 0.0;
        for (int m = lane; m < TPB / K; m += 32) acc += src[kk + K * m];
        acc = warp_sum(acc);
        if (lane == 0) s_fin[wv] = acc;
    }
    // restriction: 4 threads split the (row, corner) list of a coarse node of the window, weights from the staged
    // (s,t), vector REDs into the fp32 level-1 residual.  K = 2: both rhs per thread, one 16-byte RED per node;
    // otherwise one task per (node, rhs) (more parallelism for the 768-thread CTAs of K = 3), 8-byte REDs
    if (precond) {
        constexpr int KT = (K == 2) ? 2 : 1;              // rhs per task
        constexpr int TPN = K / KT;                       // tasks per node
        for (int t = tid; t < ((ncn * TPN * 4 + 31) & ~31); t += TPB) {
            const int task = t >> 2, sub = t & 3;
            const int jn = task / TPN, k0 = (task % TPN) * KT;
            double a[KT][2];
#pragma unroll
            for (int kk = 0; kk < KT; ++kk) { a[kk][0] = 0.0; a[kk][1] = 0.0; }
            if (jn < ncn) {
                const unsigned b0 = cn_beg[cn0 + jn] - ent0, b1 = cn_beg[cn0 + jn + 1] - ent0;
                for (unsigned e = b0 + sub; e < b1; e += 4) {
                    const int en = s_ent[e];
                    const int rl = en >> 2, c = en & 3;
                    const float2 st = s_st[rl];
                    const double wgt = (double)((c & 1) ? st.x : 1.0f - st.x) * (double)((c & 2) ? st.y : 1.0f - st.y);
#pragma unroll
                    for (int kk = 0; kk < KT; ++kk) {
                        const double2 v = s_r[rl * K + k0 + kk];
                        a[kk][0] += wgt * v.x; a[kk][1] += wgt * v.y;
                    }
                }
            }
#pragma unroll
            for (int kk = 0; kk < KT; ++kk)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    a[kk][c] += __shfl_xor_sync(0xffffffffu, a[kk][c], 1);
                    a[kk][c] += __shfl_xor_sync(0xffffffffu, a[kk][c], 2);
                }
            if (jn < ncn && sub == 0) {
                float* d = rc1 + (((size_t)sys * G + cn_node[cn0 + jn]) * K + k0) * 2;
                if (KT == 2) atomicAdd((float4*)d, make_float4((float)a[0][0], (float)a[0][1], (float)a[KT - 1][0], (float)a[KT - 1][1]));
                else atomicAdd((float2*)d, make_float2((float)a[0][0], (float)a[0][1]));
            }
        }
    }
    __syncthreads();
    if (tid < 8) part[(size_t)win * 8 + tid] = s_fin[tid];    // [rhs][rr, r.Dinv r]; k_fin_update adds them per system
}


// fp32 copies of a coarse level: A4 = A, H4_ij = A_ij (omega Dinv_j) (so that the residual after the
// first damped-Jacobi step, rc - A (omega Dinv rc), is ONE stencil pass over rc).  thread per (sys,slot,node)
__global__ void k_coarse_pack(int n_sys, Grid g, const double* __restrict__ A, const double* __restrict__ dinv,
                              stv* __restrict__ A4, stv* __restrict__ H4) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)n_sys * 25 * g.G) return;
    const int I = (int)(gid % g.G);
    const int slot = (int)((gid / g.G) % 25);
    const int sys = (int)(gid / ((long long)25 * g.G));
    const double* a = A + (size_t)sys * 100 * g.G + (size_t)slot * 4 * g.G + I;
    const double a00 = a[0], a01 = a[g.G], a10 = a[2 * (size_t)g.G], a11 = a[3 * (size_t)g.G];
    A4[gid] = st_pack((float)a00, (float)a01, (float)a10, (float)a11);
    int nb;
    double h00 = 0, h01 = 0, h10 = 0, h11 = 0;
    if (grid_node(g, I % g.nx + slot % 5 - 2, I / g.nx + slot / 5 - 2, nb)) {
        const double* d = dinv + ((size_t)sys * g.G + nb) * 4;
        const double om = RVE_OMEGA;
        h00 = om * (a00 * d[0] + a01 * d[2]); h01 = om * (a00 * d[1] + a01 * d[3]);
        h10 = om * (a10 * d[0] + a11 * d[2]); h11 = om * (a10 * d[1] + a11 * d[3]);
    }
    H4[gid] = st_pack((float)h00, (float)h01, (float)h10, (float)h11);
}

// Coarse V-cycle on the restricted residual (damped block-Jacobi V(1,1) on 25-point block stencils).
// out += S x over the 25 slots of node (i,j); branch-free: entries towards nodes outside the coarse
// space are zero by construction, so their x is read from the node itself.
template <int K>
__device__ __forceinline__ void stencil_apply(const Grid& g, const stv* __restrict__ S4,
                                              const cvec* __restrict__ x, int i, int j, cvec* out) {
    const int idx = j * g.nx + i;
#pragma unroll
    for (int dj = -2; dj <= 2; ++dj)
#pragma unroll
        for (int di = -2; di <= 2; ++di) {
            int nb;
            if (!grid_node_fast(g, i + di, j + dj, nb)) nb = idx;
            const float4 a = st_unpack(S4[(size_t)((dj + 2) * 5 + (di + 2)) * g.G + idx]);
            const cvec* xn = x + (size_t)nb * K;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const cvec xv = xn[k];
                out[k].x += a.x * xv.x + a.y * xv.y;
                out[k].y += a.z * xv.x + a.w * xv.y;
            }
        }
}

// Same sum for the grid-wide kernels of the large levels (thread per coarse node, nothing to amortise a latency with):
// the 25 stencil blocks of the node (the only DRAM stream, 200 B per node and pass, read once per V-cycle, hence
// evict-first loads) are ALL requested before the first one is used -- with the loop above the compiler keeps two
// slots in flight and the node pays about twelve dependent round trips; the neighbours' vectors (L1 / L2 hits) follow,
// both right-hand sides of K = 2 in one 16-byte load.
#ifndef RVE_VC_PRELOAD
#define RVE_VC_PRELOAD 1
#endif
#ifndef RVE_VC_MINB
#define RVE_VC_MINB 4              // CTAs of 256 threads per SM (64 registers, a few spilled words): measured against 3
#endif
// the 25 stencil blocks of node idx (evict-first: every block is read once per V-cycle)
__device__ __forceinline__ void stencil_load(const stv* __restrict__ S4, int G, int idx, stv* s) {
#pragma unroll
    for (int t = 0; t < 25; ++t) s[t] = __ldcs(S4 + (size_t)t * G + idx);
}
// out += S x with the stencil blocks in registers; x in global or shared memory (K = 2: 16-byte aligned)
template <int K>
__device__ __forceinline__ void stencil_apply_regs(const Grid& g, const stv* s, const cvec* x, int i, int j, cvec* out) {
    const int idx = j * g.nx + i;
#pragma unroll
    for (int dj = -2; dj <= 2; ++dj)
#pragma unroll
        for (int di = -2; di <= 2; ++di) {
            int nb;
            if (!grid_node_fast(g, i + di, j + dj, nb)) nb = idx;
            const float4 a = st_unpack(s[(dj + 2) * 5 + (di + 2)]);
            const cvec* xn = x + (size_t)nb * K;
            if (K == 2) {
                const float4 xv = *(const float4*)xn;
                out[0].x += a.x * xv.x + a.y * xv.y; out[0].y += a.z * xv.x + a.w * xv.y;
                out[K - 1].x += a.x * xv.z + a.y * xv.w; out[K - 1].y += a.z * xv.z + a.w * xv.w;
            } else {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const cvec xv = xn[k];
                    out[k].x += a.x * xv.x + a.y * xv.y;
                    out[k].y += a.z * xv.x + a.w * xv.y;
                }
            }
        }
}
template <int K>
__device__ __forceinline__ void stencil_apply_wide(const Grid& g, const stv* __restrict__ S4,
                                                   const cvec* __restrict__ x, int i, int j, cvec* out) {
#if RVE_VC_PRELOAD
    stv s[25];
    stencil_load(S4, g.G, j * g.nx + i, s);
    stencil_apply_regs<K>(g, s, x, i, j, out);
#else
    stencil_apply<K>(g, S4, x, i, j, out);
#endif
}

// wide level l, down: x_l = om Dinv rc_l ; t_l = rc_l - H rc_l.   thread per (sys, node), grid (ceil(G_l/256), n_sys)
template <int K>
__global__ void __launch_bounds__(256, RVE_VC_MINB)
k_vc_down1(Levels LV, int l, const int* __restrict__ active) {
    const int sys = blockIdx.y;
    if (!active[sys]) return;
    const Grid g = LV.g[l];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= g.G) return;
    const cvec* rc = (const cvec*)LV.rc[l] + (size_t)sys * g.G * K;
    cvec* xc = (cvec*)LV.xc[l] + (size_t)sys * g.G * K;
    cvec* tc = (cvec*)LV.tc[l] + (size_t)sys * g.G * K;
    const double* d = LV.dinv[l] + ((size_t)sys * g.G + idx) * 4;
    const float om = (float)RVE_OMEGA;
    const float d0 = om * (float)d[0], d1 = om * (float)d[1], d2 = om * (float)d[2], d3 = om * (float)d[3];
    cvec acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = make_float2(0.f, 0.f);
    stencil_apply_wide<K>(g, LV.H4[l] + (size_t)sys * 25 * g.G, rc, idx % g.nx, idx / g.nx, acc);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const cvec rv = rc[(size_t)idx * K + k];
        xc[(size_t)idx * K + k] = make_float2(d0 * rv.x + d1 * rv.y, d2 * rv.x + d3 * rv.y);
        tc[(size_t)idx * K + k] = make_float2(rv.x - acc[k].x, rv.y - acc[k].y);
    }
}

// levels >= l0 (the small ones): one CTA per system.  Restricts t_{l0-1}, runs the rest of the V-cycle down to
// the coarsest level and back up to level l0 (post-smoothed result in xc[l0]).
// (4 CTAs per SM: the kernel is bound by the latency of its phases, 48 registers with a few spilled words beat 68)
template <int K>
__global__ void __launch_bounds__(320, 4)
k_vc_mid(Levels LV, int l0, const int* __restrict__ active) {
    const int sys = blockIdx.x;
    if (!active[sys]) return;
    const int nth = blockDim.x, tid = threadIdx.x;
    const float om = (float)RVE_OMEGA;
    for (int l = l0; l < LV.L; ++l) {
        const Grid g = LV.g[l], gf = LV.g[l - 1];
        const double* Di = LV.dinv[l] + (size_t)sys * g.G * 4;
        cvec* rc = (cvec*)LV.rc[l] + (size_t)sys * g.G * K;
        cvec* xc = (cvec*)LV.xc[l] + (size_t)sys * g.G * K;
        cvec* tc = (cvec*)LV.tc[l] + (size_t)sys * g.G * K;
        const cvec* tf = (const cvec*)LV.tc[l - 1] + (size_t)sys * gf.G * K;
        // rc_l = P^T t_{l-1} (full weighting; t vanishes outside the coarse space)
        for (int idx = tid; idx < g.G; idx += nth) {
            const int I = idx % g.nx, J = idx / g.nx;
            cvec acc[K];
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] = make_float2(0.f, 0.f);
            if (grid_valid(g, I, J)) {
#pragma unroll
                for (int aj = -1; aj <= 1; ++aj)
#pragma unroll
                    for (int ai = -1; ai <= 1; ++ai) {
                        int nf;
                        if (!grid_node_fast(gf, 2 * I + ai, 2 * J + aj, nf)) continue;
                        const float wgt = (ai ? 0.5f : 1.0f) * (aj ? 0.5f : 1.0f);
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            const cvec tv = tf[(size_t)nf * K + k];
                            acc[k].x += wgt * tv.x; acc[k].y += wgt * tv.y;
                        }
                    }
            }
#pragma unroll
            for (int k = 0; k < K; ++k) rc[(size_t)idx * K + k] = acc[k];
        }
        __syncthreads();
        // x = om Dinv rc ; t = rc - H rc
        const stv* H4 = LV.H4[l] + (size_t)sys * 25 * g.G;
        for (int idx = tid; idx < g.G; idx += nth) {
            const double* d = Di + 4 * (size_t)idx;
            const float d0 = om * (float)d[0], d1 = om * (float)d[1], d2 = om * (float)d[2], d3 = om * (float)d[3];
            cvec acc[K];
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] = make_float2(0.f, 0.f);
            stencil_apply<K>(g, H4, rc, idx % g.nx, idx / g.nx, acc);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const cvec rv = rc[(size_t)idx * K + k];
                xc[(size_t)idx * K + k] = make_float2(d0 * rv.x + d1 * rv.y, d2 * rv.x + d3 * rv.y);
                tc[(size_t)idx * K + k] = make_float2(rv.x - acc[k].x, rv.y - acc[k].y);
            }
        }
        __syncthreads();
    }
    // coarsest level: two more damped-Jacobi sweeps (x += om Dinv t ; t = rc - A x)
    {
        const int l = LV.L - 1;
        const Grid g = LV.g[l];
        const double* Di = LV.dinv[l] + (size_t)sys * g.G * 4;
        const stv* A4 = LV.A4[l] + (size_t)sys * 25 * g.G;
        const cvec* rc = (const cvec*)LV.rc[l] + (size_t)sys * g.G * K;
        cvec* xc = (cvec*)LV.xc[l] + (size_t)sys * g.G * K;
        cvec* tc = (cvec*)LV.tc[l] + (size_t)sys * g.G * K;
        for (int sw = 0; sw < 2; ++sw) {
            for (int idx = tid; idx < g.G; idx += nth) {
                const double* d = Di + 4 * (size_t)idx;
                const float d0 = om * (float)d[0], d1 = om * (float)d[1], d2 = om * (float)d[2], d3 = om * (float)d[3];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const cvec tv = tc[(size_t)idx * K + k];
                    cvec xv = xc[(size_t)idx * K + k];
                    xv.x += d0 * tv.x + d1 * tv.y; xv.y += d2 * tv.x + d3 * tv.y;
                    xc[(size_t)idx * K + k] = xv;
                }
            }
            __syncthreads();
            if (sw == 1) break;
            for (int idx = tid; idx < g.G; idx += nth) {
                cvec acc[K];
#pragma unroll
                for (int k = 0; k < K; ++k) acc[k] = make_float2(0.f, 0.f);
                stencil_apply<K>(g, A4, xc, idx % g.nx, idx / g.nx, acc);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const cvec rv = rc[(size_t)idx * K + k];
                    tc[(size_t)idx * K + k] = make_float2(rv.x - acc[k].x, rv.y - acc[k].y);
                }
            }
            __syncthreads();
        }
    }
    // up sweep: x_l += P x_{l+1}, then post-smoothing, for the levels this kernel owns
    for (int l = LV.L - 2; l >= l0; --l) {
        const Grid g = LV.g[l], gc = LV.g[l + 1];
        cvec* xc = (cvec*)LV.xc[l] + (size_t)sys * g.G * K;
        const cvec* xcc = (const cvec*)LV.xc[l + 1] + (size_t)sys * gc.G * K;
        for (int idx = tid; idx < g.G; idx += nth) {
            const int i = idx % g.nx, j = idx / g.nx;
            if (!grid_valid(g, i, j)) continue;
            const int I0 = i >> 1, J0 = j >> 1;
            const int ni = (i & 1) ? 2 : 1, nj = (j & 1) ? 2 : 1;
            const float wgt = ((i & 1) ? 0.5f : 1.0f) * ((j & 1) ? 0.5f : 1.0f);
            cvec acc[K];
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] = xc[(size_t)idx * K + k];
            for (int bj = 0; bj < nj; ++bj)
                for (int bi = 0; bi < ni; ++bi) {
                    int nc;
                    if (!grid_node_fast(gc, I0 + bi, J0 + bj, nc)) continue;
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const cvec v = xcc[(size_t)nc * K + k];
                        acc[k].x += wgt * v.x; acc[k].y += wgt * v.y;
                    }
                }
#pragma unroll
            for (int k = 0; k < K; ++k) xc[(size_t)idx * K + k] = acc[k];
        }
        __syncthreads();
        const double* Di = LV.dinv[l] + (size_t)sys * g.G * 4;
        const stv* A4 = LV.A4[l] + (size_t)sys * 25 * g.G;
        const cvec* rc = (const cvec*)LV.rc[l] + (size_t)sys * g.G * K;
        cvec* tc = (cvec*)LV.tc[l] + (size_t)sys * g.G * K;
        for (int idx = tid; idx < g.G; idx += nth) {
            cvec acc[K];
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] = make_float2(0.f, 0.f);
            stencil_apply<K>(g, A4, xc, idx % g.nx, idx / g.nx, acc);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const cvec rv = rc[(size_t)idx * K + k];
                tc[(size_t)idx * K + k] = make_float2(rv.x - acc[k].x, rv.y - acc[k].y);
            }
        }
        __syncthreads();
        for (int idx = tid; idx < g.G; idx += nth) {
            const double* d = Di + 4 * (size_t)idx;
            const float d0 = om * (float)d[0], d1 = om * (float)d[1], d2 = om * (float)d[2], d3 = om * (float)d[3];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const cvec tv = tc[(size_t)idx * K + k];
                cvec xv = xc[(size_t)idx * K + k];
                xv.x += d0 * tv.x + d1 * tv.y; xv.y += d2 * tv.x + d3 * tv.y;
                xc[(size_t)idx * K + k] = xv;
            }
        }
        __syncthreads();
    }
}

// The same V-cycle tail with every vector of the levels >= l0 in SHARED memory (rc, xc, tc per level: 3 K G_l float2),
// so that the ~20 barrier-separated phases of a system no longer pay a global round trip each: the only global
// accesses left are the restriction input t_{l0-1}, the stencil blocks + Dinv of each pass -- requested BEFORE the
// barrier-separated phase that precedes their use -- and the result xc[l0].  Residual and damped-Jacobi update of the
// post-smoothing are done by the node's own thread (one barrier in between: the neighbours read x).
// Bitwise the same arithmetic as k_vc_mid.  Dynamic shared memory: vc_mid_smem(LV, l0, K).
#ifndef RVE_VC_MID_SMEM
#define RVE_VC_MID_SMEM 1
#endif
#ifndef RVE_VC_MID_MINB
#define RVE_VC_MID_MINB(K) ((K) == 3 ? 3 : 2)      // measured: K = 3 prefers three CTAs per SM with a few spilled words
#endif
__host__ inline size_t vc_mid_smem(const Levels& LV, int l0, int K) {
    size_t n = 0;
    for (int l = l0; l < LV.L; ++l) n += (size_t)LV.g[l].G;
    return n * 3 * K * sizeof(cvec) + 16;
}
template <int K>
__global__ void __launch_bounds__(320, RVE_VC_MID_MINB(K))
k_vc_mid_s(Levels LV, int l0, const int* __restrict__ active) {
    extern __shared__ __align__(16) unsigned char s_mid_raw[];
    cvec* s_v = (cvec*)s_mid_raw;
    const int sys = blockIdx.x;
    if (!active[sys]) return;
    const int nth = blockDim.x, tid = threadIdx.x;
    const float om = (float)RVE_OMEGA;
    stv s[25];
    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
    auto load_node = [&](const stv* __restrict__ S4, const double* __restrict__ Di, int G, int idx) {
        stencil_load(S4, G, idx, s);
        const double2 da = *(const double2*)(Di + 4 * (size_t)idx), db = *(const double2*)(Di + 4 * (size_t)idx + 2);
        d0 = om * (float)da.x; d1 = om * (float)da.y; d2 = om * (float)db.x; d3 = om * (float)db.y;
    };
    int off = 0, off_f = 0;                          // float2 offsets of level l and of level l-1 in s_v
    // ---- down: rc_l = P^T t_{l-1} ; x_l = om Dinv rc_l ; t_l = rc_l - H rc_l
    for (int l = l0; l < LV.L; ++l) {
        const Grid g = LV.g[l], gf = LV.g[l - 1];
        cvec* rc = s_v + off; cvec* xc = rc + g.G * K; cvec* tc = xc + g.G * K;
        const stv* H4 = LV.H4[l] + (size_t)sys * 25 * g.G;
        const double* Di = LV.dinv[l] + (size_t)sys * g.G * 4;
        const bool single = g.G <= nth;              // one node per thread: its stencil is requested a phase ahead
        if (single && tid < g.G) load_node(H4, Di, g.G, tid);
        const cvec* tf = (l == l0) ? (const cvec*)LV.tc[l - 1] + (size_t)sys * gf.G * K : s_v + off_f + 2 * gf.G * K;
        for (int idx = tid; idx < g.G; idx += nth) {
            const int I = idx % g.nx, J = idx / g.nx;
            cvec acc[K];
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] = make_float2(0.f, 0.f);
            if (grid_valid(g, I, J)) {
#pragma unroll
                for (int aj = -1; aj <= 1; ++aj)
#pragma unroll
                    for (int ai = -1; ai <= 1; ++ai) {
                        int nf;
                        if (!grid_node_fast(gf, 2 * I + ai, 2 * J + aj, nf)) continue;
                        const float wgt = (ai ? 0.5f : 1.0f) * (aj ? 0.5f : 1.0f);
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            const cvec tv = tf[(size_t)nf * K + k];
                            acc[k].x += wgt * tv.x; acc[k].y += wgt * tv.y;
                        }
                    }
            }
#pragma unroll
            for (int k = 0; k < K; ++k) rc[idx * K + k] = acc[k];
        }
        __syncthreads();
        for (int idx = tid; idx < g.G; idx += nth) {
            if (!single) load_node(H4, Di, g.G, idx);
            cvec acc[K];
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] = make_float2(0.f, 0.f);
            stencil_apply_regs<K>(g, s, rc, idx % g.nx, idx / g.nx, acc);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const cvec rv = rc[idx * K + k];
                xc[idx * K + k] = make_float2(d0 * rv.x + d1 * rv.y, d2 * rv.x + d3 * rv.y);
                tc[idx * K + k] = make_float2(rv.x - acc[k].x, rv.y - acc[k].y);
            }
        }
        if (l + 1 < LV.L) { off_f = off; off += 3 * g.G * K; }
        else {
            // ---- coarsest level: two more damped-Jacobi sweeps.  x += om Dinv t (own node) ; t = rc - A x ; x += om Dinv t
            const stv* A4 = LV.A4[l] + (size_t)sys * 25 * g.G;
            for (int idx = tid; idx < g.G; idx += nth) {
                if (!single) load_node(A4, Di, g.G, idx);           // (for Dinv; never taken in practice: the coarsest level is tiny)
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const cvec tv = tc[idx * K + k];
                    cvec xv = xc[idx * K + k];
                    xv.x += d0 * tv.x + d1 * tv.y; xv.y += d2 * tv.x + d3 * tv.y;
                    xc[idx * K + k] = xv;
                }
            }
            if (single && tid < g.G) load_node(A4, Di, g.G, tid);
            __syncthreads();
            for (int idx = tid; idx < g.G; idx += nth) {
                if (!single) load_node(A4, Di, g.G, idx);
                cvec acc[K];
#pragma unroll
                for (int k = 0; k < K; ++k) acc[k] = make_float2(0.f, 0.f);
                stencil_apply_regs<K>(g, s, xc, idx % g.nx, idx / g.nx, acc);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const cvec rv = rc[idx * K + k];
                    const float t0 = rv.x - acc[k].x, t1 = rv.y - acc[k].y;
                    tc[idx * K + k] = make_float2(d0 * t0 + d1 * t1, d2 * t0 + d3 * t1);     // the update, applied after the barrier
                }
            }
            __syncthreads();
            for (int idx = tid; idx < g.G; idx += nth)
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const cvec uv = tc[idx * K + k];
                    cvec xv = xc[idx * K + k];
                    xv.x += uv.x; xv.y += uv.y;
                    xc[idx * K + k] = xv;
                }
        }
        __syncthreads();
    }
    // ---- up: x_l += P x_{l+1} ; x_l += om Dinv (rc_l - A x_l)
    for (int l = LV.L - 2; l >= l0; --l) {
        const Grid g = LV.g[l], gc = LV.g[l + 1];
        const int off_c = off;                        // level l+1
        off -= 3 * g.G * K;                           // level l
        cvec* rc = s_v + off; cvec* xc = rc + g.G * K; cvec* tc = xc + g.G * K;
        const cvec* xcc = s_v + off_c + gc.G * K;
        const stv* A4 = LV.A4[l] + (size_t)sys * 25 * g.G;
        const double* Di = LV.dinv[l] + (size_t)sys * g.G * 4;
        const bool single = g.G <= nth;
        if (single && tid < g.G) load_node(A4, Di, g.G, tid);
        for (int idx = tid; idx < g.G; idx += nth) {
            const int i = idx % g.nx, j = idx / g.nx;
            if (!grid_valid(g, i, j)) continue;
            const int I0 = i >> 1, J0 = j >> 1;
            const int ni = (i & 1) ? 2 : 1, nj = (j & 1) ? 2 : 1;
            const float wgt = ((i & 1) ? 0.5f : 1.0f) * ((j & 1) ? 0.5f : 1.0f);
            cvec acc[K];
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] = xc[idx * K + k];
            for (int bj = 0; bj < nj; ++bj)
                for (int bi = 0; bi < ni; ++bi) {
                    int nc;
                    if (!grid_node_fast(gc, I0 + bi, J0 + bj, nc)) continue;
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const cvec v = xcc[nc * K + k];
                        acc[k].x += wgt * v.x; acc[k].y += wgt * v.y;
                    }
                }
#pragma unroll
            for (int k = 0; k < K; ++k) xc[idx * K + k] = acc[k];
        }
        __syncthreads();
        for (int idx = tid; idx < g.G; idx += nth) {
            if (!single) load_node(A4, Di, g.G, idx);
            cvec acc[K];
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] = make_float2(0.f, 0.f);
            stencil_apply_regs<K>(g, s, xc, idx % g.nx, idx / g.nx, acc);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const cvec rv = rc[idx * K + k];
                const float t0 = rv.x - acc[k].x, t1 = rv.y - acc[k].y;
                tc[idx * K + k] = make_float2(d0 * t0 + d1 * t1, d2 * t0 + d3 * t1);
            }
        }
        __syncthreads();
        cvec* xg = (cvec*)LV.xc[l] + (size_t)sys * g.G * K;
        for (int idx = tid; idx < g.G; idx += nth)
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const cvec uv = tc[idx * K + k];
                cvec xv = xc[idx * K + k];
                xv.x += uv.x; xv.y += uv.y;
                if (l == l0) xg[(size_t)idx * K + k] = xv; else xc[idx * K + k] = xv;
            }
        if (l > l0) __syncthreads();
    }
    if (LV.L - 1 == l0) {                             // a single small level: its result is the coarsest sweep's
        const Grid g = LV.g[l0];
        const cvec* xc = s_v + off + g.G * K;
        cvec* xg = (cvec*)LV.xc[l0] + (size_t)sys * g.G * K;
        for (int idx = tid; idx < g.G * K; idx += nth) xg[idx] = xc[idx];
    }
}

// Large coarse levels (>= RVE_WIDE_NODES nodes; always level 1) run as grid-wide kernels, thread per (system,
// node), grid (ceil(G_l/256), n_sys): restriction, pre-smoothing + residual, prolongation, post-smoothing.
#ifndef RVE_WIDE_NODES
#define RVE_WIDE_NODES 1024
#endif

// rc_l = P^T t_{l-1} (full weighting; t vanishes outside the coarse space)
template <int K>
__global__ void __launch_bounds__(256)
k_vc_restrict(Levels LV, int l, const int* __restrict__ active) {
    const int sys = blockIdx.y;
    if (!active[sys]) return;
    const Grid g = LV.g[l], gf = LV.g[l - 1];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= g.G) return;
    const cvec* tf = (const cvec*)LV.tc[l - 1] + (size_t)sys * gf.G * K;
    cvec* rc = (cvec*)LV.rc[l] + (size_t)sys * g.G * K;
    const int I = idx % g.nx, J = idx / g.nx;
    cvec acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = make_float2(0.f, 0.f);
    if (grid_valid(g, I, J)) {
#pragma unroll
        for (int aj = -1; aj <= 1; ++aj)
#pragma unroll
            for (int ai = -1; ai <= 1; ++ai) {
                int nf;
                if (!grid_node_fast(gf, 2 * I + ai, 2 * J + aj, nf)) continue;
                const float wgt = (ai ? 0.5f : 1.0f) * (aj ? 0.5f : 1.0f);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const cvec tv = tf[(size_t)nf * K + k];
                    acc[k].x += wgt * tv.x; acc[k].y += wgt * tv.y;
                }
            }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) rc[(size_t)idx * K + k] = acc[k];
}

// x_l += P src_{l+1}   (src = the post-smoothed correction of the next coarser level)
template <int K>
__global__ void __launch_bounds__(256)
k_vc_prolong(Levels LV, int l, const float* __restrict__ src, const int* __restrict__ active) {
    const int sys = blockIdx.y;
    if (!active[sys]) return;
    const Grid g = LV.g[l], gc = LV.g[l + 1];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= g.G) return;
    const int i = idx % g.nx, j = idx / g.nx;
    if (!grid_valid(g, i, j)) return;
    cvec* xc = (cvec*)LV.xc[l] + (size_t)sys * g.G * K;
    const cvec* xcc = (const cvec*)src + (size_t)sys * gc.G * K;
    const int I0 = i >> 1, J0 = j >> 1;
    const int ni = (i & 1) ? 2 : 1, nj = (j & 1) ? 2 : 1;
    const float wgt = ((i & 1) ? 0.5f : 1.0f) * ((j & 1) ? 0.5f : 1.0f);
    cvec acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = xc[(size_t)idx * K + k];
    for (int bj = 0; bj < nj; ++bj)
        for (int bi = 0; bi < ni; ++bi) {
            int nc;
            if (!grid_node_fast(gc, I0 + bi, J0 + bj, nc)) continue;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const cvec v = xcc[(size_t)nc * K + k];
                acc[k].x += wgt * v.x; acc[k].y += wgt * v.y;
            }
        }
#pragma unroll
    for (int k = 0; k < K; ++k) xc[(size_t)idx * K + k] = acc[k];
}

// post-smoothing of a wide level l >= 1: tc_l = x_l + om Dinv (rc_l - A x_l)
template <int K>
__global__ void __launch_bounds__(256, RVE_VC_MINB)
k_vc_upl(Levels LV, int l, const int* __restrict__ active) {
    const int sys = blockIdx.y;
    if (!active[sys]) return;
    const Grid g = LV.g[l];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= g.G) return;
    const cvec* rc = (const cvec*)LV.rc[l] + (size_t)sys * g.G * K;
    const cvec* xc = (const cvec*)LV.xc[l] + (size_t)sys * g.G * K;
    cvec* zc = (cvec*)LV.tc[l] + (size_t)sys * g.G * K;
    const double* d = LV.dinv[l] + ((size_t)sys * g.G + idx) * 4;
    const float om = (float)RVE_OMEGA;
    const float d0 = om * (float)d[0], d1 = om * (float)d[1], d2 = om * (float)d[2], d3 = om * (float)d[3];
    cvec acc[K], rv[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { acc[k] = make_float2(0.f, 0.f); rv[k] = rc[(size_t)idx * K + k]; }    // (requested with the stencil, not after it)
    stencil_apply_wide<K>(g, LV.A4[l] + (size_t)sys * 25 * g.G, xc, idx % g.nx, idx / g.nx, acc);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float t0 = rv[k].x - acc[k].x, t1 = rv[k].y - acc[k].y;
        cvec xv = xc[(size_t)idx * K + k];
        xv.x += d0 * t0 + d1 * t1; xv.y += d2 * t0 + d3 * t1;
        zc[(size_t)idx * K + k] = xv;
    }
}

// level-1 up: z1 = x1 + om Dinv (rc - A x1) written to tc[0] (k_direction prolongs it), coarse part of
// r.z = rc . z1, beta; clears rc for the next restriction.  grid (ceil(G/256), n_sys)
template <int K>
__global__ void __launch_bounds__(256, RVE_VC_MINB)
k_vc_up1(Levels LV, const int* __restrict__ active, double* __restrict__ part) {   // block size 128 or 256
    const int sys = blockIdx.y;
    if (!active[sys]) return;
    const Grid g = LV.g[0];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
    double dot[K];
#pragma unroll
    for (int k = 0; k < K; ++k) dot[k] = 0.0;
    if (idx < g.G) {
        cvec* rc = (cvec*)LV.rc[0] + (size_t)sys * g.G * K;
        const cvec* xc = (const cvec*)LV.xc[0] + (size_t)sys * g.G * K;
        cvec* zc = (cvec*)LV.tc[0] + (size_t)sys * g.G * K;
        const double* d = LV.dinv[0] + ((size_t)sys * g.G + idx) * 4;
        const float om = (float)RVE_OMEGA;
        const float d0 = om * (float)d[0], d1 = om * (float)d[1], d2 = om * (float)d[2], d3 = om * (float)d[3];
        cvec acc[K], rv[K];
#pragma unroll
        for (int k = 0; k < K; ++k) { acc[k] = make_float2(0.f, 0.f); rv[k] = rc[(size_t)idx * K + k]; }   // (requested with the stencil, not after it)
        stencil_apply_wide<K>(g, LV.A4[0] + (size_t)sys * 25 * g.G, xc, idx % g.nx, idx / g.nx, acc);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float t0 = rv[k].x - acc[k].x, t1 = rv[k].y - acc[k].y;
            cvec xv = xc[(size_t)idx * K + k];
            xv.x += d0 * t0 + d1 * t1; xv.y += d2 * t0 + d3 * t1;
            zc[(size_t)idx * K + k] = xv;
            dot[k] = (double)rv[k].x * (double)xv.x + (double)rv[k].y * (double)xv.y;
            rc[(size_t)idx * K + k] = make_float2(0.f, 0.f);
        }
    }
    // per-WARP partial of rc.z1 (no block barrier: finished warps retire); k_fin_beta adds them per system in a fixed order
    const int nwp = blockDim.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const double v = warp_sum(dot[k]);
        if (lane == 0) part[(((size_t)sys * gridDim.x + blockIdx.x) * nwp + wv) * 4 + k] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// per-system scalar recurrences of the PCG, one warp per system.  The kernels above leave per-window (per-CTA)
// partial sums; these add them in a fixed order (lane-strided, then a butterfly), so the scalars are reproducible
// run to run, and the streaming kernels carry no fence / atomic / "last CTA" tail.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double sys_sum(const double* __restrict__ v, int first, int last, int stride, int off) {
    double acc = 0.0;
    for (int c = first + (int)(threadIdx.x & 31); c < last; c += 32) acc += v[(size_t)c * stride + off];
    return warp_sum(acc);
}

// alpha = r.z / p.q after the operator application
__global__ void __launch_bounds__(256)
k_fin_alpha(int n_sys, int K, const int* __restrict__ sys_win0, const double* __restrict__ part, const int* __restrict__ active,
            double* __restrict__ sc, double* __restrict__ win_scal) {
    const int sys = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (sys >= n_sys || !active[sys]) return;
    const int lane = threadIdx.x & 31;
    const int w0 = sys_win0[sys], w1 = sys_win0[sys + 1];
    for (int k = 0; k < K; ++k) {
        const double pq = sys_sum(part, w0, w1, 4, k);
        double* s_ = sc + ((size_t)sys * 4 + k) * SC_N;
        const double alpha = (pq > 0.0) ? s_[SC_RZ] / pq : 0.0;
        if (lane == 0) { s_[SC_PQ] = pq; s_[SC_ALPHA] = alpha; }
        // per-window copy ([window][8]: alpha[0..3], beta[4..7]): the persistent kernels read it next to the descriptor
        for (int w = w0 + lane; w < w1; w += 32) win_scal[(size_t)w * 8 + k] = alpha;
    }
}

// r.r, r.Dinv r after the update; convergence test; a converged system switches itself and its windows off
__global__ void __launch_bounds__(256)
k_fin_update(int n_sys, int K, const int* __restrict__ sys_win0, const double* __restrict__ part, int* __restrict__ active,
             unsigned char* __restrict__ win_active, double* __restrict__ sc, int init, int iter, double rtol2, int precond,
             int* __restrict__ iters, int* n_active, double* __restrict__ win_scal) {
    const int sys = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (sys >= n_sys || !active[sys]) return;
    const int lane = threadIdx.x & 31;
    const int w0 = sys_win0[sys], w1 = sys_win0[sys + 1];
    bool done = true;
    for (int k = 0; k < K; ++k) {
        const double vrr = sys_sum(part, w0, w1, 8, 2 * k), vrdr = sys_sum(part, w0, w1, 8, 2 * k + 1);
        double* s = sc + ((size_t)sys * 4 + k) * SC_N;
        // init: 0 = after an update, 1 = first residual (sets the reference norm), 2 = after a residual replacement
        const double bb = (init == 1) ? vrr : s[SC_BB];
        if (!(vrr <= rtol2 * bb || vrr <= 1e-28 * s[SC_BREF])) done = false;
        if (precond == 0) {  // block Jacobi: z = Dinv r, finish the scalar recurrences here
            const double rzold = (init == 2) ? s[SC_RZOLD] : s[SC_RZ];
            const double beta = (init != 1 && rzold > 0.0) ? vrdr / rzold : 0.0;
            __syncwarp();
            if (lane == 0) { s[SC_RZOLD] = rzold; s[SC_RZ] = vrdr; s[SC_BETA] = beta; }
            for (int w = w0 + lane; w < w1; w += 32) win_scal[(size_t)w * 8 + 4 + k] = beta;
        }
        if (lane == 0) {
            if (init == 1) s[SC_BB] = vrr;
            s[SC_RR] = vrr; s[SC_RDR] = vrdr;
        }
    }
    if (!done) return;
    if (lane == 0) {
        active[sys] = 0;
        iters[sys] = (init == 1) ? 0 : iter + 1;
        atomicSub(n_active, 1);
    }
    for (int wq = w0 + lane; wq < w1; wq += 32) win_active[wq] = 0;
}

// residual replacement: r = b - q with q = A x just applied (thread per (row, rhs) value pair of the active windows)
__global__ void __launch_bounds__(256)
k_true_residual(WinTab wt, int K, const double2* __restrict__ b, const double2* __restrict__ q, double2* __restrict__ r) {
    const int w = blockIdx.x;
    if (!wt.win_active[w]) return;
    const size_t base = (size_t)wt.win_row0[w] * K;
    for (int e = threadIdx.x; e < wt.win_n[w] * K; e += 256) {
        const double2 bv = b[base + e], qv = q[base + e];
        r[base + e] = make_double2(bv.x - qv.x, bv.y - qv.y);
    }
}

// r.z = r.Dinv r + rc.z1 after the V-cycle; beta
__global__ void __launch_bounds__(256)
k_fin_beta(int n_sys, int K, int nb, const int* __restrict__ sys_win0, const double* __restrict__ cpart,
           const int* __restrict__ active, double* __restrict__ sc, int first, double* __restrict__ win_scal) {
    const int sys = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (sys >= n_sys || !active[sys]) return;
    const int lane = threadIdx.x & 31;
    const int w0 = sys_win0[sys], w1 = sys_win0[sys + 1];
    for (int k = 0; k < K; ++k) {
        const double cz = sys_sum(cpart, sys * nb, sys * nb + nb, 4, k);
        double* s = sc + ((size_t)sys * 4 + k) * SC_N;
        const double rz = s[SC_RDR] + cz;
        const double rzold = s[SC_RZ];
        const double beta = (!first && rzold > 0.0) ? rz / rzold : 0.0;
        __syncwarp();
        if (lane == 0) { s[SC_RZOLD] = rzold; s[SC_RZ] = rz; s[SC_BETA] = beta; }
        for (int w = w0 + lane; w < w1; w += 32) win_scal[(size_t)w * 8 + 4 + k] = beta;
    }
}

// p = Dinv r + Z xc1 + beta p     (thread = one (row, rhs) pair; the level-1 values the window
// touches are staged in shared memory)
template <int K>
__global__ void __launch_bounds__(RVE_DIR_THREADS(K), RVE_DIR_MINB(K))
k_direction(WinTab wt, double2* __restrict__ p, const double2* __restrict__ r, const float4* __restrict__ dinv32,
            const RowInfo* __restrict__ rowinfo, const int* __restrict__ cn_node, const cvec* __restrict__ xc1,
            int G, const int* __restrict__ active, const double* __restrict__ sc, int precond) {
    extern __shared__ double2 s_xc[];             // [ncn*K]
    constexpr int EPT = RVE_EPT_DIR(K), TPB = RVE_DIR_THREADS(K);
    const int win = blockIdx.x;
    const int sys = wt.win_sys[win];
    if (!wt.win_active[win]) return;
    const int tid = threadIdx.x;
    const int row0 = wt.win_row0[win], n = wt.win_n[win];
    const int k = tid % K;
    const size_t base = (size_t)row0 * K;
    // the streaming loads do not depend on the staged coarse values: request them first
    double2 rv[EPT], pv[EPT];
    float4 dv[EPT];
    RowInfo ri[EPT];
#pragma unroll
    for (int it = 0; it < EPT; ++it) {
        const int e = tid + it * TPB;
        if (e < n * K) {
            rv[it] = r[base + e]; pv[it] = p[base + e];
            dv[it] = dinv32[row0 + e / K];
            if (precond) ri[it] = rowinfo[row0 + e / K];
        }
    }
    const double beta = sc[((size_t)sys * 4 + k) * SC_N + SC_BETA];
    if (precond) {
        const int cn0 = wt.win_cn0[win], ncn = wt.win_ncn[win];
        for (int t = tid; t < ncn * K; t += TPB)
        {
            const cvec v = xc1[((size_t)sys * G + cn_node[cn0 + t / K]) * K + (t % K)];
            s_xc[t] = make_double2((double)v.x, (double)v.y);
        }
        __syncthreads();
    }
#pragma unroll
    for (int it = 0; it < EPT; ++it) {
        const int e = tid + it * TPB;
        if (e < n * K) {
            const float4 d = dv[it];
            double z0 = (double)d.x * rv[it].x + (double)d.y * rv[it].y, z1 = (double)d.z * rv[it].x + (double)d.w * rv[it].y;
            if (precond) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int slot = ri[it].slot[c];
                    if (slot != 0xffff) {
                        const double wgt = (double)((c & 1) ? ri[it].s : 1.0f - ri[it].s) * (double)((c & 2) ? ri[it].t : 1.0f - ri[it].t);
                        const double2 v = s_xc[slot * K + k];
                        z0 += wgt * v.x; z1 += wgt * v.y;
                    }
                }
            }
            p[base + e] = make_double2(z0 + beta * pv[it].x, z1 + beta * pv[it].y);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// (4) right-hand sides
// ---------------------------------------------------------------------------------------------

// dnn pre-step (micro_model_dnn.py:85-94): B = -(1/vol) oint uD (x) n ds (exact trapezoid on the
// P1 boundary of Vref), eps_eff = eps - voigt(B); thread per (sys, rhs).
__global__ void k_dnn_prep(int n_sys, int K, int nside, const double* __restrict__ vbox /*[4]*/,
                           const double* __restrict__ uD /*[sys][K][2*(4*nside+1)] or null*/,
                           const double* __restrict__ eps, double* __restrict__ eps_eff,
                           double* __restrict__ Bmat /*[sys][K][4]*/, int sym_B) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n_sys * K) return;
    double B00 = 0, B01 = 0, B10 = 0, B11 = 0;
    if (uD) {
        const int nb = 4 * nside;
        const double* u = uD + (size_t)gid * 2 * (nb + 1);
        const double hx = vbox[2] / nside, hy = vbox[3] / nside;
        for (int k = 0; k < nb; ++k) {
            const int k2 = (k + 1) % nb, side = k / nside;
            // CCW boundary: outward normals (0,-1), (1,0), (0,1), (-1,0); segment lengths hx,hy,hx,hy
            const double nx = (side == 1) ? 1.0 : (side == 3 ? -1.0 : 0.0);
            const double ny = (side == 0) ? -1.0 : (side == 2 ? 1.0 : 0.0);
            const double len = (side & 1) ? hy : hx;
            const double m0 = 0.5 * len * (u[2 * k] + u[2 * k2]), m1 = 0.5 * len * (u[2 * k + 1] + u[2 * k2 + 1]);
            B00 += m0 * nx; B01 += m0 * ny; B10 += m1 * nx; B11 += m1 * ny;
        }
        const double iv = -1.0 / (vbox[2] * vbox[3]);
        B00 *= iv; B01 *= iv; B10 *= iv; B11 *= iv;
        if (sym_B) { const double m = 0.5 * (B01 + B10); B01 = m; B10 = m; }   // micro_model_gen.py:122
    }
    Bmat[4 * (size_t)gid] = B00; Bmat[4 * (size_t)gid + 1] = B01;
    Bmat[4 * (size_t)gid + 2] = B10; Bmat[4 * (size_t)gid + 3] = B11;
    eps_eff[3 * (size_t)gid] = eps[3 * (size_t)gid] - B00;
    eps_eff[3 * (size_t)gid + 1] = eps[3 * (size_t)gid + 1] - B11;
    eps_eff[3 * (size_t)gid + 2] = eps[3 * (size_t)gid + 2] - (B01 + B10);
}

// Dirichlet data at the RVE boundary nodes: g = mean over the edge end vertices of the P1 trace of
// (uD + B x) ([EXT] dolfin restriction semantics); thread per (boundary node, rhs).
__global__ void k_bnd_values(int n_bnd, int K, int nside, const int* __restrict__ bnd_sys,
                             const int* __restrict__ bnd_node, const double* __restrict__ bnd_s,
                             const double* __restrict__ node_xy, const double* __restrict__ uD,
                             const double* __restrict__ Bmat, double* __restrict__ gval /*[n_bnd][K][2]*/) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)n_bnd * K) return;
    const int ib = (int)(gid / K), k = (int)(gid % K);
    const int sys = bnd_sys[ib];
    const int nb = 4 * nside;
    const double* u = uD + ((size_t)sys * K + k) * 2 * (nb + 1);
    double g0 = 0, g1 = 0;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const double s = bnd_s[2 * (size_t)ib + e];
        int k0 = (int)floor(s + 1e-12);
        const double t = s - k0;
        k0 = k0 % nb;
        const int k1 = (k0 + 1) % nb;
        g0 += 0.5 * ((1.0 - t) * u[2 * k0] + t * u[2 * k1]);
        g1 += 0.5 * ((1.0 - t) * u[2 * k0 + 1] + t * u[2 * k1 + 1]);
    }
    const double* B = Bmat + 4 * ((size_t)sys * K + k);
    const int node = bnd_node[ib];
    const double x = node_xy[2 * (size_t)node], y = node_xy[2 * (size_t)node + 1];
    gval[2 * (size_t)gid] = g0 + B[0] * x + B[1] * y;
    gval[2 * (size_t)gid + 1] = g1 + B[2] * x + B[3] * y;
}

// Deterministic load vector: thread per block row (CTA per window) walks the elements around its row, like k_assemble:
// b_row = sum_e [ -(sigma(eps) . int grad phi_a) - sum_{b prescribed} K_e[a][b] g_b ], one writer per row, no atomics,
// the same summation order on every run.  The per-window partial of the squared unassembled contributions (absolute
// convergence floor) goes to part[win][8]; k_fin_bref adds them per system in a fixed order.
__global__ void __launch_bounds__(RVE_WIN)
k_rhs_rows(WinTab wt, int K, const unsigned* __restrict__ adjf_ptr, const int* __restrict__ adjf, const int* __restrict__ elem_node,
           const unsigned char* __restrict__ elem_reg, const double* __restrict__ node_xy, const int* __restrict__ node_row,
           const int* __restrict__ node_bnd, const double* __restrict__ eps_eff, const double* __restrict__ gval,
           double* __restrict__ b, double* __restrict__ part, Mat8 mat) {
    __shared__ double s_mat[8];
    __shared__ double s_m2[RVE_WIN / 32][4];
    if (threadIdx.x < 8) s_mat[threadIdx.x] = mat.v[threadIdx.x];
    __syncthreads();
    const int w = blockIdx.x, t = threadIdx.x, lane = t & 31, wv = t >> 5;
    const int sys = wt.win_sys[w];
    double acc[RVE_MAX_RHS][2], m2[RVE_MAX_RHS];
#pragma unroll
    for (int k = 0; k < RVE_MAX_RHS; ++k) { acc[k][0] = acc[k][1] = 0.0; m2[k] = 0.0; }
    if (t < wt.win_n[w]) {
        const int r = wt.win_row0[w] + t;
        for (unsigned i = adjf_ptr[r]; i < adjf_ptr[r + 1]; ++i) {
            const int e = adjf[i];
            const int* en = elem_node + 6 * (size_t)e;
            int a = 0;
#pragma unroll
            for (int k = 0; k < 6; ++k) if (node_row[en[k]] == r) a = k;
            const double* p0 = node_xy + 2 * (size_t)en[0]; const double* p1 = node_xy + 2 * (size_t)en[1]; const double* p2 = node_xy + 2 * (size_t)en[2];
            TriGeom g;
            tri_geom(p0[0], p0[1], p1[0], p1[1], p2[0], p2[1], g);
            const int reg = elem_reg[e];
            const double lam = s_mat[2 * reg], mu = s_mat[2 * reg + 1];
            double gx, gy;
            p2_grad_int(g, a, gx, gy);
            for (int k = 0; k < K; ++k) {
                const double* ev = eps_eff + 3 * ((size_t)sys * K + k);
                const double s00 = (lam + 2 * mu) * ev[0] + lam * ev[1];
                const double s11 = lam * ev[0] + (lam + 2 * mu) * ev[1];
                const double s01 = mu * ev[2];
                const double f0 = -(s00 * gx + s01 * gy), f1 = -(s01 * gx + s11 * gy);
                acc[k][0] += f0; acc[k][1] += f1;
                m2[k] += f0 * f0 + f1 * f1;
            }
            if (gval) {
                for (int bb = 0; bb < 6; ++bb) {
                    if (bb == a) continue;
                    const int nb = en[bb];
                    if (node_row[nb] >= 0) continue;
                    const int ib = node_bnd[nb];
                    if (ib < 0) continue;
                    double Kb[4];
                    p2_stiff_block(g, a, bb, lam, mu, Kb);
                    for (int k = 0; k < K; ++k) {
                        const double g0 = gval[((size_t)ib * K + k) * 2], g1 = gval[((size_t)ib * K + k) * 2 + 1];
                        acc[k][0] -= Kb[0] * g0 + Kb[1] * g1;
                        acc[k][1] -= Kb[2] * g0 + Kb[3] * g1;
                    }
                }
            }
        }
        for (int k = 0; k < K; ++k) { b[((size_t)r * K + k) * 2] = acc[k][0]; b[((size_t)r * K + k) * 2 + 1] = acc[k][1]; }
    }
    for (int k = 0; k < K; ++k) {
        const double v = warp_sum(m2[k]);
        if (lane == 0) s_m2[wv][k] = v;
    }
    __syncthreads();
    if (t < 4) {
        double tot = 0.0;
        if (t < K) for (int i = 0; i < RVE_WIN / 32; ++i) tot += s_m2[i][t];
        part[(size_t)w * 8 + t] = tot;
    }
}

__global__ void __launch_bounds__(256)
k_fin_bref(int n_sys, int K, const int* __restrict__ sys_win0, const double* __restrict__ part, double* __restrict__ sc) {
    const int sys = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (sys >= n_sys) return;
    const int w0 = sys_win0[sys], w1 = sys_win0[sys + 1];
    for (int k = 0; k < K; ++k) {
        const double v = sys_sum(part, w0, w1, 8, k);
        if ((threadIdx.x & 31) == 0) sc[((size_t)sys * 4 + k) * SC_N + SC_BREF] = v;
    }
}

// MR: the three unit-traction load vectors T_j = C_o^T e_j (j = 00, 11, 01+10); thread per boundary edge.
__global__ void k_rhs_mr(int n_bedge, const int* __restrict__ bedge, const double* __restrict__ bedge_nl,
                         const int* __restrict__ node_row, double* __restrict__ b /*K=3*/) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n_bedge) return;
    const double nx = bedge_nl[2 * (size_t)gid], ny = bedge_nl[2 * (size_t)gid + 1];
    for (int k = 0; k < 3; ++k) {
        const int row = node_row[bedge[3 * (size_t)gid + k]];
        const double wq = (k == 2) ? 4.0 / 6.0 : 1.0 / 6.0;
        double* br = b + (size_t)row * 3 * 2;
        atomicAdd(br + 0, wq * nx);              // T_0: (comp 0, n_x)
        atomicAdd(br + 2 + 1, wq * ny);          // T_1: (comp 1, n_y)
        atomicAdd(br + 4 + 0, wq * ny);          // T_2: (comp 0, n_y) + (comp 1, n_x)
        atomicAdd(br + 4 + 1, wq * nx);
    }
}

// ---------------------------------------------------------------------------------------------
// (5) constraint post-processing ("Schur step" of MR, zero average of per) and the nodal field
// ---------------------------------------------------------------------------------------------
// per-system mixing data: U[node][l] = sum_j X[row][j] Mx[j][l] + G_l y + a_l (+ g on Dirichlet nodes)
struct MixPtrs {
    double* Mx;    // [sys][4][4]  (j, l)
    double* Gaff;  // [sys][4][4]  (l, 2x2)
    double* aaff;  // [sys][4][2]
};

// one CTA per system: reductions C_m X (wmean . x) and, for MR, C_o X over the boundary edges,
// then the small dense algebra by thread 0.
__global__ void __launch_bounds__(256)
k_mix_coeffs(int model, int K /*solver rhs*/, int NL /*loads*/, const double* __restrict__ x,
             const double* __restrict__ wmean, const int* __restrict__ row_base,
             const int* __restrict__ bedge_base, const int* __restrict__ bedge,
             const double* __restrict__ bedge_nl, const int* __restrict__ node_row,
             const double* __restrict__ box, const double* __restrict__ eps_eff, MixPtrs M) {
    const int sys = blockIdx.x, tid = threadIdx.x;
    __shared__ double s_red[256];
    __shared__ double s_cm[4][2];      // [j][c]
    __shared__ double s_co[4][4];      // [j][2*i+jn] = int x_i n_jn ds
    double cm[8], co[16];
    for (int k = 0; k < 8; ++k) cm[k] = 0.0;
    for (int k = 0; k < 16; ++k) co[k] = 0.0;
    if (model == RVE_MODEL_PER || model == RVE_MODEL_MR) {
        for (int r = row_base[sys] + tid; r < row_base[sys + 1]; r += 256) {
            const double wv = wmean[r];
            for (int j = 0; j < K; ++j) {
                cm[2 * j] += wv * x[((size_t)r * K + j) * 2];
                cm[2 * j + 1] += wv * x[((size_t)r * K + j) * 2 + 1];
            }
        }
    }
    if (model == RVE_MODEL_MR) {
        for (int e = bedge_base[sys] + tid; e < bedge_base[sys + 1]; e += 256) {
            const double nx = bedge_nl[2 * (size_t)e], ny = bedge_nl[2 * (size_t)e + 1];
            for (int k = 0; k < 3; ++k) {
                const int row = node_row[bedge[3 * (size_t)e + k]];
                const double wq = (k == 2) ? 4.0 / 6.0 : 1.0 / 6.0;
                for (int j = 0; j < K; ++j) {
                    const double u0 = x[((size_t)row * K + j) * 2], u1 = x[((size_t)row * K + j) * 2 + 1];
                    co[4 * j + 0] += wq * u0 * nx; co[4 * j + 1] += wq * u0 * ny;
                    co[4 * j + 2] += wq * u1 * nx; co[4 * j + 3] += wq * u1 * ny;
                }
            }
        }
    }
    for (int v = 0; v < 24; ++v) {
        s_red[tid] = (v < 8) ? cm[v] : co[v - 8];
        __syncthreads();
        for (int st = 128; st > 0; st >>= 1) {
            if (tid < st) s_red[tid] += s_red[tid + st];
            __syncthreads();
        }
        if (tid == 0) { if (v < 8) s_cm[v / 2][v % 2] = s_red[0]; else s_co[(v - 8) / 4][(v - 8) % 4] = s_red[0]; }
        __syncthreads();
    }
    if (tid != 0) return;
    double* Mx = M.Mx + (size_t)sys * 16;
    double* Ga = M.Gaff + (size_t)sys * 16;
    double* aa = M.aaff + (size_t)sys * 8;
    for (int k = 0; k < 16; ++k) { Mx[k] = 0.0; Ga[k] = 0.0; }
    for (int k = 0; k < 8; ++k) aa[k] = 0.0;
    const double vol = box[4 * sys + 2] * box[4 * sys + 3];
    const double cx = box[4 * sys] + 0.5 * box[4 * sys + 2], cy = box[4 * sys + 1] + 0.5 * box[4 * sys + 3];
    if (model != RVE_MODEL_MR) {
        for (int l = 0; l < NL; ++l) {
            Mx[4 * l + l] = 1.0;
            if (model == RVE_MODEL_PER) { aa[2 * l] = -s_cm[l][0] / vol; aa[2 * l + 1] = -s_cm[l][1] / vol; }
        }
        return;
    }
    // MR: S = T^T Y (3x3), S[i][j] = T_i . Y_j
    double S[3][3], Si[3][3];
    for (int j = 0; j < 3; ++j) { S[0][j] = s_co[j][0]; S[1][j] = s_co[j][3]; S[2][j] = s_co[j][1] + s_co[j][2]; }
    const double det = S[0][0] * (S[1][1] * S[2][2] - S[1][2] * S[2][1]) - S[0][1] * (S[1][0] * S[2][2] - S[1][2] * S[2][0]) +
                       S[0][2] * (S[1][0] * S[2][1] - S[1][1] * S[2][0]);
    const double id = 1.0 / det;
    Si[0][0] = (S[1][1] * S[2][2] - S[1][2] * S[2][1]) * id; Si[0][1] = (S[0][2] * S[2][1] - S[0][1] * S[2][2]) * id;
    Si[0][2] = (S[0][1] * S[1][2] - S[0][2] * S[1][1]) * id; Si[1][0] = (S[1][2] * S[2][0] - S[1][0] * S[2][2]) * id;
    Si[1][1] = (S[0][0] * S[2][2] - S[0][2] * S[2][0]) * id; Si[1][2] = (S[0][2] * S[1][0] - S[0][0] * S[1][2]) * id;
    Si[2][0] = (S[1][0] * S[2][1] - S[1][1] * S[2][0]) * id; Si[2][1] = (S[0][1] * S[2][0] - S[0][0] * S[2][1]) * id;
    Si[2][2] = (S[0][0] * S[1][1] - S[0][1] * S[1][0]) * id;
    for (int l = 0; l < NL; ++l) {
        const double* ev = eps_eff + 3 * ((size_t)sys * NL + l);
        double pi[3];
        for (int j = 0; j < 3; ++j) pi[j] = vol * (Si[j][0] * ev[0] + Si[j][1] * ev[1] + Si[j][2] * ev[2]);
        for (int j = 0; j < 3; ++j) Mx[4 * j + l] = pi[j];
        // u_t = Y pi ; rotation amount from the antisymmetric part of C_o u_t
        double co10 = 0, co01 = 0, cm0 = 0, cm1 = 0;
        for (int j = 0; j < 3; ++j) { co01 += s_co[j][1] * pi[j]; co10 += s_co[j][2] * pi[j]; cm0 += s_cm[j][0] * pi[j]; cm1 += s_cm[j][1] * pi[j]; }
        const double th = (co10 - co01) / (2.0 * vol);
        // u~ = u_t - th*rho - eps y - t,   rho(y) = (-y2, y1)
        const double G00 = -ev[0], G01 = -0.5 * ev[2] + th, G10 = -0.5 * ev[2] - th, G11 = -ev[1];
        Ga[4 * l] = G00; Ga[4 * l + 1] = G01; Ga[4 * l + 2] = G10; Ga[4 * l + 3] = G11;
        // zero average: t = (C_m u_t + G int y dx)/vol
        aa[2 * l] = -(cm0 / vol + G00 * cx + G01 * cy);
        aa[2 * l + 1] = -(cm1 / vol + G10 * cx + G11 * cy);
    }
}

// nodal fluctuation field on ALL P2 nodes: thread per (node, load)
__global__ void k_nodal_field(long long n_nodes, int K, int NL, const int* __restrict__ node_sys,
                              const int* __restrict__ node_row, const int* __restrict__ node_bnd,
                              const double* __restrict__ node_xy, const double* __restrict__ x,
                              const double* __restrict__ gval, MixPtrs M, int identity_mix,
                              double* __restrict__ U /*[node][NL][2]*/) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n_nodes * NL) return;
    const long long node = gid / NL;
    const int l = (int)(gid % NL);
    const int sys = node_sys[node];
    const int row = node_row[node];
    double u0 = 0, u1 = 0;
    if (row >= 0) {
        if (identity_mix) {
            u0 = x[((size_t)row * K + l) * 2]; u1 = x[((size_t)row * K + l) * 2 + 1];
        } else {
            const double* Mx = M.Mx + (size_t)sys * 16;
            for (int j = 0; j < K; ++j) {
                u0 += x[((size_t)row * K + j) * 2] * Mx[4 * j + l];
                u1 += x[((size_t)row * K + j) * 2 + 1] * Mx[4 * j + l];
            }
        }
        const double* G = M.Gaff + (size_t)sys * 16 + 4 * l;
        const double* a = M.aaff + (size_t)sys * 8 + 2 * l;
        const double px = node_xy[2 * (size_t)node], py = node_xy[2 * (size_t)node + 1];
        u0 += G[0] * px + G[1] * py + a[0];
        u1 += G[2] * px + G[3] * py + a[1];
    } else if (gval) {
        const int ib = node_bnd[node];
        u0 = gval[((size_t)ib * NL + l) * 2]; u1 = gval[((size_t)ib * NL + l) * 2 + 1];
    }
    U[2 * (size_t)gid] = u0; U[2 * (size_t)gid + 1] = u1;
}

// ---------------------------------------------------------------------------------------------
// (6) fused epilogues
// ---------------------------------------------------------------------------------------------
// accumulators per system: [0..3] vol_r, [4..5] int y over regions {0,1}; per load l (base 6+16*l):
//   [0..1] int u~ (L), [2..5] int grad u~ (L), [6..8] int sigma(u~) (L: 00,11,01), [9..11] same over all regions,
//   [12..15] int grad u~ over all regions
#define ACC_PER_SYS (6 + 16 * RVE_MAX_RHS)

// add v from every lane to *addr: one atomic per warp when all lanes target the same system
__device__ __forceinline__ void warp_acc(double v, double* addr, bool uniform, bool valid) {
    if (uniform) {
        v = warp_sum(v);
        if ((threadIdx.x & 31) == 0) atomicAdd(addr, v);
    } else if (valid) {
        atomicAdd(addr, v);
    }
}

__global__ void __launch_bounds__(128)
k_epi_elements(int n_elem, int NL, const int* __restrict__ elem_node, const unsigned char* __restrict__ elem_reg,
               const int* __restrict__ elem_sys, const double* __restrict__ node_xy,
               const double* __restrict__ U, double* __restrict__ acc, Mat8 mat) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = e < n_elem;
    const int sys0 = __shfl_sync(0xffffffffu, valid ? elem_sys[e] : -1, 0);
    if (sys0 < 0) return;                               // whole warp past the end
    const int sys = valid ? elem_sys[e] : sys0;
    const bool uniform = __all_sync(0xffffffffu, sys == sys0);
    int nd[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) nd[k] = valid ? elem_node[(size_t)e * 6 + k] : 0;
    TriGeom g;
    g.area = 0.0;
    int reg = 0;
    double cxa = 0.0, cya = 0.0;
    if (valid) {
        tri_geom(node_xy[2 * (size_t)nd[0]], node_xy[2 * (size_t)nd[0] + 1], node_xy[2 * (size_t)nd[1]],
                 node_xy[2 * (size_t)nd[1] + 1], node_xy[2 * (size_t)nd[2]], node_xy[2 * (size_t)nd[2] + 1], g);
        reg = elem_reg[e];
        cxa = g.area * (node_xy[2 * (size_t)nd[0]] + node_xy[2 * (size_t)nd[1]] + node_xy[2 * (size_t)nd[2]]) / 3.0;
        cya = g.area * (node_xy[2 * (size_t)nd[0] + 1] + node_xy[2 * (size_t)nd[1] + 1] + node_xy[2 * (size_t)nd[2] + 1]) / 3.0;
    }
    const double lam = mat.v[2 * reg], mu = mat.v[2 * reg + 1];
    double* A = acc + (size_t)sys * ACC_PER_SYS;
    const bool inL = valid && reg < 2;
#pragma unroll
    for (int rg = 0; rg < 4; ++rg) warp_acc((valid && reg == rg) ? g.area : 0.0, A + rg, uniform, valid && reg == rg);
    warp_acc(inL ? cxa : 0.0, A + 4, uniform, inL);
    warp_acc(inL ? cya : 0.0, A + 5, uniform, inL);
    double gx[6], gy[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) { gx[a] = 0.0; gy[a] = 0.0; if (valid) p2_grad_int(g, a, gx[a], gy[a]); }
    for (int l = 0; l < NL; ++l) {
        double G00 = 0, G01 = 0, G10 = 0, G11 = 0, ui0 = 0, ui1 = 0;
        if (valid) {
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                const double u0 = U[((size_t)nd[a] * NL + l) * 2], u1 = U[((size_t)nd[a] * NL + l) * 2 + 1];
                G00 += u0 * gx[a]; G01 += u0 * gy[a]; G10 += u1 * gx[a]; G11 += u1 * gy[a];
                if (a >= 3) { ui0 += u0; ui1 += u1; }
            }
        }
        ui0 *= g.area / 3.0; ui1 *= g.area / 3.0;
        const double tr = G00 + G11;
        const double s00 = lam * tr + 2 * mu * G00, s11 = lam * tr + 2 * mu * G11, s01 = mu * (G01 + G10);
        double* Al = A + 6 + 16 * l;
        const double vL[9] = {ui0, ui1, G00, G01, G10, G11, s00, s11, s01};
#pragma unroll
        for (int q = 0; q < 9; ++q) warp_acc(inL ? vL[q] : 0.0, Al + q, uniform, inL);
        warp_acc(valid ? s00 : 0.0, Al + 9, uniform, valid);
        warp_acc(valid ? s11 : 0.0, Al + 10, uniform, valid);
        warp_acc(valid ? s01 : 0.0, Al + 11, uniform, valid);
        warp_acc(valid ? G00 : 0.0, Al + 12, uniform, valid);
        warp_acc(valid ? G01 : 0.0, Al + 13, uniform, valid);
        warp_acc(valid ? G10 : 0.0, Al + 14, uniform, valid);
        warp_acc(valid ? G11 : 0.0, Al + 15, uniform, valid);
    }
}

// thread per (sys, load): sigma over {0,1}, sigma over all, a, B, tangent column
__global__ void k_epi_finalize(int n_sys, int NL, const double* __restrict__ acc, const double* __restrict__ eps_eff,
                               Mat8 mat, double* __restrict__ sigL, double* __restrict__ sigT,
                               double* __restrict__ aout, double* __restrict__ Bout, double* __restrict__ gradT) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n_sys * NL) return;
    const int sys = gid / NL, l = gid % NL;
    const double* A = acc + (size_t)sys * ACC_PER_SYS;
    const double* Al = A + 6 + 16 * l;
    const double* ev = eps_eff + 3 * (size_t)gid;
    const double volL = A[0] + A[1], volT = A[0] + A[1] + A[2] + A[3];
    double sL[3] = {Al[6], Al[7], Al[8]}, sT[3] = {Al[9], Al[10], Al[11]};
    for (int rg = 0; rg < 4; ++rg) {
        const double lam = mat.v[2 * rg], mu = mat.v[2 * rg + 1], v = A[rg];
        const double s00 = (lam + 2 * mu) * ev[0] + lam * ev[1], s11 = lam * ev[0] + (lam + 2 * mu) * ev[1], s01 = mu * ev[2];
        if (rg < 2) { sL[0] += v * s00; sL[1] += v * s11; sL[2] += v * s01; }
        sT[0] += v * s00; sT[1] += v * s11; sT[2] += v * s01;
    }
    for (int k = 0; k < 3; ++k) { sigL[3 * (size_t)gid + k] = sL[k] / volL; sigT[3 * (size_t)gid + k] = sT[k] / volT; }
    const double yG0 = A[4] / volL, yG1 = A[5] / volL;
    const double G00 = Al[2] / volL, G01 = Al[3] / volL, G10 = Al[4] / volL, G11 = Al[5] / volL;
    aout[2 * (size_t)gid] = -Al[0] / volL + G00 * yG0 + G01 * yG1;
    aout[2 * (size_t)gid + 1] = -Al[1] / volL + G10 * yG0 + G11 * yG1;
    Bout[4 * (size_t)gid] = -G00; Bout[4 * (size_t)gid + 1] = -G01;
    Bout[4 * (size_t)gid + 2] = -G10; Bout[4 * (size_t)gid + 3] = -G11;
    if (gradT)
        for (int k = 0; k < 4; ++k) gradT[4 * (size_t)gid + k] = Al[12 + k] / volT;
}

// internal-boundary DOF extraction: P2 evaluation at the Vref points; thread per (sys, point, load)
__global__ void k_epi_sample(long long n, int nvref, int NL, const int* __restrict__ samp_elem /*global elem*/,
                             const double* __restrict__ samp_lam, const int* __restrict__ elem_node,
                             const double* __restrict__ U, double* __restrict__ sol /*[sys][NL][2*nvref]*/) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n) return;
    const int l = (int)(gid % NL);
    const long long sp = gid / NL;           // sys*nvref + point
    const int pt = (int)(sp % nvref);
    const long long sys = sp / nvref;
    const int e = samp_elem[sp];
    const double l1 = samp_lam[2 * sp], l2 = samp_lam[2 * sp + 1];
    double N[6];
    p2_shape(1.0 - l1 - l2, l1, l2, N);
    double u0 = 0, u1 = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        const int nd = elem_node[(size_t)e * 6 + a];
        u0 += N[a] * U[((size_t)nd * NL + l) * 2];
        u1 += N[a] * U[((size_t)nd * NL + l) * 2 + 1];
    }
    double* o = sol + ((size_t)sys * NL + l) * 2 * nvref + 2 * pt;
    o[0] = u0; o[1] = u1;
}

// RB projection Y = (U + translate) M W^T with MWt[l][q][j] precomputed; CTA per (sys, load)
__global__ void k_project(int NL, int nh, int nmax, const double* __restrict__ sol,
                          const double* __restrict__ translate, const double* __restrict__ MWt,
                          double* __restrict__ Y) {
    extern __shared__ double s_u[];
    const int sl = blockIdx.x;  // sys*NL + l
    const int l = sl % NL;
    for (int q = threadIdx.x; q < nh; q += blockDim.x) {
        double v = sol[(size_t)sl * nh + q];
        if (translate) v += translate[2 * (size_t)sl + (q & 1)];
        s_u[q] = v;
    }
    __syncthreads();
    const double* Wl = MWt + (size_t)l * nh * nmax;
    for (int j = threadIdx.x; j < nmax; j += blockDim.x) {
        double acc = 0.0;
        for (int q = 0; q < nh; ++q) acc += s_u[q] * Wl[(size_t)q * nmax + j];
        Y[(size_t)sl * nmax + j] = acc;
    }
}
