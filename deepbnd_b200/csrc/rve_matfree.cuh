// rve_matfree.cuh — matrix-free application of the P2 elasticity operator (sm_100a).
//
// The PCG's q = A p does not need the assembled matrix: the stored SELL-32 operator is 34 B per 2x2 block (about
// 15 MB per `full` RVE, read once per iteration), while the same operator is defined by 72 B per element (about
// 1.7 MB per RVE, counting the elements on window borders twice): 6 tile-local columns, 6 destination slots, and
// 6 doubles of geometry x material.  The fp64 pipe of the B200 has the headroom (about 150 fp64 operations per
// element and right-hand side), so the application becomes a gather -> element product -> gather pass over
// shared memory with coalesced streams only:
//
//   window w (RVE_WIN rows, one CTA): the elements touching its own rows ("window elements", found by
//   k_sym_elems with a shared-memory bitmap over the elements of the system; elements on a border belong to
//   several windows and are applied by each of them, no inter-CTA exchange, no atomics, deterministic)
//     phase 0  p-tile: own rows (coalesced) + halo rows (list-driven gather) -> shared memory   [same as k_spmv]
//     phase 1  thread per (window element, rhs): u_e from the tile, f_e = K_e u_e by p2_apply (K_e never formed),
//              the nodal forces of the element's OWN-row nodes go to their slots in shared memory
//              (slot = position j of the element in that row's fan, sliced ELL [slice][j][32 rows])
//     phase 2  thread per (row, rhs): q = sum of the row's slots (conflict-free), p.q partial, coalesced store
//
// Replaces mp.block_assemble(a) + PETSc's matrix-vector product inside the solve (micro_model.py:60-66): the
// assembled matrix is only built on demand (rve_get_csr, rve_bench_spmv, RVE_OPT_OPERATOR = assembled).
#pragma once
#include "rve_kernels.cuh"

struct MfTab {
    const int* win_we0;           // [n_win] first window element (multiple of 32); arrays below are indexed 6*we0 + a*nep + e
    const int* win_nwe;           // [n_win] window elements (nep = nwe rounded up to 32)
    const unsigned char* win_sw;  // [n_win][8] slots per row of every 32-row slice (max fan size in the slice)
    const unsigned char* row_nadj;// [R] elements around the row
    const uint32_t* we_idx;       // [6][nep] per window: words 0-2 = tile-local columns of the element's 6 nodes (two uint16 per word,
                                  // 0xffff: no row = homogeneous Dirichlet), words 3-5 = slots of the nodal forces (0xffff: not an own row)
    const double2* we_geo;        // [3][nep] per window: (g0x,g0y) (g1x,g1y) (w*lam,w*mu)
};

#define RVE_MF_MAXWE 1024         // window elements per window (a 256-row window touches about 200)

// counters: [2] error code (5: window-element overflow), [6] window elements used, [7] max slots of a window
__global__ void __launch_bounds__(RVE_WIN)
k_sym_elems(WinTab wt, Sell sl, const int* __restrict__ elem_base, const unsigned* __restrict__ adjf_ptr,
            const int* __restrict__ adjf, const int* __restrict__ elem_node, const int* __restrict__ node_row,
            int* __restrict__ win_we0, int* __restrict__ win_nwe, unsigned char* __restrict__ win_sw,
            unsigned char* __restrict__ row_nadj, int* __restrict__ we_elem, uint32_t* __restrict__ we_idx,
            long long we_cap, int* counters) {
    extern __shared__ unsigned s_bm[];           // [words] bitmap over the elements of the system, then popcount prefix
    __shared__ int s_scan[RVE_WIN];
    __shared__ int s_el[RVE_MF_MAXWE];
    __shared__ int s_sb[RVE_WIN / 32 + 1];
    __shared__ int s_w0;
    const int w = blockIdx.x, tid = threadIdx.x, wv = tid >> 5;
    const int sys = wt.win_sys[w];
    const int e0 = elem_base[sys], ne = elem_base[sys + 1] - e0;
    const int words = (ne + 31) >> 5;
    unsigned* s_pre = s_bm + words;
    const int r0 = wt.win_row0[w], n = wt.win_n[w];
    for (int i = tid; i < words; i += RVE_WIN) s_bm[i] = 0u;
    __syncthreads();
    unsigned a0 = 0; int cnt = 0;
    if (tid < n) {
        a0 = adjf_ptr[r0 + tid]; cnt = (int)(adjf_ptr[r0 + tid + 1] - a0);
        for (int j = 0; j < cnt; ++j) { const int e = adjf[a0 + j] - e0; atomicOr(&s_bm[e >> 5], 1u << (e & 31)); }
        row_nadj[r0 + tid] = (unsigned char)cnt;
    }
    // slots per row of the slice = largest fan in it
    int mx = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    // every lane stores the warp maximum (same value): a store predicated on (tid & 31) == 0 made nvcc 12.9 reuse the
    // address (tid >> 3), valid under that predicate only, for the unpredicated s_sb[tid >> 5] read further down
    // (LEA.HI without the mask -> "misaligned address" on the device)
    s_sb[wv + 1] = mx;
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
    const int nwe = s_scan[RVE_WIN - 1];
    const int nep = (nwe + 31) & ~31;
    if (tid == 0) {
        int tot = 0;
        for (int s = 0; s < RVE_WIN / 32; ++s) {
            const int wd = s_sb[s + 1];
            win_sw[(size_t)w * 8 + s] = (unsigned char)wd;
            s_sb[s] = tot; tot += wd * 32;             // first slot of the slice
        }
        s_sb[RVE_WIN / 32] = tot;
        atomicMax(counters + 7, tot);
        int we0 = atomicAdd(counters + 6, nep);
        if ((long long)we0 + nep > we_cap || nwe > RVE_MF_MAXWE) { atomicExch(counters + 2, 5); we0 = -1; }
        s_w0 = we0;
        win_we0[w] = we0 < 0 ? 0 : we0;
        win_nwe[w] = we0 < 0 ? 0 : nwe;
    }
    __syncthreads();
    const int we0 = s_w0;
    if (we0 < 0) return;
    for (int i = tid; i < words; i += RVE_WIN) {       // window-local element list (ascending element id)
        unsigned bits = s_bm[i];
        int pos = (int)s_pre[i];
        while (bits) {
            const int b = __ffs(bits) - 1;
            s_el[pos++] = i * 32 + b;
            bits &= bits - 1;
        }
    }
    // packed record, viewed as uint16: column of node a = half (a & 1) of word a >> 1, slot of node a = the same in word 3 + (a >> 1)
    uint16_t* rec = (uint16_t*)(we_idx + 6 * (size_t)we0);
#define WE_COL(a, el) rec[2 * (((a) >> 1) * nep + (el)) + ((a) & 1)]
#define WE_DST(a, el) rec[2 * ((3 + ((a) >> 1)) * nep + (el)) + ((a) & 1)]
    for (int i = tid; i < 6 * nep; i += RVE_WIN) rec[6 * nep + i] = (uint16_t)0xffffu;
    __syncthreads();
    // columns of the 6 nodes of every window element: own rows first, then the position in the (sorted) halo list
    const int h0 = wt.win_halo0[w], nh = wt.win_nhalo[w];
    for (int el = tid; el < nep; el += RVE_WIN) {
        if (el >= nwe) {
#pragma unroll
            for (int a = 0; a < 6; ++a) WE_COL(a, el) = (uint16_t)0xffffu;
            we_elem[we0 + el] = -1;
            continue;
        }
        const int e = e0 + s_el[el];
        we_elem[we0 + el] = e;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            const int c = node_row[elem_node[6 * (size_t)e + a]];
            int lc = 0xffff;
            if (c >= r0 && c < r0 + n) lc = c - r0;
            else if (c >= 0) {
                int lo = 0, hi = nh;                   // first halo entry >= c
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (sl.halo[h0 + mid] < c) lo = mid + 1; else hi = mid; }
                lc = n + lo;
            }
            WE_COL(a, el) = (uint16_t)lc;
        }
    }
    // destination slots: row t, j-th element of its fan
    if (tid < n) {
        const int r = r0 + tid;
        const int sbase = s_sb[wv] + (tid & 31);
        for (int j = 0; j < cnt; ++j) {
            const int e = adjf[a0 + j];
            const int* en = elem_node + 6 * (size_t)e;
            int a = 0;
#pragma unroll
            for (int k = 0; k < 6; ++k) if (node_row[en[k]] == r) a = k;
            const int el = (int)s_pre[(e - e0) >> 5] + __popc(s_bm[(e - e0) >> 5] & ((1u << ((e - e0) & 31)) - 1u));
            WE_DST(a, el) = (uint16_t)(sbase + j * 32);
        }
    }
}

// geometry x material of the window elements (numeric half: redone when coordinates or materials change)
__global__ void __launch_bounds__(RVE_WIN)
k_we_geo(int n_win, const int* __restrict__ win_we0, const int* __restrict__ win_nwe, const int* __restrict__ we_elem,
         const int* __restrict__ elem_node, const unsigned char* __restrict__ elem_reg,
         const double* __restrict__ node_xy, double2* __restrict__ we_geo, Mat8 mat) {
    const int w = blockIdx.x;
    const int we0 = win_we0[w], nwe = win_nwe[w], nep = (nwe + 31) & ~31;
    double2* geo = we_geo + 3 * (size_t)we0;
    for (int el = threadIdx.x; el < nep; el += RVE_WIN) {
        double gv[6] = {0, 0, 0, 0, 0, 0};
        if (el < nwe) {
            const int e = we_elem[we0 + el];
            const int* en = elem_node + 6 * (size_t)e;
            const double* p0 = node_xy + 2 * (size_t)en[0]; const double* p1 = node_xy + 2 * (size_t)en[1]; const double* p2 = node_xy + 2 * (size_t)en[2];
            TriGeom g;
            tri_geom(p0[0], p0[1], p1[0], p1[1], p2[0], p2[1], g);
            const int reg = elem_reg[e];
            p2_apply_geo(g, mat.v[2 * reg], mat.v[2 * reg + 1], gv);
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) geo[j * nep + el] = make_double2(gv[2 * j], gv[2 * j + 1]);
    }
}

// block-Jacobi diagonal and the mean-value weights without the assembled matrix: thread per block row walks the
// elements around it (the diagonal block of K_e for its own local node); CTA per window
__global__ void __launch_bounds__(RVE_WIN)
k_diag(WinTab wt, const unsigned* __restrict__ adjf_ptr, const int* __restrict__ adjf, const int* __restrict__ elem_node,
       const unsigned char* __restrict__ elem_reg, const double* __restrict__ node_xy, const int* __restrict__ node_row,
       float4* __restrict__ dinv32, double* __restrict__ wmean, Mat8 mat) {
    const int w = blockIdx.x, t = threadIdx.x;
    if (t >= wt.win_n[w]) return;
    const int r = wt.win_row0[w] + t;
    double d00 = 0, d01 = 0, d10 = 0, d11 = 0, wm = 0;
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
        double Kb[4];
        p2_stiff_block(g, a, a, mat.v[2 * reg], mat.v[2 * reg + 1], Kb);
        d00 += Kb[0]; d01 += Kb[1]; d10 += Kb[2]; d11 += Kb[3];
        if (a >= 3) wm += g.area * (1.0 / 3.0);
    }
    const double det = d00 * d11 - d01 * d10;
    const double inv = det != 0.0 ? 1.0 / det : 0.0;
    dinv32[r] = make_float4((float)(d11 * inv), (float)(-d01 * inv), (float)(-d10 * inv), (float)(d00 * inv));
    wmean[r] = wm;
}

#define RVE_APPLY_THREADS 256
// p-tile layout in shared memory: interleaved [row][rhs] (default) or planar [rhs][row] with the planes pitched to
// 4 mod 8 rows.  Planar was meant to spread the K = 2 gathers of one rhs over all 8 16-byte bank groups instead of
// every other one; measured (r02f, 1000 full / 10k reduced RVEs): K = 2 +1 %, K = 3 +5 % time per launch -> not used.
#ifndef RVE_APPLY_PLANAR
#define RVE_APPLY_PLANAR 0
#endif
__host__ __device__ inline int apply_tile_pitch(int tile_cap) { return RVE_APPLY_PLANAR ? ((tile_cap + 3) & ~7) + 4 : tile_cap; }
#ifndef RVE_APPLY_MINB
#define RVE_APPLY_MINB(K) ((K) == 3 ? 3 : 4)
#endif

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// q = A p and the per-window partial of p.q: see the header of this file.
// Latency plan: the p-tile travels global -> shared with cp.async (LDGSTS, no registers, no wait), the halo row ids
// and the element record of the thread's first element are requested before anything is waited for, so the three
// dependent latencies (halo ids -> halo rows, tile -> element product) overlap.
template <int K>
__global__ void __launch_bounds__(RVE_APPLY_THREADS, RVE_APPLY_MINB(K))
k_apply(WinTab wt, Sell sl, MfTab mf, const double2* __restrict__ p, double2* __restrict__ q, int tile_cap, int slot_cap,
        double* __restrict__ part) {
    extern __shared__ double2 s_tile[];           // [tile_cap * K] p-tile (row-major, rhs interleaved), then K planes [slot_cap] of nodal forces
    constexpr int TPB = RVE_APPLY_THREADS;
    const int win = blockIdx.x;
    // (requesting the window metadata together with the active flag -- volatile loads ahead of the early exit -- was
    //  measured: no change, r02f)
    if (!wt.win_active[win]) return;
    const int tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;
    const int row0 = wt.win_row0[win], n = wt.win_n[win];
    const int h0 = wt.win_halo0[win], nh = wt.win_nhalo[win];
    const int we0 = mf.win_we0[win], nwe = mf.win_nwe[win], nep = (nwe + 31) & ~31;
    const int pitch = apply_tile_pitch(tile_cap);
#if RVE_APPLY_PLANAR
#define TIDX(c, k) ((k) * pitch + (c))
#else
#define TIDX(c, k) ((c) * K + (k))
#endif
    double2* s_f = s_tile + (size_t)pitch * K;
    __shared__ int s_sb[RVE_WIN / 32];
    // halo row ids first (the gather below depends on them)
    int hr[2];
#pragma unroll
    for (int it = 0; it < 2; ++it) { const int e = tid + it * TPB; hr[it] = (e < nh) ? sl.halo[h0 + e] : -1; }
    if (tid == 0) {
        const uint2 v = *(const uint2*)(mf.win_sw + (size_t)win * 8);
        int tot = 0;
#pragma unroll
        for (int s = 0; s < 8; ++s) { s_sb[s] = tot; tot += 32 * (int)(((s < 4 ? v.x : v.y) >> (8 * (s & 3))) & 0xffu); }
    }
    {   // own rows: one contiguous copy
        const double2* src = p + (size_t)row0 * K;
        for (int e = tid; e < n * K; e += TPB) cp_async16(s_tile + TIDX(e / K, e % K), src + e);
    }
    // element record of the first element of this thread, rows' fan sizes
    const uint32_t* idx = mf.we_idx + 6 * (size_t)we0;
    const double2* geo = mf.we_geo + 3 * (size_t)we0;
    uint32_t ix[6];
    double2 gv[3];
    if (tid < nwe) {
#pragma unroll
        for (int j = 0; j < 6; ++j) ix[j] = idx[j * nep + tid];
#pragma unroll
        for (int j = 0; j < 3; ++j) gv[j] = geo[j * nep + tid];
    }
    const int na = (tid < n) ? mf.row_nadj[row0 + tid] : 0;
    {   // halo rows
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int e = tid + it * TPB;
            if (e < nh) {
#pragma unroll
                for (int k = 0; k < K; ++k) cp_async16(s_tile + TIDX(n + e, k), p + (size_t)hr[it] * K + k);
            }
        }
        for (int e = tid + 2 * TPB; e < nh; e += TPB) {           // (a window has fewer than 512 halo rows in practice)
            const int h = sl.halo[h0 + e];
#pragma unroll
            for (int k = 0; k < K; ++k) cp_async16(s_tile + TIDX(n + e, k), p + (size_t)h * K + k);
        }
    }
    cp_async_wait_all();
    __syncthreads();
    // phase 1: element products, thread per window element, right-hand sides in turn
    for (int e = tid; e < nwe; e += TPB) {
        if (e != tid) {
#pragma unroll
            for (int j = 0; j < 6; ++j) ix[j] = idx[j * nep + e];
#pragma unroll
            for (int j = 0; j < 3; ++j) gv[j] = geo[j * nep + e];
        }
        const double g6[6] = {gv[0].x, gv[0].y, gv[1].x, gv[1].y, gv[2].x, gv[2].y};
#pragma unroll 1
        for (int k = 0; k < K; ++k) {
            double u[12], f[12];
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                const unsigned c = (ix[a >> 1] >> (16 * (a & 1))) & 0xffffu;
                const double2 v = (c != 0xffffu) ? s_tile[TIDX(c, k)] : make_double2(0.0, 0.0);
                u[2 * a] = v.x; u[2 * a + 1] = v.y;
            }
            p2_apply(g6, u, f);
            double2* fk = s_f + (size_t)k * slot_cap;
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                const unsigned d = (ix[3 + (a >> 1)] >> (16 * (a & 1))) & 0xffffu;
                if (d != 0xffffu) fk[d] = make_double2(f[2 * a], f[2 * a + 1]);
            }
        }
    }
    __syncthreads();
    // phase 2: thread per row gathers its slots (conflict-free: consecutive rows, consecutive slots), p.q partials
    double dot[K];
#pragma unroll
    for (int k = 0; k < K; ++k) dot[k] = 0.0;
    if (tid < n) {
        const int sb = s_sb[wv] + lane;
        double2 acc[K];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = make_double2(0.0, 0.0);
        for (int j = 0; j < na; ++j)
#pragma unroll
            for (int k = 0; k < K; ++k) { const double2 v = s_f[(size_t)k * slot_cap + sb + j * 32]; acc[k].x += v.x; acc[k].y += v.y; }
        double2* qd = q + ((size_t)row0 + tid) * K;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            qd[k] = acc[k];
            const double2 pv = s_tile[TIDX(tid, k)];
            dot[k] = acc[k].x * pv.x + acc[k].y * pv.y;
        }
    }
    __shared__ double s_dot[RVE_APPLY_THREADS / 32][4];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const double v = warp_sum(dot[k]);
        if (lane == 0) s_dot[wv][k] = v;
    }
    __syncthreads();
    if (tid < K) {                                   // per-window partial of p.q; k_fin_alpha adds them per system
        double my = 0.0;
        for (int i = 0; i < TPB / 32; ++i) my += s_dot[i][tid];
        part[(size_t)win * 4 + tid] = my;
    }
#undef TIDX
}

// ------------------------------------------------------------------------------------------------------------------
// Persistent k_apply -- MEASURED AND NOT USED (compiled with -DRVE_APPLY_PERSISTENT=1 only).
// With one window per CTA every window pays four dependent global round trips before its first element product
// (active flag -> window metadata -> halo row ids -> halo rows of p): 31 % of the stall samples of k_apply<2>
// (profiles/r02f_ncu_apply2_source_top.txt).  Here a CTA walks the windows w = blockIdx.x, += gridDim.x and requests,
// together with the p-tile of window i, the halo ids of window i+1 and the metadata of window i+2 -- all with cp.async
// into small shared-memory rings, no registers -- so that every window only waits for ONE round trip (its tile).
// Same results (tests green), but 20 % (K = 2) / 6 % (K = 3) SLOWER per launch on 1000 full / 10k reduced RVEs
// (r02f: 1756 vs 1457 us, 2315 vs 2180 us): the resident CTAs of an SM start together and stay in step, so their
// tile waits coincide instead of being covered by each other's element products the way the staggered one-window CTAs
// cover them; hiding the tile wait inside one CTA needs a second tile buffer = one CTA per SM less.
// ------------------------------------------------------------------------------------------------------------------
#ifndef RVE_APPLY_PERSISTENT
#define RVE_APPLY_PERSISTENT 0
#endif
#if RVE_APPLY_PERSISTENT
#define RVE_APPLY_HCAP 512           // halo ids staged per window (more: read straight from global memory)
#ifndef RVE_APPLYP_MINB
#define RVE_APPLYP_MINB(K) ((K) == 3 ? 3 : 4)
#endif
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async8b(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gmem_src) : "memory");
}

// metadata words of a window in shared memory
enum { AM_ACT = 0, AM_ROW0, AM_N, AM_H0, AM_NH, AM_WE0, AM_NWE, AM_PAD, AM_SW0, AM_SW1, AM_WORDS = 12 };

template <int K>
__global__ void __launch_bounds__(RVE_APPLY_THREADS, RVE_APPLYP_MINB(K))
k_apply_p(WinTab wt, Sell sl, MfTab mf, const double2* __restrict__ p, double2* __restrict__ q, int tile_cap, int slot_cap,
          double* __restrict__ part, int n_win) {
    extern __shared__ double2 s_tile[];           // p-tile, then K planes [slot_cap] of nodal forces (as in k_apply)
    constexpr int TPB = RVE_APPLY_THREADS;
    __shared__ int s_halo[2][RVE_APPLY_HCAP];
    __shared__ __align__(8) int s_meta[2][AM_WORDS];
    __shared__ double s_dot[2][TPB / 32][4];
    __shared__ int s_sb[RVE_WIN / 32];
    const int tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;
    const int stride = (int)gridDim.x;
    const int pitch = apply_tile_pitch(tile_cap);
#if RVE_APPLY_PLANAR
#define TIDX(c, k) ((k) * pitch + (c))
#else
#define TIDX(c, k) ((c) * K + (k))
#endif
    double2* s_f = s_tile + (size_t)pitch * K;
    // request the metadata of window w into ring slot `slot` (threads 0..7; the flag word holds 4 neighbouring flags)
    auto fetch_meta = [&](int w, int slot) {
        if (w >= n_win) { if (tid == 0) s_meta[slot][AM_N] = -1; return; }
        int* m = s_meta[slot];
        switch (tid) {
            case 0: cp_async4(m + AM_ACT, (const int*)(wt.win_active + (w & ~3))); break;
            case 1: cp_async4(m + AM_ROW0, wt.win_row0 + w); break;
            case 2: cp_async4(m + AM_N, wt.win_n + w); break;
            case 3: cp_async4(m + AM_H0, wt.win_halo0 + w); break;
            case 4: cp_async4(m + AM_NH, wt.win_nhalo + w); break;
            case 5: cp_async4(m + AM_WE0, mf.win_we0 + w); break;
            case 6: cp_async4(m + AM_NWE, mf.win_nwe + w); break;
            case 7: cp_async8b(m + AM_SW0, mf.win_sw + (size_t)w * 8); break;
            default: break;
        }
    };
    auto is_active = [&](const int* m, int w) { return m[AM_N] >= 0 && ((((unsigned)m[AM_ACT]) >> (8 * (w & 3))) & 0xffu) != 0u; };
    auto fetch_halo = [&](const int* m, int w, int slot) {
        if (!is_active(m, w)) return;
        const int h0 = m[AM_H0], nh = min(m[AM_NH], RVE_APPLY_HCAP);
        for (int e = tid; e < nh; e += TPB) cp_async4(&s_halo[slot][e], sl.halo + h0 + e);
    };
    // prologue: metadata of the first two windows, halo ids of the first
    int w = (int)blockIdx.x;
    fetch_meta(w, 0);
    fetch_meta(w + stride, 1);
    cp_async_wait_all();
    __syncthreads();
    fetch_halo(s_meta[0], w, 0);
    int prev_win = -1;
    for (int it = 0; w < n_win; ++it, w += stride) {
        const int cur = it & 1;
        cp_async_wait_all();                      // (prologue halo ids; otherwise nothing is pending here)
        __syncthreads();                          // (A) rings of this window visible; the previous window's tile, forces and dots are free / final
        if (prev_win >= 0 && tid < K) {           // p.q partial of the previous window
            double my = 0.0;
            for (int i = 0; i < TPB / 32; ++i) my += s_dot[cur ^ 1][i][tid];
            part[(size_t)prev_win * 4 + tid] = my;
        }
        const int* m = s_meta[cur];
        const bool act = is_active(m, w);
        const int row0 = m[AM_ROW0], n = m[AM_N], h0 = m[AM_H0], nh = m[AM_NH], we0 = m[AM_WE0], nwe = m[AM_NWE];
        const int nep = (nwe + 31) & ~31;
        const uint32_t* idx = mf.we_idx + 6 * (size_t)we0;
        const double2* geo = mf.we_geo + 3 * (size_t)we0;
        uint32_t ix[6];
        double2 gv[3];
        int na = 0;
        if (act) {
            if (tid == 0) {
                int tot = 0;
#pragma unroll
                for (int s = 0; s < 8; ++s) { s_sb[s] = tot; tot += 32 * (int)((((unsigned)m[AM_SW0 + (s >> 2)]) >> (8 * (s & 3))) & 0xffu); }
            }
            const double2* src = p + (size_t)row0 * K;
            for (int e = tid; e < n * K; e += TPB) cp_async16(s_tile + TIDX(e / K, e % K), src + e);
            for (int e = tid; e < nh; e += TPB) {
                const int hrow = e < RVE_APPLY_HCAP ? s_halo[cur][e] : sl.halo[h0 + e];
#pragma unroll
                for (int k = 0; k < K; ++k) cp_async16(s_tile + TIDX(n + e, k), p + (size_t)hrow * K + k);
            }
            if (tid < nwe) {
#pragma unroll
                for (int j = 0; j < 6; ++j) ix[j] = idx[j * nep + tid];
#pragma unroll
                for (int j = 0; j < 3; ++j) gv[j] = geo[j * nep + tid];
            }
            na = (tid < n) ? mf.row_nadj[row0 + tid] : 0;
        }
        // the next window's halo ids and the metadata of the one after it travel with this window's tile
        fetch_halo(s_meta[cur ^ 1], w + stride, cur ^ 1);
        cp_async_wait_all();
        __syncthreads();                          // (B) tile complete; s_meta[cur] has been read by every thread
        fetch_meta(w + 2 * stride, cur);          // lands during the phases below, waited for at the top of the next window
        if (!act) { prev_win = -1; continue; }    // (uniform)
        // phase 1: element products
        for (int e = tid; e < nwe; e += TPB) {
            if (e != tid) {
#pragma unroll
                for (int j = 0; j < 6; ++j) ix[j] = idx[j * nep + e];
#pragma unroll
                for (int j = 0; j < 3; ++j) gv[j] = geo[j * nep + e];
            }
            const double g6[6] = {gv[0].x, gv[0].y, gv[1].x, gv[1].y, gv[2].x, gv[2].y};
#pragma unroll 1
            for (int k = 0; k < K; ++k) {
                double u[12], f[12];
#pragma unroll
                for (int a = 0; a < 6; ++a) {
                    const unsigned c = (ix[a >> 1] >> (16 * (a & 1))) & 0xffffu;
                    const double2 v = (c != 0xffffu) ? s_tile[TIDX(c, k)] : make_double2(0.0, 0.0);
                    u[2 * a] = v.x; u[2 * a + 1] = v.y;
                }
                p2_apply(g6, u, f);
                double2* fk = s_f + (size_t)k * slot_cap;
#pragma unroll
                for (int a = 0; a < 6; ++a) {
                    const unsigned d = (ix[3 + (a >> 1)] >> (16 * (a & 1))) & 0xffffu;
                    if (d != 0xffffu) fk[d] = make_double2(f[2 * a], f[2 * a + 1]);
                }
            }
        }
        __syncthreads();                          // (C)
        // phase 2: row sums, p.q partials per warp
        double dot[K];
#pragma unroll
        for (int k = 0; k < K; ++k) dot[k] = 0.0;
        if (tid < n) {
            const int sb = s_sb[wv] + lane;
            double2 acc[K];
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] = make_double2(0.0, 0.0);
            for (int j = 0; j < na; ++j)
#pragma unroll
                for (int k = 0; k < K; ++k) { const double2 v = s_f[(size_t)k * slot_cap + sb + j * 32]; acc[k].x += v.x; acc[k].y += v.y; }
            double2* qd = q + ((size_t)row0 + tid) * K;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                qd[k] = acc[k];
                const double2 pv = s_tile[TIDX(tid, k)];
                dot[k] = acc[k].x * pv.x + acc[k].y * pv.y;
            }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const double v = warp_sum(dot[k]);
            if (lane == 0) s_dot[cur][wv][k] = v;
        }
        prev_win = w;
    }
    __syncthreads();
    if (prev_win >= 0 && tid < K) {
        const int it_last = (prev_win - (int)blockIdx.x) / stride;
        double my = 0.0;
        for (int i = 0; i < TPB / 32; ++i) my += s_dot[it_last & 1][i][tid];
        part[(size_t)prev_win * 4 + tid] = my;
    }
#undef TIDX
}
#endif  // RVE_APPLY_PERSISTENT
