// rve_stream.cuh — persistent, bulk-copy-pipelined versions of the PCG vector kernels (sm_100a).
//
// k_update / k_direction stream 6 resp. 3 vector passes per window and then do a short, latency-bound shared-memory
// phase (restriction to the level-1 grid, block sums, the last-CTA bookkeeping).  With one window per CTA nothing is
// in flight for that CTA during the second phase, and the measured bandwidth stops at 60-66 % of the copy peak.
// Here a CTA is persistent (grid = resident CTAs), walks the windows w = blockIdx.x, += gridDim.x and keeps the NEXT
// window's operands in flight while it works on the current one:
//
//   producer = one extra warp (warp specialisation): finds the CTA's next active window (32 candidates per ballot), loads
//   its 32-byte WinDesc, waits for the stage to be released ("empty" mbarrier) and issues cp.async.bulk (UBLKCP, the
//   1-D TMA path) of the contiguous per-window pieces r, p, q, x, dinv, rowinfo, restriction lists into a 2-stage
//   shared-memory ring, completion counted in bytes by the "full" mbarrier of the stage;
//   consumers = the other warps: mbarrier wait, elementwise update out of shared memory, coalesced stores of x and
//   r, the new residual stays in the stage for the restriction; they synchronise among themselves on a named barrier
//   (the producer never joins), so issuing copies is never on their critical path.
//
// No register staging, no per-thread address arithmetic for the loads, and the copy engine keeps HBM busy during the
// reduction phases.
#pragma once
#include "rve_kernels.cuh"

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// global -> shared bulk copy (bytes: multiple of 16, both addresses 16-byte aligned), completion on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// order generic-proxy accesses of shared memory before the async proxy overwrites it
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// barrier among the consumer warps only (id 1; the producer warp never joins)
__device__ __forceinline__ void consumer_sync(int nthreads) { asm volatile("bar.sync 1, %0;\n" ::"r"(nthreads) : "memory"); }

// first active window at or after w in the sequence w, w+stride, ... (whole warp; returns n_win when there is none)
__device__ __forceinline__ int next_active_window(const unsigned char* __restrict__ act, int w, int n_win, int stride) {
    const int lane = threadIdx.x & 31;
    while (w < n_win) {
        const long long cand = (long long)w + (long long)lane * stride;
        const bool a = cand < n_win && __ldcg(act + cand) != 0;
        const unsigned m = __ballot_sync(0xffffffffu, a);
        if (m) return w + (__ffs(m) - 1) * stride;
        if ((long long)w + 32LL * stride >= n_win) break;
        w += 32 * stride;
    }
    return n_win;
}

// Producer-side walk over the active windows of a CTA (whole warp).  The flags, the descriptor and the per-window
// scalars of the NEXT candidate are requested before the current window's copies are issued and looked at afterwards:
// the producer pays one memory latency per window at most, and none when the next window of the sequence is active.
struct WinCursor {
    int w; WinDesc d; double sv;                 // current window (w >= n_win: none), its descriptor, lane's scalar
    long long w2; unsigned char f2; WinDesc d2; double sv2;
    __device__ __forceinline__ void fetch(const WinDesc* __restrict__ desc, const double* __restrict__ scal, int off, int K, int n_win) {
        d.n = -1; sv = 0.0;
        if (w < n_win) {
            d = desc[w];
            if ((int)(threadIdx.x & 31) < K) sv = scal[(size_t)w * 8 + off + (threadIdx.x & 31)];
        }
    }
    __device__ __forceinline__ void init(const unsigned char* __restrict__ act, const WinDesc* __restrict__ desc,
                                         const double* __restrict__ scal, int off, int K, int n_win, int stride, int first) {
        w = first < n_win ? next_active_window(act, first, n_win, stride) : n_win;
        fetch(desc, scal, off, K, n_win);
    }
    __device__ __forceinline__ void prefetch(const unsigned char* __restrict__ act, const WinDesc* __restrict__ desc,
                                             const double* __restrict__ scal, int off, int K, int n_win, int stride) {
        const int lane = threadIdx.x & 31;
        w2 = (long long)w + stride;
        f2 = 0; d2.n = -1; sv2 = 0.0;
        if (w < n_win && w2 < n_win) {
            const long long cand = w2 + (long long)lane * stride;
            f2 = cand < n_win ? __ldcg(act + cand) : (unsigned char)0;
            d2 = desc[w2];
            if (lane < K) sv2 = scal[(size_t)w2 * 8 + off + lane];
        }
    }
    __device__ __forceinline__ void advance(const unsigned char* __restrict__ act, const WinDesc* __restrict__ desc,
                                            const double* __restrict__ scal, int off, int K, int n_win, int stride) {
        if (w >= n_win || w2 >= n_win) { w = n_win; d.n = -1; return; }
        const unsigned m = __ballot_sync(0xffffffffu, f2 != 0);
        if (m & 1u) { w = (int)w2; d = d2; sv = sv2; return; }
        if (m) w = (int)w2 + (__ffs(m) - 1) * stride;
        else w = (w2 + 32LL * stride < n_win) ? next_active_window(act, (int)(w2 + 32LL * stride), n_win, stride) : n_win;
        fetch(desc, scal, off, K, n_win);
    }
};

// stages of the shared-memory ring (measured): K <= 2: two stages and two CTAs per SM; K = 3: three stages, one CTA per SM
#define RVE_PIPE_STAGES(K) ((K) == 3 ? 3 : 2)
#define RVE_UPDP_MINB(K) ((K) == 3 ? 1 : 2)
#ifndef RVE_RESTRICT_SUB
#define RVE_RESTRICT_SUB 2      // threads sharing the (row, corner) list of one level-1 node in the restriction (2 or 4)
#endif
#define RVE_UPDP_THREADS(K) ((K) == 3 ? 384 : 256)

// shared-memory footprint of one stage (bytes); cn_pad = level-1 nodes per window rounded up to 4, + 4
__host__ __device__ inline size_t updp_stage_bytes(int K, int cn_pad, bool precond) {
    size_t b = 4 * (size_t)RVE_WIN * K * 16 + (size_t)RVE_WIN * 16;                     // r p q x, dinv
    if (precond) b += (size_t)RVE_WIN * 16 + (4 * (size_t)RVE_WIN + 8) * 2 + 2 * (size_t)cn_pad * 4;   // rowinfo, ent, beg, node
    return (b + 127) & ~(size_t)127;
}

// x += alpha p ; r -= alpha q ; per-window partials of r.r and r.Dinv r ; rc += Z^T r.
// Same arithmetic as k_update (rve_kernels.cuh), persistent + bulk-copy pipelined; launch with RVE_UPDP_THREADS(K) + 32 threads.
template <int K>
__global__ void __launch_bounds__(RVE_UPDP_THREADS(K) + 32, RVE_UPDP_MINB(K))
k_update_p(WinTab wt, const WinDesc* __restrict__ desc, int n_win, const double2* __restrict__ p, const double2* __restrict__ q,
           double2* __restrict__ x, double2* __restrict__ r, const float4* __restrict__ dinv32, const RowInfo* __restrict__ rowinfo,
           const int* __restrict__ cn_node, const unsigned* __restrict__ cn_beg, const uint16_t* __restrict__ cn_ent,
           float* __restrict__ rc1, int G, double* __restrict__ part, const double* __restrict__ win_scal, int init, int precond, int cn_pad) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    constexpr int TPB = RVE_UPDP_THREADS(K);              // consumer threads
    constexpr int NST = RVE_PIPE_STAGES(K);
    __shared__ __align__(8) uint64_t s_full[NST], s_empty[NST];
    __shared__ WinDesc s_desc[NST];
    __shared__ int s_win[NST];
    __shared__ double s_alpha[NST][4];
    __shared__ double s_a[TPB], s_b[TPB];
    __shared__ double s_fin[8];
    const int tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;
    const size_t stage_bytes = updp_stage_bytes(K, cn_pad, precond != 0);
    const int stride = (int)gridDim.x;

    // stage layout
    auto st_r = [&](int s) { return (double2*)(s_raw + (size_t)s * stage_bytes); };
    auto st_p = [&](int s) { return st_r(s) + (size_t)RVE_WIN * K; };
    auto st_q = [&](int s) { return st_r(s) + 2 * (size_t)RVE_WIN * K; };
    auto st_x = [&](int s) { return st_r(s) + 3 * (size_t)RVE_WIN * K; };
    auto st_d = [&](int s) { return (float4*)(st_r(s) + 4 * (size_t)RVE_WIN * K); };
    auto st_ri = [&](int s) { return (RowInfo*)(st_d(s) + RVE_WIN); };
    auto st_ent = [&](int s) { return (uint16_t*)(st_ri(s) + RVE_WIN); };
    auto st_beg = [&](int s) { return (unsigned*)(st_ent(s) + 4 * RVE_WIN + 8); };
    auto st_node = [&](int s) { return (int*)(st_beg(s) + cn_pad); };

    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 1); }
        mbar_fence_init();
    }
    __syncthreads();

    if (wv == TPB / 32) {
        // ---------------- producer warp ----------------
        WinCursor cur;
        cur.init(wt.win_active, desc, win_scal, 0, K, n_win, stride, (int)blockIdx.x);
        for (int it = 0;; ++it) {
            const int s = it % NST;
            cur.prefetch(wt.win_active, desc, win_scal, 0, K, n_win, stride);
            const int w = cur.w;
            const WinDesc d = cur.d;
            if (it >= NST) mbar_wait(&s_empty[s], ((it / NST) - 1) & 1);
            if (lane < K) s_alpha[s][lane] = cur.sv;
            __syncwarp();
            if (lane == 0) {
                s_win[s] = w;
                s_desc[s] = d;
                if (w >= n_win) {
                    mbar_arrive(&s_full[s]);                 // end marker: consumers wake up and leave
                } else {
                    fence_proxy_async();
                    const unsigned vb = (unsigned)d.n * K * 16, rb = (unsigned)d.n * 16;
                    unsigned tot = vb + rb + (init ? 0u : 3u * vb);
                    unsigned eb = 0, bb = 0, nb = 0;
                    if (precond) {
                        eb = ((d.nent + 7u) & ~7u) * 2u; bb = (((unsigned)d.ncn + 1u + 3u) & ~3u) * 4u; nb = (((unsigned)d.ncn + 3u) & ~3u) * 4u;
                        tot += rb + eb + bb + nb;
                    }
                    mbar_expect_tx(&s_full[s], tot);
                    const size_t base = (size_t)d.row0 * K;
                    bulk_g2s(st_r(s), r + base, vb, &s_full[s]);
                    if (!init) {
                        bulk_g2s(st_p(s), p + base, vb, &s_full[s]);
                        bulk_g2s(st_q(s), q + base, vb, &s_full[s]);
                        bulk_g2s(st_x(s), x + base, vb, &s_full[s]);
                    }
                    bulk_g2s(st_d(s), dinv32 + d.row0, rb, &s_full[s]);
                    if (precond) {
                        bulk_g2s(st_ri(s), rowinfo + d.row0, rb, &s_full[s]);
                        if (eb) bulk_g2s(st_ent(s), cn_ent + d.ent0, eb, &s_full[s]);
                        bulk_g2s(st_beg(s), cn_beg + d.cn0, bb, &s_full[s]);
                        if (nb) bulk_g2s(st_node(s), cn_node + d.cn0, nb, &s_full[s]);
                    }
                }
            }
            if (w >= n_win) break;
            cur.advance(wt.win_active, desc, win_scal, 0, K, n_win, stride);
        }
        return;
    }

    // ---------------- consumer warps ----------------
    const int k = tid % K;
    for (int it = 0;; ++it) {
        const int stage = it % NST;
        mbar_wait(&s_full[stage], (it / NST) & 1);
        const int win = s_win[stage];
        if (win >= n_win) break;
        const WinDesc d = s_desc[stage];
        const int sys = d.sys, row0 = d.row0, n = d.n, ncn = d.ncn;
        const double alpha = s_alpha[stage][k];
        double2* sr = st_r(stage);
        double rr = 0.0, rdr = 0.0;
        {
            const double2* sp = st_p(stage); const double2* sq = st_q(stage); const double2* sx = st_x(stage);
            const float4* sd = st_d(stage);
            const size_t base = (size_t)row0 * K;
            for (int e = tid; e < n * K; e += TPB) {
                double2 rw = sr[e];
                if (!init) {
                    double2 xw = sx[e];
                    const double2 pv = sp[e], qv = sq[e];
                    xw.x += alpha * pv.x; xw.y += alpha * pv.y;
                    rw.x -= alpha * qv.x; rw.y -= alpha * qv.y;
                    x[base + e] = xw; r[base + e] = rw;
                    sr[e] = rw;
                }
                const float4 dv = sd[e / K];
                rr += rw.x * rw.x + rw.y * rw.y;
                rdr += rw.x * ((double)dv.x * rw.x + (double)dv.y * rw.y) + rw.y * ((double)dv.z * rw.x + (double)dv.w * rw.y);
            }
        }
        s_a[tid] = rr; s_b[tid] = rdr;
        if (tid < 8) s_fin[tid] = 0.0;
        consumer_sync(TPB);
        if (wv < 2 * K) {
            const int kk = wv >> 1;
            const double* src = (wv & 1) ? s_b : s_a;
            double acc = 0.0;
            for (int m = lane; m < TPB / K; m += 32) acc += src[kk + K * m];
            acc = warp_sum(acc);
            if (lane == 0) s_fin[wv] = acc;
        }
        if (precond) {
            const RowInfo* sri = st_ri(stage);
            const uint16_t* sent = st_ent(stage);
            const unsigned* sbeg = st_beg(stage);
            const int* snode = st_node(stage);
            const unsigned ent0 = d.ent0;
            constexpr int KT = (K == 2) ? 2 : 1;              // rhs per task
            constexpr int TPN = K / KT;                       // tasks per node
            for (int t = tid; t < ((ncn * TPN * RVE_RESTRICT_SUB + 31) & ~31); t += TPB) {
                const int task = t / RVE_RESTRICT_SUB, sub = t % RVE_RESTRICT_SUB;
                const int jn = task / TPN, k0 = (task % TPN) * KT;
                double a[KT][2];
#pragma unroll
                for (int kk = 0; kk < KT; ++kk) { a[kk][0] = 0.0; a[kk][1] = 0.0; }
                if (jn < ncn) {
                    const unsigned b0 = sbeg[jn] - ent0, b1 = sbeg[jn + 1] - ent0;
                    for (unsigned e = b0 + sub; e < b1; e += RVE_RESTRICT_SUB) {
                        const int en = sent[e];
                        const int rl = en >> 2, c = en & 3;
                        const RowInfo ri = sri[rl];
                        const double wgt = (double)((c & 1) ? ri.s : 1.0f - ri.s) * (double)((c & 2) ? ri.t : 1.0f - ri.t);
#pragma unroll
                        for (int kk = 0; kk < KT; ++kk) {
                            const double2 v = sr[rl * K + k0 + kk];
                            a[kk][0] += wgt * v.x; a[kk][1] += wgt * v.y;
                        }
                    }
                }
#pragma unroll
                for (int kk = 0; kk < KT; ++kk)
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        a[kk][c] += __shfl_xor_sync(0xffffffffu, a[kk][c], 1);
                        if (RVE_RESTRICT_SUB == 4) a[kk][c] += __shfl_xor_sync(0xffffffffu, a[kk][c], 2);
                    }
                if (jn < ncn && sub == 0) {
                    float* dd = rc1 + (((size_t)sys * G + snode[jn]) * K + k0) * 2;
                    if (KT == 2) atomicAdd((float4*)dd, make_float4((float)a[0][0], (float)a[0][1], (float)a[KT - 1][0], (float)a[KT - 1][1]));
                    else atomicAdd((float2*)dd, make_float2((float)a[0][0], (float)a[0][1]));
                }
            }
        }
        consumer_sync(TPB);                                   // the stage has been read for the last time
        if (tid == 0) mbar_arrive(&s_empty[stage]);
        if (tid < 8) part[(size_t)win * 8 + tid] = s_fin[tid];    // [rhs][rr, r.Dinv r]; k_fin_update adds them per system
    }
}

// ------------------------------------------------------------------------------------------------------------------
// p = Dinv r + Z xc1 + beta p, persistent + pipelined.  Three stages: while window i is computed, the bulk copies of
// window i+2 (r, p, dinv, rowinfo, level-1 node list) are in flight (producer warp) and the level-1 values of window
// i+1 are gathered by the consumers with 8-byte cp.async (its node list landed one window ago), so neither the streams
// nor the dependent gather are on the critical path.  Launch with RVE_UPDP_THREADS(K) + 32 threads.
// ------------------------------------------------------------------------------------------------------------------
#define RVE_DIRP_STAGES 3
#define RVE_DIRP_MINB(K) 2             // two CTAs per SM for every K (3 stages of r, p, dinv, rowinfo fit twice)
__host__ __device__ inline size_t dirp_stage_bytes(int K, int cn_pad, bool precond) {
    size_t b = 2 * (size_t)RVE_WIN * K * 16 + (size_t)RVE_WIN * 16;                                   // r p, dinv
    if (precond) b += (size_t)RVE_WIN * 16 + (size_t)cn_pad * 4 + (size_t)cn_pad * K * 8;             // rowinfo, node list, xc (float2)
    return (b + 127) & ~(size_t)127;
}

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

template <int K>
__global__ void __launch_bounds__(RVE_UPDP_THREADS(K) + 32, RVE_DIRP_MINB(K))
k_direction_p(WinTab wt, const WinDesc* __restrict__ desc, int n_win, double2* __restrict__ p, const double2* __restrict__ r,
              const float4* __restrict__ dinv32, const RowInfo* __restrict__ rowinfo, const int* __restrict__ cn_node,
              const cvec* __restrict__ xc1, int G, const double* __restrict__ win_scal, int precond, int cn_pad) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    constexpr int TPB = RVE_UPDP_THREADS(K);
    constexpr int NST = RVE_DIRP_STAGES;
    __shared__ __align__(8) uint64_t s_full[NST], s_empty[NST];
    __shared__ WinDesc s_desc[NST];
    __shared__ int s_win[NST];
    __shared__ double s_beta[NST][4];
    const int tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;
    const size_t stage_bytes = dirp_stage_bytes(K, cn_pad, precond != 0);
    const int stride = (int)gridDim.x;
    auto st_r = [&](int s) { return (double2*)(s_raw + (size_t)s * stage_bytes); };
    auto st_p = [&](int s) { return st_r(s) + (size_t)RVE_WIN * K; };
    auto st_d = [&](int s) { return (float4*)(st_r(s) + 2 * (size_t)RVE_WIN * K); };
    auto st_ri = [&](int s) { return (RowInfo*)(st_d(s) + RVE_WIN); };
    auto st_node = [&](int s) { return (int*)(st_ri(s) + RVE_WIN); };
    auto st_xc = [&](int s) { return (cvec*)(st_node(s) + cn_pad); };

    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 1); }
        mbar_fence_init();
    }
    __syncthreads();

    if (wv == TPB / 32) {
        // ---------------- producer warp ----------------
        WinCursor cur;
        cur.init(wt.win_active, desc, win_scal, 4, K, n_win, stride, (int)blockIdx.x);
        for (int it = 0;; ++it) {
            const int s = it % NST;
            cur.prefetch(wt.win_active, desc, win_scal, 4, K, n_win, stride);
            const int w = cur.w;
            const WinDesc d = cur.d;
            if (it >= NST) mbar_wait(&s_empty[s], ((it / NST) - 1) & 1);
            if (lane < K) s_beta[s][lane] = cur.sv;
            __syncwarp();
            if (lane == 0) {
                s_win[s] = w;
                s_desc[s] = d;
                if (w >= n_win) {
                    mbar_arrive(&s_full[s]);
                } else {
                    fence_proxy_async();
                    const unsigned vb = (unsigned)d.n * K * 16, rb = (unsigned)d.n * 16;
                    unsigned tot = 2u * vb + rb, nb = 0;
                    if (precond) { nb = (((unsigned)d.ncn + 3u) & ~3u) * 4u; tot += rb + nb; }
                    mbar_expect_tx(&s_full[s], tot);
                    const size_t base = (size_t)d.row0 * K;
                    bulk_g2s(st_r(s), r + base, vb, &s_full[s]);
                    bulk_g2s(st_p(s), p + base, vb, &s_full[s]);
                    bulk_g2s(st_d(s), dinv32 + d.row0, rb, &s_full[s]);
                    if (precond) {
                        bulk_g2s(st_ri(s), rowinfo + d.row0, rb, &s_full[s]);
                        if (nb) bulk_g2s(st_node(s), cn_node + d.cn0, nb, &s_full[s]);
                    }
                }
            }
            if (w >= n_win) break;
            cur.advance(wt.win_active, desc, win_scal, 4, K, n_win, stride);
        }
        return;
    }

    // ---------------- consumer warps ----------------
    // gather of the level-1 values of the window in stage s (its bulk copies must have landed)
    auto gather = [&](int s) {
        if (precond) {
            const WinDesc d = s_desc[s];
            const int* sn = st_node(s);
            cvec* sx = st_xc(s);
            const cvec* src = xc1 + (size_t)d.sys * G * K;
            for (int t = tid; t < d.ncn * K; t += TPB) cp_async8(sx + t, src + (size_t)sn[t / K] * K + (t % K));
        }
        cp_async_commit();
    };
    const int k = tid % K;
    mbar_wait(&s_full[0], 0);
    bool have = s_win[0] < n_win;
    if (have) gather(0);
    for (int it = 0; have; ++it) {
        const int stage = it % NST, nxt = (it + 1) % NST;
        // next window: wait for its streams, start its gather
        mbar_wait(&s_full[nxt], ((it + 1) / NST) & 1);
        const bool have_next = s_win[nxt] < n_win;
        if (have_next) gather(nxt); else cp_async_commit();
        cp_async_wait_group<1>();                             // this window's gather (committed one iteration ago)
        consumer_sync(TPB);
        const WinDesc d = s_desc[stage];
        const int row0 = d.row0, n = d.n;
        const double beta = s_beta[stage][k];
        const double2* sr = st_r(stage); const double2* sp = st_p(stage);
        const float4* sd = st_d(stage); const RowInfo* sri = st_ri(stage); const cvec* sx = st_xc(stage);
        const size_t base = (size_t)row0 * K;
        for (int e = tid; e < n * K; e += TPB) {
            const int rl = e / K;
            const double2 rv = sr[e], pv = sp[e];
            const float4 dv = sd[rl];
            double z0 = (double)dv.x * rv.x + (double)dv.y * rv.y, z1 = (double)dv.z * rv.x + (double)dv.w * rv.y;
            if (precond) {
                const RowInfo ri = sri[rl];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int slot = ri.slot[c];
                    if (slot != 0xffff) {
                        const double wgt = (double)((c & 1) ? ri.s : 1.0f - ri.s) * (double)((c & 2) ? ri.t : 1.0f - ri.t);
                        const cvec v = sx[slot * K + k];
                        z0 += wgt * (double)v.x; z1 += wgt * (double)v.y;
                    }
                }
            }
            p[base + e] = make_double2(z0 + beta * pv.x, z1 + beta * pv.y);
        }
        consumer_sync(TPB);                                   // the stage has been read for the last time
        if (tid == 0) mbar_arrive(&s_empty[stage]);
        have = have_next;
    }
}
