// rve_kernels.cuh — sm_100a kernels of the batched RVE solver.
//
// Data layout (all systems of a batch concatenated; "row" = active P2 node = 2x2 block row):
//   node_xy[N][2], node_row[N] (global row or -1), node_bnd[N]; elem_node[E][6] (global node ids),
//   elem_reg[E], elem_sys[E];  row_ptr[R+1] (uint32 block offsets), bcol[NB] (global row ids),
//   bval[NB][4] (K00 K01 K10 K11), dinv[R][4], row_xy[R][2], wmean[R].
//   PCG vectors: p[R][4][2] (padded: one 64-byte line segment per gathered node),
//                x,r,q,b[R][K][2] compact.
//   Coarse level l (structured (ncx+1)x(ncy+1) grid per system, 25-point block stencil):
//                A[sys][slot 25][4][G], dinv[sys][G][4], vectors [sys][G][K][2].
// Rows are processed in chunks of <= RVE_CH rows that never straddle a system; per-chunk partial
// dot products are combined in a fixed order by the last CTA of each system (deterministic).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rve_b200.h"
#include "rve_math.cuh"

#define RVE_CH 64          // rows per chunk (CTA of 256 threads: 8 warps x 8 rows, or 64 rows x 4 rhs)
#define RVE_MAXLEV 8
#define RVE_OMEGA 0.8      // damped block-Jacobi smoother on the coarse levels

struct Mat8 { double v[8]; };  // [4][2] lambda*, mu per region

struct Grid {
    int ncx, ncy, nx, ny, G;
    int lox, hix, loy, hiy;  // valid node index range (inclusive)
    int wrap;                // periodic wrap
};

struct Levels {
    int L;          // number of coarse levels
    int K;          // rhs count
    Grid g[RVE_MAXLEV];
    double* A[RVE_MAXLEV];      // [n_sys][25][4][G]
    double* dinv[RVE_MAXLEV];   // [n_sys][G][4]
    double* rc[RVE_MAXLEV];     // [n_sys][G][K][2]
    double* xc[RVE_MAXLEV];
    double* tc[RVE_MAXLEV];
};

struct ChunkTab {
    const int* chunk_sys;   // [n_chunks]
    const int* chunk_row0;  // global first row
    const int* chunk_n;     // rows in the chunk
    const int* sys_chunk0;  // [n_sys+1]
};

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

__device__ __forceinline__ bool grid_valid(const Grid& g, int i, int j) {
    return i >= g.lox && i <= g.hix && j >= g.loy && j <= g.hiy;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------
// (1) numeric assembly: warp per element, 21 lanes compute the upper-triangular 2x2 blocks and
//     scatter them (and their transposes) with fp64 atomics into the node-block CSR.
//     Connectivity / coordinates are staged through warp shuffles, the material table lives in
//     shared memory.  Replaces mp.block_assemble(a) (micro_model.py:60).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int find_block(const unsigned* __restrict__ row_ptr, const int* __restrict__ bcol,
                                          int row, int col) {
    unsigned lo = row_ptr[row], hi = row_ptr[row + 1];
    while (lo < hi) {
        unsigned mid = (lo + hi) >> 1;
        int c = bcol[mid];
        if (c < col) lo = mid + 1; else hi = mid;
    }
    return (int)lo;  // caller guarantees presence
}

__global__ void __launch_bounds__(256)
k_assemble(int n_elem, const int* __restrict__ elem_node, const unsigned char* __restrict__ elem_reg,
           const double* __restrict__ node_xy, const int* __restrict__ node_row,
           const unsigned* __restrict__ row_ptr, const int* __restrict__ bcol, double* __restrict__ bval,
           double* __restrict__ wmean, Mat8 mat) {
    __shared__ double s_mat[8];
    if (threadIdx.x < 8) s_mat[threadIdx.x] = mat.v[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (e >= n_elem) return;
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
    int a = 0, b = 0;
    pair_ab(lane < 21 ? lane : 0, a, b);
    const int ra = __shfl_sync(0xffffffffu, row, a);
    const int rb = __shfl_sync(0xffffffffu, row, b);
    if (lane < 21 && ra >= 0 && rb >= 0) {
        double K[4];
        p2_stiff_block(g, a, b, lam, mu, K);
        double* d = bval + 4 * (size_t)find_block(row_ptr, bcol, ra, rb);
        atomicAdd(d + 0, K[0]); atomicAdd(d + 1, K[1]); atomicAdd(d + 2, K[2]); atomicAdd(d + 3, K[3]);
        if (a != b) {
            double* dt = bval + 4 * (size_t)find_block(row_ptr, bcol, rb, ra);
            atomicAdd(dt + 0, K[0]); atomicAdd(dt + 1, K[2]); atomicAdd(dt + 2, K[1]); atomicAdd(dt + 3, K[3]);
        }
    }
    // int phi_a dx: A/3 for the mid-edge functions (mean-value constraint of 'per' / 'MR')
    if (lane >= 3 && lane < 6 && row >= 0) atomicAdd(wmean + row, g.area * (1.0 / 3.0));
}

// 2x2 inverse of the diagonal blocks (block-Jacobi)
__global__ void k_dinv(int n_rows, const unsigned* __restrict__ row_ptr, const int* __restrict__ bcol,
                       const double* __restrict__ bval, double* __restrict__ dinv) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const double* d = bval + 4 * (size_t)find_block(row_ptr, bcol, r, r);
    double a = d[0], b = d[1], c = d[2], e = d[3];
    double det = a * e - b * c;
    double inv = det != 0.0 ? 1.0 / det : 0.0;
    dinv[4 * (size_t)r + 0] = e * inv; dinv[4 * (size_t)r + 1] = -b * inv;
    dinv[4 * (size_t)r + 2] = -c * inv; dinv[4 * (size_t)r + 3] = a * inv;
}

// ---------------------------------------------------------------------------------------------
// (2) level-1 Galerkin operator A1 = Z^T A Z, element by element: warp per element, K_e and the
//     bilinear weights staged in shared memory, 9x9 window of coarse nodes, atomics into the
//     25-slot stencil arrays.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_galerkin1(int n_elem, const int* __restrict__ elem_node, const unsigned char* __restrict__ elem_reg,
            const int* __restrict__ elem_sys, const double* __restrict__ node_xy,
            const int* __restrict__ node_row, const double* __restrict__ box, Grid g1,
            double* __restrict__ A1, Mat8 mat, int* __restrict__ err_flag) {
    __shared__ double s_mat[8];
    __shared__ double s_K[8][36][4];
    __shared__ double s_W[8][6][9];
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
    if (lane < 21) {
        int a, b;
        pair_ab(lane, a, b);
        double K[4];
        p2_stiff_block(g, a, b, lam, mu, K);
        s_K[w][a * 6 + b][0] = K[0]; s_K[w][a * 6 + b][1] = K[1];
        s_K[w][a * 6 + b][2] = K[2]; s_K[w][a * 6 + b][3] = K[3];
        if (a != b) {
            s_K[w][b * 6 + a][0] = K[0]; s_K[w][b * 6 + a][1] = K[2];
            s_K[w][b * 6 + a][2] = K[1]; s_K[w][b * 6 + a][3] = K[3];
        }
    }
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
        int oi = ci - imin, oj = cj - jmin;
        if (oi > 1 || oj > 1) {
            atomicExch(err_flag, 1);  // element spans more than two coarse cells: grid too fine
        } else if (row >= 0) {
            s_W[w][lane][oj * 3 + oi] = (1 - s) * (1 - t);
            s_W[w][lane][oj * 3 + oi + 1] = s * (1 - t);
            s_W[w][lane][(oj + 1) * 3 + oi] = (1 - s) * t;
            s_W[w][lane][(oj + 1) * 3 + oi + 1] = s * t;
        }
    }
    __syncwarp();
    double* Asys = A1 + (size_t)sys * 100 * g1.G;
    for (int idx = lane; idx < 81; idx += 32) {
        const int I = idx / 9, J = idx % 9;
        double wI = 0, wJ = 0;
#pragma unroll
        for (int a = 0; a < 6; ++a) { wI += s_W[w][a][I]; wJ += s_W[w][a][J]; }
        if (wI == 0.0 || wJ == 0.0) continue;
        int Ii = imin + I % 3, Ij = jmin + I / 3, Ji = imin + J % 3, Jj = jmin + J / 3;
        int nI, nJ;
        if (!grid_node(g1, Ii, Ij, nI) || !grid_node(g1, Ji, Jj, nJ)) continue;
        double acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
        for (int a = 0; a < 6; ++a) {
            double wa = s_W[w][a][I];
            if (wa == 0.0) continue;
            for (int b = 0; b < 6; ++b) {
                double wb = s_W[w][b][J];
                if (wb == 0.0) continue;
                double ww = wa * wb;
                const double* k4 = s_K[w][a * 6 + b];
                acc0 += ww * k4[0]; acc1 += ww * k4[1]; acc2 += ww * k4[2]; acc3 += ww * k4[3];
            }
        }
        const int slot = (Jj - Ij + 2) * 5 + (Ji - Ii + 2);
        double* d = Asys + (size_t)slot * 4 * g1.G + nI;
        atomicAdd(d, acc0); atomicAdd(d + g1.G, acc1);
        atomicAdd(d + 2 * (size_t)g1.G, acc2); atomicAdd(d + 3 * (size_t)g1.G, acc3);
    }
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
// (3) PCG kernels
// ---------------------------------------------------------------------------------------------

// Deterministic combination of per-chunk partials by the last CTA of a system.
// part layout: [chunk][4][NQ]; result written to out[rq][q] (shared) for rq<4, q<NQ.
template <int NQ>
__device__ __forceinline__ bool last_block_sum(double* part, int* cnt, int sys, int chunk, int c0, int c1,
                                               const double* my /*[4*NQ] valid in thread 0..4*NQ-1? */,
                                               double (*s_out)[NQ]) {
    __shared__ int s_last;
    __shared__ double s_red[256];
    const int tid = threadIdx.x;
    if (tid < 4 * NQ) part[(size_t)chunk * 4 * NQ + tid] = my[0];
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        int old = atomicAdd(cnt + sys, 1);
        s_last = (old == (c1 - c0 - 1));
        if (s_last) cnt[sys] = 0;
    }
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    // 256 threads: value v = tid % (4*NQ), lane group = tid / (4*NQ)
    const int nv = 4 * NQ;
    const int grp = tid / nv, v = tid % nv, ngrp = 256 / nv;
    double acc = 0.0;
    if (grp < ngrp)
        for (int c = c0 + grp; c < c1; c += ngrp) acc += __ldcg(part + (size_t)c * nv + v);
    s_red[tid] = (grp < ngrp) ? acc : 0.0;
    __syncthreads();
    if (tid < nv) {
        double t = 0.0;
        for (int gq = 0; gq < ngrp; ++gq) t += s_red[gq * nv + tid];
        s_out[tid / NQ][tid % NQ] = t;
    }
    __syncthreads();
    return true;
}

// q = A p, pq = p.q  — warp per block row, lane quad per 2x2 block, lane (lane&3) owns one rhs.
// Per 8-block pass a warp issues: 2 LDG.128 on a contiguous 256 B of bval, one LDG on 32 B of bcol
// and one LDG.128 gather that touches exactly one 64-byte segment of p per block.
__global__ void __launch_bounds__(256)
k_spmv_dot(const unsigned* __restrict__ row_ptr, const int* __restrict__ bcol,
           const double2* __restrict__ bval, const double2* __restrict__ p, double2* __restrict__ q,
           int K, ChunkTab ct, const int* __restrict__ active, double* part, int* cnt, double* sc) {
    const int chunk = blockIdx.x;
    const int sys = ct.chunk_sys[chunk];
    if (!active[sys]) return;
    const int row0 = ct.chunk_row0[chunk], nrow = ct.chunk_n[chunk];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int grp = lane >> 2, rq = lane & 3;
    double dot = 0.0;
    for (int lr = w; lr < nrow; lr += 8) {
        const int row = row0 + lr;
        const unsigned beg = row_ptr[row], end = row_ptr[row + 1];
        double y0 = 0.0, y1 = 0.0;
        unsigned b = beg + grp;
        // two passes in flight
        for (; b + 8 < end; b += 16) {
            const int c0 = bcol[b], c1 = bcol[b + 8];
            const double2 k00 = bval[2 * (size_t)b], k01 = bval[2 * (size_t)b + 1];
            const double2 k10 = bval[2 * (size_t)(b + 8)], k11 = bval[2 * (size_t)(b + 8) + 1];
            const double2 xa = p[(size_t)c0 * 4 + rq], xb = p[(size_t)c1 * 4 + rq];
            y0 += k00.x * xa.x + k00.y * xa.y; y1 += k01.x * xa.x + k01.y * xa.y;
            y0 += k10.x * xb.x + k10.y * xb.y; y1 += k11.x * xb.x + k11.y * xb.y;
        }
        if (b < end) {
            const int c0 = bcol[b];
            const double2 k00 = bval[2 * (size_t)b], k01 = bval[2 * (size_t)b + 1];
            const double2 xa = p[(size_t)c0 * 4 + rq];
            y0 += k00.x * xa.x + k00.y * xa.y; y1 += k01.x * xa.x + k01.y * xa.y;
        }
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
            y0 += __shfl_xor_sync(0xffffffffu, y0, o);
            y1 += __shfl_xor_sync(0xffffffffu, y1, o);
        }
        if (grp == 0 && rq < K) {
            q[(size_t)row * K + rq] = make_double2(y0, y1);
            const double2 pv = p[(size_t)row * 4 + rq];
            dot += y0 * pv.x + y1 * pv.y;
        }
    }
    // CTA reduction of the 4 per-rhs dots in a fixed order
    __shared__ double s_dot[8][4];
    __shared__ double s_out[4][1];
    if (grp == 0) s_dot[w][rq] = dot;
    __syncthreads();
    double my = 0.0;
    if (threadIdx.x < 4) {
        for (int i = 0; i < 8; ++i) my += s_dot[i][threadIdx.x];
    }
    const int c0 = ct.sys_chunk0[sys], c1 = ct.sys_chunk0[sys + 1];
    if (last_block_sum<1>(part, cnt, sys, chunk, c0, c1, &my, s_out)) {
        if (threadIdx.x < 4) {
            double* s = sc + ((size_t)sys * 4 + threadIdx.x) * SC_N;
            const double pq = s_out[threadIdx.x][0];
            s[SC_PQ] = pq;
            s[SC_ALPHA] = (pq > 0.0) ? s[SC_RZ] / pq : 0.0;
        }
    }
}

// x += alpha p ; r -= alpha q ; partial rr and r.Dinv r ; convergence test by the last CTA.
// thread = (row in chunk, rhs).
__global__ void __launch_bounds__(256)
k_update(const double2* __restrict__ p, const double2* __restrict__ q, double2* __restrict__ x,
         double2* __restrict__ r, const double* __restrict__ dinv, int K, ChunkTab ct, int* active,
         double* part, int* cnt, double* sc, int init, int iter, double rtol2, int precond,
         int* iters, int* n_active) {
    const int chunk = blockIdx.x;
    const int sys = ct.chunk_sys[chunk];
    if (!active[sys]) return;
    const int row0 = ct.chunk_row0[chunk], nrow = ct.chunk_n[chunk];
    const int lr = threadIdx.x >> 2, rq = threadIdx.x & 3;
    double rr = 0.0, rdr = 0.0;
    if (lr < nrow && rq < K) {
        const int row = row0 + lr;
        const double alpha = sc[((size_t)sys * 4 + rq) * SC_N + SC_ALPHA];
        const size_t i = (size_t)row * K + rq;
        double2 rv = r[i];
        if (!init) {
            const double2 pv = p[(size_t)row * 4 + rq], qv = q[i];
            double2 xv = x[i];
            xv.x += alpha * pv.x; xv.y += alpha * pv.y;
            rv.x -= alpha * qv.x; rv.y -= alpha * qv.y;
            x[i] = xv; r[i] = rv;
        }
        const double* d = dinv + 4 * (size_t)row;
        rr = rv.x * rv.x + rv.y * rv.y;
        rdr = rv.x * (d[0] * rv.x + d[1] * rv.y) + rv.y * (d[2] * rv.x + d[3] * rv.y);
    }
    __shared__ double s_a[256], s_b[256];
    __shared__ double s_out[4][2];
    s_a[threadIdx.x] = rr; s_b[threadIdx.x] = rdr;
    __syncthreads();
    for (int st = 128; st >= 4; st >>= 1) {
        if (threadIdx.x < st) { s_a[threadIdx.x] += s_a[threadIdx.x + st]; s_b[threadIdx.x] += s_b[threadIdx.x + st]; }
        __syncthreads();
    }
    // thread t<8 publishes value (rq = t/2, which = t%2)
    double my = 0.0;
    if (threadIdx.x < 8) my = (threadIdx.x & 1) ? s_b[threadIdx.x >> 1] : s_a[threadIdx.x >> 1];
    const int c0 = ct.sys_chunk0[sys], c1 = ct.sys_chunk0[sys + 1];
    if (last_block_sum<2>(part, cnt, sys, chunk, c0, c1, &my, s_out)) {
        if (threadIdx.x == 0) {
            bool done = true;
            for (int k = 0; k < K; ++k) {
                double* s = sc + ((size_t)sys * 4 + k) * SC_N;
                const double vrr = s_out[k][0], vrdr = s_out[k][1];
                if (init) s[SC_BB] = vrr;
                s[SC_RR] = vrr; s[SC_RDR] = vrdr;
                if (!(vrr <= rtol2 * s[SC_BB] || vrr <= 1e-28 * s[SC_BREF])) done = false;
                if (precond == 0) {  // block Jacobi: z = Dinv r, finish the scalar recurrences here
                    const double rzold = s[SC_RZ];
                    s[SC_RZOLD] = rzold;
                    s[SC_RZ] = vrdr;
                    s[SC_BETA] = (!init && rzold > 0.0) ? vrdr / rzold : 0.0;
                }
            }
            if (done) {
                active[sys] = 0;
                iters[sys] = init ? 0 : iter + 1;
                atomicSub(n_active, 1);
            }
        }
    }
}

// Coarse V-cycle: ONE CTA per system runs restriction from the fine residual, all structured
// levels (25-point block stencils, damped block-Jacobi smoothing) and the coarse part of r.z.
__device__ __forceinline__ void stencil_apply(const Grid& g, const double* __restrict__ A,
                                              const double* x, int K, int i, int j, double* out) {
    const int idx = j * g.nx + i;
    for (int dj = -2; dj <= 2; ++dj)
        for (int di = -2; di <= 2; ++di) {
            int nb;
            if (!grid_node(g, i + di, j + dj, nb)) continue;
            const double* a = A + (size_t)((dj + 2) * 5 + (di + 2)) * 4 * g.G + idx;
            const double a00 = a[0], a01 = a[g.G], a10 = a[2 * (size_t)g.G], a11 = a[3 * (size_t)g.G];
            if (a00 == 0.0 && a01 == 0.0 && a10 == 0.0 && a11 == 0.0) continue;
            const double* xn = x + (size_t)nb * K * 2;
#pragma unroll
            for (int k = 0; k < RVE_MAX_RHS; ++k)
                if (k < K) {
                    const double x0 = xn[2 * k], x1 = xn[2 * k + 1];
                    out[2 * k] += a00 * x0 + a01 * x1;
                    out[2 * k + 1] += a10 * x0 + a11 * x1;
                }
        }
}

__global__ void __launch_bounds__(256)
k_vcycle(Levels LV, const double* __restrict__ r, const double* __restrict__ row_xy,
         const int* __restrict__ row_base, const int* __restrict__ cell_ptr, const double* __restrict__ box,
         const int* __restrict__ active, double* sc, int first) {
    const int sys = blockIdx.x;
    if (!active[sys]) return;
    const int K = LV.K, tid = threadIdx.x, nth = blockDim.x;
    const double om = RVE_OMEGA;
    // --- restrict fine residual to level 1:  rc = Z^T r
    {
        const Grid g = LV.g[0];
        double* rc = LV.rc[0] + (size_t)sys * g.G * K * 2;
        const int* cp = cell_ptr + (size_t)sys * (g.ncx * g.ncy + 1);
        const int rb = row_base[sys];
        const double bx0 = box[4 * sys], by0 = box[4 * sys + 1];
        const double invHx = g.ncx / box[4 * sys + 2], invHy = g.ncy / box[4 * sys + 3];
        for (int idx = tid; idx < g.G; idx += nth) {
            const int i = idx % g.nx, j = idx / g.nx;
            if (!grid_valid(g, i, j)) continue;
            double acc[2 * RVE_MAX_RHS];
#pragma unroll
            for (int k = 0; k < 2 * RVE_MAX_RHS; ++k) acc[k] = 0.0;
            for (int dj = -1; dj <= 0; ++dj)
                for (int di = -1; di <= 0; ++di) {
                    int ci = i + di, cj = j + dj;
                    if (g.wrap) { ci = (ci + g.ncx) % g.ncx; cj = (cj + g.ncy) % g.ncy; }
                    else if (ci < 0 || cj < 0 || ci >= g.ncx || cj >= g.ncy) continue;
                    const int c = cj * g.ncx + ci;
                    for (int rr = cp[c]; rr < cp[c + 1]; ++rr) {
                        const size_t row = (size_t)rb + rr;
                        int qi, qj; double s, t;
                        coarse_cell(row_xy[2 * row], row_xy[2 * row + 1], bx0, by0, invHx, invHy, g.ncx, g.ncy, qi, qj, s, t);
                        const double wgt = (di ? s : 1.0 - s) * (dj ? t : 1.0 - t);
                        const double* rv = r + row * K * 2;
#pragma unroll
                        for (int k = 0; k < 2 * RVE_MAX_RHS; ++k)
                            if (k < 2 * K) acc[k] += wgt * rv[k];
                    }
                }
            double* o = rc + (size_t)idx * K * 2;
            for (int k = 0; k < 2 * K; ++k) o[k] = acc[k];
        }
    }
    __syncthreads();
    // --- down sweep
    for (int l = 0; l < LV.L; ++l) {
        const Grid g = LV.g[l];
        const double* A = LV.A[l] + (size_t)sys * 100 * g.G;
        const double* Di = LV.dinv[l] + (size_t)sys * g.G * 4;
        double* rc = LV.rc[l] + (size_t)sys * g.G * K * 2;
        double* xc = LV.xc[l] + (size_t)sys * g.G * K * 2;
        double* tc = LV.tc[l] + (size_t)sys * g.G * K * 2;
        // x = om Dinv rc
        for (int idx = tid; idx < g.G; idx += nth) {
            if (!grid_valid(g, idx % g.nx, idx / g.nx)) continue;
            const double* d = Di + 4 * (size_t)idx;
            for (int k = 0; k < K; ++k) {
                const double r0 = rc[((size_t)idx * K + k) * 2], r1 = rc[((size_t)idx * K + k) * 2 + 1];
                xc[((size_t)idx * K + k) * 2] = om * (d[0] * r0 + d[1] * r1);
                xc[((size_t)idx * K + k) * 2 + 1] = om * (d[2] * r0 + d[3] * r1);
            }
        }
        __syncthreads();
        const bool coarsest = (l == LV.L - 1);
        const int nsw = coarsest ? 3 : 0;  // extra Jacobi sweeps on the coarsest level
        for (int sw = 0; sw < (coarsest ? nsw : 1); ++sw) {
            // t = rc - A x
            for (int idx = tid; idx < g.G; idx += nth) {
                const int i = idx % g.nx, j = idx / g.nx;
                if (!grid_valid(g, i, j)) continue;
                double acc[2 * RVE_MAX_RHS];
#pragma unroll
                for (int k = 0; k < 2 * RVE_MAX_RHS; ++k) acc[k] = 0.0;
                stencil_apply(g, A, xc, K, i, j, acc);
                for (int k = 0; k < 2 * K; ++k) tc[(size_t)idx * K * 2 + k] = rc[(size_t)idx * K * 2 + k] - acc[k];
            }
            __syncthreads();
            if (coarsest) {
                for (int idx = tid; idx < g.G; idx += nth) {
                    if (!grid_valid(g, idx % g.nx, idx / g.nx)) continue;
                    const double* d = Di + 4 * (size_t)idx;
                    for (int k = 0; k < K; ++k) {
                        const double t0 = tc[((size_t)idx * K + k) * 2], t1 = tc[((size_t)idx * K + k) * 2 + 1];
                        xc[((size_t)idx * K + k) * 2] += om * (d[0] * t0 + d[1] * t1);
                        xc[((size_t)idx * K + k) * 2 + 1] += om * (d[2] * t0 + d[3] * t1);
                    }
                }
                __syncthreads();
            }
        }
        if (!coarsest) {
            // restrict t to level l+1 (full weighting = P^T)
            const Grid gc = LV.g[l + 1];
            double* rcc = LV.rc[l + 1] + (size_t)sys * gc.G * K * 2;
            for (int idx = tid; idx < gc.G; idx += nth) {
                const int I = idx % gc.nx, J = idx / gc.nx;
                if (!grid_valid(gc, I, J)) continue;
                double acc[2 * RVE_MAX_RHS];
#pragma unroll
                for (int k = 0; k < 2 * RVE_MAX_RHS; ++k) acc[k] = 0.0;
                for (int aj = -1; aj <= 1; ++aj)
                    for (int ai = -1; ai <= 1; ++ai) {
                        int nf;
                        if (!grid_node(g, 2 * I + ai, 2 * J + aj, nf)) continue;
                        const double wgt = (ai ? 0.5 : 1.0) * (aj ? 0.5 : 1.0);
                        const double* tv = tc + (size_t)nf * K * 2;
#pragma unroll
                        for (int k = 0; k < 2 * RVE_MAX_RHS; ++k)
                            if (k < 2 * K) acc[k] += wgt * tv[k];
                    }
                for (int k = 0; k < 2 * K; ++k) rcc[(size_t)idx * K * 2 + k] = acc[k];
            }
            __syncthreads();
        }
    }
    // --- up sweep
    for (int l = LV.L - 2; l >= 0; --l) {
        const Grid g = LV.g[l];
        const Grid gc = LV.g[l + 1];
        const double* A = LV.A[l] + (size_t)sys * 100 * g.G;
        const double* Di = LV.dinv[l] + (size_t)sys * g.G * 4;
        double* rc = LV.rc[l] + (size_t)sys * g.G * K * 2;
        double* xc = LV.xc[l] + (size_t)sys * g.G * K * 2;
        double* tc = LV.tc[l] + (size_t)sys * g.G * K * 2;
        const double* xcc = LV.xc[l + 1] + (size_t)sys * gc.G * K * 2;
        // x += P xc
        for (int idx = tid; idx < g.G; idx += nth) {
            const int i = idx % g.nx, j = idx / g.nx;
            if (!grid_valid(g, i, j)) continue;
            const int I0 = i >> 1, J0 = j >> 1;
            const int ni = (i & 1) ? 2 : 1, nj = (j & 1) ? 2 : 1;
            for (int bj = 0; bj < nj; ++bj)
                for (int bi = 0; bi < ni; ++bi) {
                    int nc;
                    if (!grid_node(gc, I0 + bi, J0 + bj, nc)) continue;
                    const double wgt = ((i & 1) ? 0.5 : 1.0) * ((j & 1) ? 0.5 : 1.0);
                    for (int k = 0; k < 2 * K; ++k) xc[(size_t)idx * K * 2 + k] += wgt * xcc[(size_t)nc * K * 2 + k];
                }
        }
        __syncthreads();
        // post-smoothing: t = rc - A x ; x += om Dinv t
        for (int idx = tid; idx < g.G; idx += nth) {
            const int i = idx % g.nx, j = idx / g.nx;
            if (!grid_valid(g, i, j)) continue;
            double acc[2 * RVE_MAX_RHS];
#pragma unroll
            for (int k = 0; k < 2 * RVE_MAX_RHS; ++k) acc[k] = 0.0;
            stencil_apply(g, A, xc, K, i, j, acc);
            for (int k = 0; k < 2 * K; ++k) tc[(size_t)idx * K * 2 + k] = rc[(size_t)idx * K * 2 + k] - acc[k];
        }
        __syncthreads();
        for (int idx = tid; idx < g.G; idx += nth) {
            if (!grid_valid(g, idx % g.nx, idx / g.nx)) continue;
            const double* d = Di + 4 * (size_t)idx;
            for (int k = 0; k < K; ++k) {
                const double t0 = tc[((size_t)idx * K + k) * 2], t1 = tc[((size_t)idx * K + k) * 2 + 1];
                xc[((size_t)idx * K + k) * 2] += om * (d[0] * t0 + d[1] * t1);
                xc[((size_t)idx * K + k) * 2 + 1] += om * (d[2] * t0 + d[3] * t1);
            }
        }
        __syncthreads();
    }
    // --- coarse part of r.z = rc1 . xc1, then beta
    {
        const Grid g = LV.g[0];
        const double* rc = LV.rc[0] + (size_t)sys * g.G * K * 2;
        const double* xc = LV.xc[0] + (size_t)sys * g.G * K * 2;
        double acc[RVE_MAX_RHS] = {0, 0, 0, 0};
        for (int idx = tid; idx < g.G; idx += nth) {
            if (!grid_valid(g, idx % g.nx, idx / g.nx)) continue;
#pragma unroll
            for (int k = 0; k < RVE_MAX_RHS; ++k)
                if (k < K)
                    acc[k] += rc[((size_t)idx * K + k) * 2] * xc[((size_t)idx * K + k) * 2] +
                              rc[((size_t)idx * K + k) * 2 + 1] * xc[((size_t)idx * K + k) * 2 + 1];
        }
        __shared__ double s_red[256][RVE_MAX_RHS];
#pragma unroll
        for (int k = 0; k < RVE_MAX_RHS; ++k) s_red[tid][k] = acc[k];
        __syncthreads();
        for (int st = nth >> 1; st > 0; st >>= 1) {
            if (tid < st)
#pragma unroll
                for (int k = 0; k < RVE_MAX_RHS; ++k) s_red[tid][k] += s_red[tid + st][k];
            __syncthreads();
        }
        if (tid < K) {
            double* s = sc + ((size_t)sys * 4 + tid) * SC_N;
            const double rz = s[SC_RDR] + s_red[0][tid];
            const double rzold = s[SC_RZ];
            s[SC_RZOLD] = rzold;
            s[SC_RZ] = rz;
            s[SC_BETA] = (!first && rzold > 0.0) ? rz / rzold : 0.0;
        }
    }
}

// p = Dinv r + Z xc1 + beta p     (thread = (row in chunk, rhs))
__global__ void __launch_bounds__(256)
k_direction(double2* __restrict__ p, const double2* __restrict__ r, const double* __restrict__ dinv,
            const double* __restrict__ row_xy, const double* __restrict__ box, int K, ChunkTab ct,
            const int* __restrict__ active, const double* __restrict__ sc, int precond, Grid g1,
            const double* xc1) {
    const int chunk = blockIdx.x;
    const int sys = ct.chunk_sys[chunk];
    if (!active[sys]) return;
    const int row0 = ct.chunk_row0[chunk], nrow = ct.chunk_n[chunk];
    const int lr = threadIdx.x >> 2, rq = threadIdx.x & 3;
    if (lr >= nrow || rq >= K) return;
    const int row = row0 + lr;
    const double beta = sc[((size_t)sys * 4 + rq) * SC_N + SC_BETA];
    const double2 rv = r[(size_t)row * K + rq];
    const double* d = dinv + 4 * (size_t)row;
    double z0 = d[0] * rv.x + d[1] * rv.y, z1 = d[2] * rv.x + d[3] * rv.y;
    if (precond) {
        int ci, cj; double s, t;
        coarse_cell(row_xy[2 * (size_t)row], row_xy[2 * (size_t)row + 1], box[4 * sys], box[4 * sys + 1],
                    g1.ncx / box[4 * sys + 2], g1.ncy / box[4 * sys + 3], g1.ncx, g1.ncy, ci, cj, s, t);
        const double* xs = xc1 + (size_t)sys * g1.G * K * 2;
#pragma unroll
        for (int dj = 0; dj < 2; ++dj)
#pragma unroll
            for (int di = 0; di < 2; ++di) {
                int nc;
                if (!grid_node(g1, ci + di, cj + dj, nc)) continue;
                const double wgt = (di ? s : 1.0 - s) * (dj ? t : 1.0 - t);
                z0 += wgt * xs[((size_t)nc * K + rq) * 2];
                z1 += wgt * xs[((size_t)nc * K + rq) * 2 + 1];
            }
    }
    double2 pv = p[(size_t)row * 4 + rq];
    pv.x = z0 + beta * pv.x; pv.y = z1 + beta * pv.y;
    p[(size_t)row * 4 + rq] = pv;
}

// ---------------------------------------------------------------------------------------------
// (4) right-hand sides
// ---------------------------------------------------------------------------------------------

// dnn pre-step (micro_model_dnn.py:85-94): B = -(1/vol) oint uD (x) n ds (exact trapezoid on the
// P1 boundary of Vref), eps_eff = eps - voigt(B); thread per (sys, rhs).
__global__ void k_dnn_prep(int n_sys, int K, int nside, const double* __restrict__ vbox /*[4]*/,
                           const double* __restrict__ uD /*[sys][K][2*(4*nside+1)] or null*/,
                           const double* __restrict__ eps, double* __restrict__ eps_eff,
                           double* __restrict__ Bmat /*[sys][K][4]*/) {
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

// b = -int B^T D eps  - K_fb g : warp per element.  First term in closed form per node
// (sigma0 . int grad phi_a), second only where an element touches Dirichlet nodes with data.
__global__ void __launch_bounds__(256)
k_rhs(int n_elem, int K, const int* __restrict__ elem_node, const unsigned char* __restrict__ elem_reg,
      const int* __restrict__ elem_sys, const double* __restrict__ node_xy, const int* __restrict__ node_row,
      const int* __restrict__ node_bnd, const double* __restrict__ eps_eff, const double* __restrict__ gval,
      double* __restrict__ b, double* __restrict__ sc, Mat8 mat) {
    __shared__ double s_mat[8];
    if (threadIdx.x < 8) s_mat[threadIdx.x] = mat.v[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (e >= n_elem) return;
    const int sys = elem_sys[e];
    int node = 0, row = -1, bnd = -1;
    double px = 0, py = 0;
    if (lane < 6) {
        node = elem_node[(size_t)e * 6 + lane];
        row = node_row[node];
        bnd = (gval && row < 0) ? node_bnd[node] : -1;
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
    {
        double gx = 0, gy = 0;
        if (lane < 6) p2_grad_int(g, lane, gx, gy);
        for (int k = 0; k < K; ++k) {
            const double* ev = eps_eff + 3 * ((size_t)sys * K + k);
            const double s00 = (lam + 2 * mu) * ev[0] + lam * ev[1];
            const double s11 = lam * ev[0] + (lam + 2 * mu) * ev[1];
            const double s01 = mu * ev[2];
            const double f0 = -(s00 * gx + s01 * gy), f1 = -(s01 * gx + s11 * gy);
            if (lane < 6 && row >= 0) {
                atomicAdd(b + ((size_t)row * K + k) * 2, f0);
                atomicAdd(b + ((size_t)row * K + k) * 2 + 1, f1);
            }
            // reference magnitude of the load before cancellation (absolute convergence floor)
            double m2 = (lane < 6 && row >= 0) ? f0 * f0 + f1 * f1 : 0.0;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) m2 += __shfl_xor_sync(0xffffffffu, m2, o);
            if (lane == 0) atomicAdd(sc + ((size_t)sys * 4 + k) * SC_N + SC_BREF, m2);
        }
    }
    const unsigned anyb = __ballot_sync(0xffffffffu, bnd >= 0);
    if (anyb == 0) return;
    int a = 0, bb = 0;
    pair_ab(lane < 21 ? lane : 0, a, bb);
    const int ra = __shfl_sync(0xffffffffu, row, a), rb = __shfl_sync(0xffffffffu, row, bb);
    const int ba = __shfl_sync(0xffffffffu, bnd, a), bnb = __shfl_sync(0xffffffffu, bnd, bb);
    if (lane < 21 && a != bb) {
        double Kb[4];
        p2_stiff_block(g, a, bb, lam, mu, Kb);
        if (ra >= 0 && bnb >= 0) {  // row a free, column b prescribed
            for (int k = 0; k < K; ++k) {
                const double g0 = gval[((size_t)bnb * K + k) * 2], g1 = gval[((size_t)bnb * K + k) * 2 + 1];
                atomicAdd(b + ((size_t)ra * K + k) * 2, -(Kb[0] * g0 + Kb[1] * g1));
                atomicAdd(b + ((size_t)ra * K + k) * 2 + 1, -(Kb[2] * g0 + Kb[3] * g1));
            }
        }
        if (rb >= 0 && ba >= 0) {  // row b free, column a prescribed: K_ba = K_ab^T
            for (int k = 0; k < K; ++k) {
                const double g0 = gval[((size_t)ba * K + k) * 2], g1 = gval[((size_t)ba * K + k) * 2 + 1];
                atomicAdd(b + ((size_t)rb * K + k) * 2, -(Kb[0] * g0 + Kb[2] * g1));
                atomicAdd(b + ((size_t)rb * K + k) * 2 + 1, -(Kb[1] * g0 + Kb[3] * g1));
            }
        }
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
//   [0..1] int u~ (L), [2..5] int grad u~ (L), [6..8] int sigma(u~) (L: 00,11,01), [9..11] same over all regions
#define ACC_PER_SYS (6 + 16 * RVE_MAX_RHS)

__global__ void __launch_bounds__(128)
k_epi_elements(int n_elem, int NL, const int* __restrict__ elem_node, const unsigned char* __restrict__ elem_reg,
               const int* __restrict__ elem_sys, const double* __restrict__ node_xy,
               const double* __restrict__ U, double* __restrict__ acc, Mat8 mat) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_elem) return;
    const int sys = elem_sys[e];
    int nd[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) nd[k] = elem_node[(size_t)e * 6 + k];
    TriGeom g;
    tri_geom(node_xy[2 * (size_t)nd[0]], node_xy[2 * (size_t)nd[0] + 1], node_xy[2 * (size_t)nd[1]],
             node_xy[2 * (size_t)nd[1] + 1], node_xy[2 * (size_t)nd[2]], node_xy[2 * (size_t)nd[2] + 1], g);
    const int reg = elem_reg[e];
    const double lam = mat.v[2 * reg], mu = mat.v[2 * reg + 1];
    double* A = acc + (size_t)sys * ACC_PER_SYS;
    atomicAdd(A + reg, g.area);
    const bool inL = reg < 2;
    if (inL) {
        const double cx = (node_xy[2 * (size_t)nd[0]] + node_xy[2 * (size_t)nd[1]] + node_xy[2 * (size_t)nd[2]]) / 3.0;
        const double cy = (node_xy[2 * (size_t)nd[0] + 1] + node_xy[2 * (size_t)nd[1] + 1] + node_xy[2 * (size_t)nd[2] + 1]) / 3.0;
        atomicAdd(A + 4, cx * g.area); atomicAdd(A + 5, cy * g.area);
    }
    double gx[6], gy[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) p2_grad_int(g, a, gx[a], gy[a]);
    for (int l = 0; l < NL; ++l) {
        double G00 = 0, G01 = 0, G10 = 0, G11 = 0, ui0 = 0, ui1 = 0;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            const double u0 = U[((size_t)nd[a] * NL + l) * 2], u1 = U[((size_t)nd[a] * NL + l) * 2 + 1];
            G00 += u0 * gx[a]; G01 += u0 * gy[a]; G10 += u1 * gx[a]; G11 += u1 * gy[a];
            if (a >= 3) { ui0 += u0; ui1 += u1; }
        }
        ui0 *= g.area / 3.0; ui1 *= g.area / 3.0;
        const double tr = G00 + G11;
        const double s00 = lam * tr + 2 * mu * G00, s11 = lam * tr + 2 * mu * G11, s01 = mu * (G01 + G10);
        double* Al = A + 6 + 16 * l;
        if (inL) {
            atomicAdd(Al + 0, ui0); atomicAdd(Al + 1, ui1);
            atomicAdd(Al + 2, G00); atomicAdd(Al + 3, G01); atomicAdd(Al + 4, G10); atomicAdd(Al + 5, G11);
            atomicAdd(Al + 6, s00); atomicAdd(Al + 7, s11); atomicAdd(Al + 8, s01);
        }
        atomicAdd(Al + 9, s00); atomicAdd(Al + 10, s11); atomicAdd(Al + 11, s01);
    }
}

// thread per (sys, load): sigma over {0,1}, sigma over all, a, B, tangent column
__global__ void k_epi_finalize(int n_sys, int NL, const double* __restrict__ acc, const double* __restrict__ eps_eff,
                               Mat8 mat, double* __restrict__ sigL, double* __restrict__ sigT,
                               double* __restrict__ aout, double* __restrict__ Bout) {
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
