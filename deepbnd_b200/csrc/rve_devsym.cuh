// rve_devsym.cuh — the symbolic phase of rve_set_batch on the GPU.
//
// Input: the caller's raw meshes (vertices, triangles, region tags, prefix offsets), uploaded as they are.
// Output: everything the numeric kernels need — P2 numbering (mid-edge nodes), boundary edges / nodes,
// the constraint map of the boundary model, active rows ordered along the Morton curve of the level-1
// cells, the elements around every row, row capacities, length-sorted windows, SELL slice widths and the
// Vref sample points.  One CTA per RVE walks its mesh with block-wide scans and small per-thread sorts;
// every ordering is made deterministic (sorted buckets), and is the one rve_symbolic.hpp (the CPU test
// hook rve_host_symbolic) produces, so the two can be compared entry by entry (tests/test_gpu_parity.py).
//
// Replaces what dolfin does once per mesh in the reference: FunctionSpace(mesh, "CG", 2) DOF numbering,
// DirichletBC / periodic-map set-up and sparsity-pattern build (micro_model.py:24-60,
// micro_model_dnn.py:50-83), and the bbox-tree point location of usol.interpolate
// (simulation_snapshots.py:56-58).
#pragma once

#define DS_THREADS 1024
#define DS_EMASK 0x3fffffff
#define DS_EBND 0x40000000

// per-system counters written by the symbolic kernels (8 ints per system)
enum { DC_NE = 0, DC_NBE, DC_NB, DC_ERR, DC_NA, DC_NBLK, DC_NGRP, DC_NADJ };
// error codes (DC_ERR)
enum { DE_REGION = 1, DE_VERTEX, DE_DEGENERATE, DE_NONMANIFOLD, DE_PER_LR, DE_PER_BT, DE_PER_CHAIN, DE_DNN_BOX,
       DE_VREF_OUTSIDE, DE_ROW_LONG };

static const char* ds_error_text(int code) {
    switch (code) {
        case DE_REGION: return "region tag out of range (need max-min <= 3)";
        case DE_VERTEX: return "triangle vertex id out of range";
        case DE_DEGENERATE: return "degenerate triangle";
        case DE_NONMANIFOLD: return "non-manifold edge";
        case DE_PER_LR: return "periodic: left/right boundary nodes do not match";
        case DE_PER_BT: return "periodic: bottom/top boundary nodes do not match";
        case DE_PER_CHAIN: return "periodic: constraint map is not idempotent";
        case DE_DNN_BOX: return "dnn: mesh boundary does not coincide with the Vref box";
        case DE_VREF_OUTSIDE: return "Vref sample point outside the mesh";
        case DE_ROW_LONG: return "a row couples to more than 63 nodes";
        default: return "symbolic phase failed";
    }
}

// ---- block-wide helpers (blockDim.x a multiple of 32, <= 1024) --------------------------------------------
// exclusive scan of a[0..n) in place; every thread gets the total.  s_w: shared int[33]
__device__ __forceinline__ int ds_scan(int* a, int n, int* s_w) {
    const int tid = threadIdx.x, lane = tid & 31, wv = tid >> 5, nw = blockDim.x >> 5;
    int carry = 0;
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + tid;
        const int v = i < n ? a[i] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_w[wv] = inc;
        __syncthreads();
        if (wv == 0) {
            const int w = lane < nw ? s_w[lane] : 0;
            int winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
            s_w[lane] = winc - w;
            if (lane == 31) s_w[32] = winc;
        }
        __syncthreads();
        if (i < n) a[i] = carry + s_w[wv] + inc - v;
        carry += s_w[32];
        __syncthreads();
    }
    return carry;
}

__device__ __forceinline__ int ds_sum(int v, int* s_w) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (lane == 0) s_w[wv] = v;
    __syncthreads();
    if (wv == 0) {
        int w = lane < nw ? s_w[lane] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
        if (lane == 0) s_w[32] = w;
    }
    __syncthreads();
    const int r = s_w[32];
    __syncthreads();
    return r;
}

__device__ __forceinline__ int ds_min(int v, int* s_w) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (lane == 0) s_w[wv] = v;
    __syncthreads();
    if (wv == 0) {
        int w = lane < nw ? s_w[lane] : s_w[0];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w = min(w, __shfl_xor_sync(0xffffffffu, w, o));
        if (lane == 0) s_w[32] = w;
    }
    __syncthreads();
    const int r = s_w[32];
    __syncthreads();
    return r;
}

// op: 0 = min, 1 = max, 2 = sum.  s_d: shared double[33]
__device__ __forceinline__ double ds_reduce_d(double v, int op, double* s_d) {
    auto comb = [op](double a, double b) { return op == 0 ? fmin(a, b) : (op == 1 ? fmax(a, b) : a + b); };
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = comb(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (lane == 0) s_d[wv] = v;
    __syncthreads();
    if (wv == 0) {
        double w = lane < nw ? s_d[lane] : s_d[0];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w = comb(w, __shfl_xor_sync(0xffffffffu, w, o));
        if (lane == 0) s_d[32] = w;
    }
    __syncthreads();
    const double r = s_d[32];
    __syncthreads();
    return r;
}

// ascending insertion sort of a short int list (optionally with a companion array)
__device__ __forceinline__ void ds_isort(int* a, int m) {
    for (int i = 1; i < m; ++i) {
        const int x = a[i];
        int j = i - 1;
        while (j >= 0 && a[j] > x) { a[j + 1] = a[j]; --j; }
        a[j + 1] = x;
    }
}

struct DsMesh {
    const double* xy;     // [NV][2]
    int* tri;             // [NT][3]  re-oriented CCW in place
    const int* region;    // [NT]
    const int* voff;      // [n_sys+1]
    const int* toff;      // [n_sys+1]
};

// ---------------------------------------------------------------------------------------------
// (A) bounding box, orientation, edges.  The 3 nt half-edges of a system are bucketed by their smaller
//     vertex (count, scan, fill), every bucket is sorted by the larger vertex, and edge e = rank of
//     (va, vb) in that order — the numbering of a global sort by (va, vb).  emid[3t+k] = local id of the edge
//     from local vertex k to k+1 of element t, bit 30 set on mesh-boundary edges.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DS_THREADS)
k_ds_edges(DsMesh M, int* __restrict__ vptr /*[NV+n_sys]*/, int* __restrict__ vfill /*[NV]*/, int* __restrict__ he_vb /*[3NT]*/,
           int* __restrict__ he_slot /*[3NT]*/, int* __restrict__ emid /*[3NT]*/, unsigned char* __restrict__ elem_reg,
           double* __restrict__ box, int* __restrict__ sys_cnt, double* __restrict__ sys_ext) {
    __shared__ int s_w[33];
    __shared__ double s_d[33];
    __shared__ int s_err;
    const int s = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
    const int v0 = M.voff[s], nv = M.voff[s + 1] - v0, t0 = M.toff[s], nt = M.toff[s + 1] - t0;
    const double* xy = M.xy + 2 * (size_t)v0;
    int* tri = M.tri + 3 * (size_t)t0;
    const int* region = M.region + t0;
    int* vp = vptr + v0 + s;
    int* vf = vfill + v0;
    int* hv = he_vb + 3 * (size_t)t0;
    int* hs = he_slot + 3 * (size_t)t0;
    int* em = emid + 3 * (size_t)t0;
    if (tid == 0) s_err = 0;
    double xmin = 1e300, ymin = 1e300, xmax = -1e300, ymax = -1e300;
    for (int v = tid; v < nv; v += nth) {
        const double x = xy[2 * v], y = xy[2 * v + 1];
        xmin = fmin(xmin, x); xmax = fmax(xmax, x); ymin = fmin(ymin, y); ymax = fmax(ymax, y);
        vp[v] = 0; vf[v] = 0;
    }
    if (tid == 0) vp[nv] = 0;
    int rmin = 1 << 30;
    for (int t = tid; t < nt; t += nth) rmin = min(rmin, region[t]);
    xmin = ds_reduce_d(xmin, 0, s_d); ymin = ds_reduce_d(ymin, 0, s_d);
    xmax = ds_reduce_d(xmax, 1, s_d); ymax = ds_reduce_d(ymax, 1, s_d);
    rmin = ds_min(rmin, s_w);
    if (tid == 0) { box[4 * s] = xmin; box[4 * s + 1] = ymin; box[4 * s + 2] = xmax - xmin; box[4 * s + 3] = ymax - ymin; }
    double ext_sum = 0.0;
    for (int t = tid; t < nt; t += nth) {
        const int r = region[t] - rmin;
        if (r < 0 || r > 3) { atomicMax(&s_err, DE_REGION); continue; }
        elem_reg[t0 + t] = (unsigned char)r;
        int a = tri[3 * t], b = tri[3 * t + 1], c = tri[3 * t + 2];
        if (a < 0 || a >= nv || b < 0 || b >= nv || c < 0 || c >= nv) { atomicMax(&s_err, DE_VERTEX); continue; }
        const double ax = xy[2 * a], ay = xy[2 * a + 1], bx = xy[2 * b], by = xy[2 * b + 1], cx = xy[2 * c], cy = xy[2 * c + 1];
        const double a2 = (bx - ax) * (cy - ay) - (by - ay) * (cx - ax);
        if (a2 == 0.0) { atomicMax(&s_err, DE_DEGENERATE); continue; }
        if (a2 < 0) { tri[3 * t + 1] = c; tri[3 * t + 2] = b; const int q = b; b = c; c = q; }
        const double ex = fmax(ax, fmax(bx, cx)) - fmin(ax, fmin(bx, cx)), ey = fmax(ay, fmax(by, cy)) - fmin(ay, fmin(by, cy));
        ext_sum += fmax(ex, ey);
        atomicAdd(&vp[min(a, b)], 1); atomicAdd(&vp[min(b, c)], 1); atomicAdd(&vp[min(c, a)], 1);
    }
    ext_sum = ds_reduce_d(ext_sum, 2, s_d);
    if (s_err) {   // uniform after the barrier inside the reduction
        if (tid == 0) { sys_cnt[8 * s + DC_ERR] = s_err; sys_ext[s] = 0.0; }
        return;
    }
    const int nhe = ds_scan(vp, nv, s_w);
    if (tid == 0) vp[nv] = nhe;
    __syncthreads();
    for (int t = tid; t < nt; t += nth)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int a = tri[3 * t + k], b = tri[3 * t + (k + 1) % 3];
            const int lo = min(a, b), hi = max(a, b);
            const int pos = vp[lo] + atomicAdd(&vf[lo], 1);
            hv[pos] = hi; hs[pos] = 3 * t + k;
        }
    __syncthreads();
    // sort every bucket by (larger vertex, slot); count its distinct edges
    for (int v = tid; v < nv; v += nth) {
        const int b0 = vp[v], m = vp[v + 1] - b0;
        for (int i = 1; i < m; ++i) {
            const int xv = hv[b0 + i], xs = hs[b0 + i];
            int j = i - 1;
            while (j >= 0 && (hv[b0 + j] > xv || (hv[b0 + j] == xv && hs[b0 + j] > xs))) {
                hv[b0 + j + 1] = hv[b0 + j]; hs[b0 + j + 1] = hs[b0 + j]; --j;
            }
            hv[b0 + j + 1] = xv; hs[b0 + j + 1] = xs;
        }
        int nu = 0;
        for (int i = 0; i < m; ++i) if (i == 0 || hv[b0 + i] != hv[b0 + i - 1]) ++nu;
        vf[v] = nu;
    }
    __syncthreads();
    const int ne = ds_scan(vf, nv, s_w);
    int nbe = 0;
    for (int v = tid; v < nv; v += nth) {
        const int b0 = vp[v], m = vp[v + 1] - b0;
        int e = vf[v] - 1, i = 0;
        while (i < m) {
            int j = i + 1;
            while (j < m && hv[b0 + j] == hv[b0 + i]) ++j;
            ++e;
            const int run = j - i;
            if (run > 2) atomicMax(&s_err, DE_NONMANIFOLD);
            const int flag = run == 1 ? DS_EBND : 0;
            nbe += run == 1;
            for (int q = i; q < j; ++q) em[hs[b0 + q]] = e | flag;
            i = j;
        }
    }
    nbe = ds_sum(nbe, s_w);
    // boundary vertices (the boundary nodes are those plus one mid-edge node per boundary edge)
    for (int v = tid; v < nv; v += nth) vp[v] = 0;
    __syncthreads();
    for (int hh = tid; hh < 3 * nt; hh += nth)
        if (em[hh] & DS_EBND) {
            const int t = hh / 3, k = hh - 3 * t;
            vp[tri[3 * t + k]] = 1; vp[tri[3 * t + (k + 1) % 3]] = 1;
        }
    __syncthreads();
    int nbv = 0;
    for (int v = tid; v < nv; v += nth) nbv += vp[v];
    nbv = ds_sum(nbv, s_w);
    if (tid == 0) {
        sys_cnt[8 * s + DC_NE] = ne; sys_cnt[8 * s + DC_NBE] = nbe; sys_cnt[8 * s + DC_NB] = nbe + nbv;
        sys_cnt[8 * s + DC_ERR] = s_err;
        sys_ext[s] = nt > 0 ? ext_sum / nt : 0.0;
    }
}

// Vref perimeter parameter of a point on the box boundary: node k of Vref sits at s = k (rve_symbolic.hpp)
__device__ __forceinline__ double ds_perimeter_param(double x, double y, const double* vb, int nside) {
    const double x0 = vb[0], y0 = vb[1], Lx = vb[2], Ly = vb[3];
    const double tol = 1e-9 * fmax(Lx, Ly);
    if (fabs(y - y0) < tol && x < x0 + Lx - tol) return (x - x0) / Lx * nside;
    if (fabs(x - (x0 + Lx)) < tol && y < y0 + Ly - tol) return nside + (y - y0) / Ly * nside;
    if (fabs(y - (y0 + Ly)) < tol && x > x0 + tol) return 2 * nside + (x0 + Lx - x) / Lx * nside;
    if (fabs(x - x0) < tol) return 3 * nside + (y0 + Ly - y) / Ly * nside;
    return -1.0;
}

struct DsRows {
    // inputs
    const int* emid; const int* node_base; const int* bnd_base; const int* bedge_base; const int* cellrank;
    const double* box; const double* vbox;     // vbox: [4] shared Vref box or nullptr (then the mesh box)
    int model, ncx, ncy, nside;
    // final arrays in the global node / element / boundary index spaces
    double* node_xy; int* node_va; int* node_vb; int* node_sys; int* node_bnd; int* elem_node; int* elem_sys;
    int* bnd_sys; int* bnd_node; double* bnd_s; int* bedge; double* bedge_nl;
    // scratch: [N] node-indexed unless noted
    int* master; int* ncell; int* pnode; int* prow; int* nrow0; int* cur; int* len0; int* newid; int* oldid;
    int* adjc;          // [N + n_sys]
    int* adj;           // [6 NT]
    int* scan3;         // [3 NT]
    int* ccnt;          // [n_sys][ncells+1]
    int* cfill;         // [n_sys][ncells]
    unsigned char* rowb;
    // per-system results in local numbering for k_ds_finish
    int* row_node_loc; int* node_row_loc;
    int* loc_rowptr;    // [N + n_sys]
    int* loc_adjptr;    // [N + n_sys]
    int* loc_slw;       // [N/32 + 2 n_sys]  offset node_base/32 + 2 s
    int* sys_cnt;
};

// ---------------------------------------------------------------------------------------------
// (B) nodes, boundary lists, constraint map, Morton row order, elements around rows, capacities, windows
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DS_THREADS)
k_ds_rows(DsMesh M, DsRows A) {
    __shared__ int s_w[33];
    __shared__ int s_err;
    __shared__ int s_hist[4][8][64], s_tot[4][64], s_pre[4][64];   // window sort
    const int s = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
    if (A.sys_cnt[8 * s + DC_ERR]) return;
    const int v0 = M.voff[s], nv = M.voff[s + 1] - v0, t0 = M.toff[s], nt = M.toff[s + 1] - t0;
    const double* xy = M.xy + 2 * (size_t)v0;
    const int* tri = M.tri + 3 * (size_t)t0;
    const int* em = A.emid + 3 * (size_t)t0;
    const int nb0 = A.node_base[s], nn = A.node_base[s + 1] - nb0;
    const int bb0 = A.bnd_base[s], be0 = A.bedge_base[s];
    double* nxy = A.node_xy + 2 * (size_t)nb0;
    int* nva = A.node_va + nb0; int* nvb = A.node_vb + nb0; int* nsys = A.node_sys + nb0; int* nbnd = A.node_bnd + nb0;
    int* en = A.elem_node + 6 * (size_t)t0;
    int* master = A.master + nb0; int* ncell = A.ncell + nb0; int* pnode = A.pnode + nb0; int* prow = A.prow + nb0;
    int* nrow0 = A.nrow0 + nb0; int* cur = A.cur + nb0; int* len0 = A.len0 + nb0; int* newid = A.newid + nb0;
    int* oldid = A.oldid + nb0; int* adjc = A.adjc + nb0 + s; int* adj = A.adj + 6 * (size_t)t0;
    int* scan3 = A.scan3 + 3 * (size_t)t0;
    unsigned char* rowb = A.rowb + nb0;
    if (tid == 0) s_err = 0;
    // ---- nodes and element connectivity
    for (int v = tid; v < nv; v += nth) {
        nxy[2 * v] = xy[2 * v]; nxy[2 * v + 1] = xy[2 * v + 1];
        nva[v] = v; nvb[v] = v; nsys[v] = s; nbnd[v] = -1;
    }
    for (int hh = tid; hh < 3 * nt; hh += nth) {
        const int t = hh / 3, k = hh - 3 * t;
        const int a = tri[3 * t + k], b = tri[3 * t + (k + 1) % 3];
        const int e = em[hh] & DS_EMASK, n = nv + e;
        nxy[2 * n] = 0.5 * (xy[2 * a] + xy[2 * b]); nxy[2 * n + 1] = 0.5 * (xy[2 * a + 1] + xy[2 * b + 1]);
        nva[n] = min(a, b); nvb[n] = max(a, b); nsys[n] = s; nbnd[n] = -1;
        en[6 * t + k] = nb0 + a; en[6 * t + 3 + k] = nb0 + n;
        scan3[hh] = (em[hh] & DS_EBND) ? 1 : 0;
    }
    for (int t = tid; t < nt; t += nth) A.elem_sys[t0 + t] = s;
    __syncthreads();
    // ---- boundary edges in (element, local edge) order with outward normal * length; boundary nodes
    ds_scan(scan3, 3 * nt, s_w);
    for (int hh = tid; hh < 3 * nt; hh += nth)
        if (em[hh] & DS_EBND) {
            const int t = hh / 3, k = hh - 3 * t;
            const int a = tri[3 * t + k], b = tri[3 * t + (k + 1) % 3], m = nv + (em[hh] & DS_EMASK);
            const size_t ge = (size_t)be0 + scan3[hh];
            A.bedge[3 * ge] = nb0 + a; A.bedge[3 * ge + 1] = nb0 + b; A.bedge[3 * ge + 2] = nb0 + m;
            A.bedge_nl[2 * ge] = xy[2 * b + 1] - xy[2 * a + 1]; A.bedge_nl[2 * ge + 1] = -(xy[2 * b] - xy[2 * a]);
            nbnd[a] = 0; nbnd[b] = 0; nbnd[m] = 0;
        }
    __syncthreads();
    for (int n = tid; n < nn; n += nth) { cur[n] = nbnd[n] == 0 ? 1 : 0; master[n] = n; }
    __syncthreads();
    const int nb = ds_scan(cur, nn, s_w);
    for (int n = tid; n < nn; n += nth)
        if (nbnd[n] == 0) {
            const int i = bb0 + cur[n];
            nbnd[n] = i; A.bnd_node[i] = nb0 + n; A.bnd_sys[i] = s;
        }
    __syncthreads();
    const int* bnode = A.bnd_node + bb0;           // global node ids, ascending
    const double x0 = A.box[4 * s], y0 = A.box[4 * s + 1], Lx = A.box[4 * s + 2], Ly = A.box[4 * s + 3];
    const double tol = 1e-9 * fmax(Lx, Ly), tol7 = 1e-7 * fmax(Lx, Ly);
    const double* vb = A.vbox ? A.vbox : A.box + 4 * s;
    if (A.model == RVE_MODEL_DNN) {
        for (int i = tid; i < nb; i += nth) {
            const int n = bnode[i] - nb0;
            const int a = nva[n], b = nvb[n];
            const double sa = ds_perimeter_param(xy[2 * a], xy[2 * a + 1], vb, A.nside);
            const double sb = ds_perimeter_param(xy[2 * b], xy[2 * b + 1], vb, A.nside);
            if (sa < 0 || sb < 0) atomicMax(&s_err, DE_DNN_BOX);
            A.bnd_s[2 * ((size_t)bb0 + i)] = sa; A.bnd_s[2 * ((size_t)bb0 + i) + 1] = sb;
        }
    }
    // ---- constraint map: master[n] = n (own row), another node (periodic slave) or -1 (Dirichlet)
    if (A.model == RVE_MODEL_LIN || A.model == RVE_MODEL_DNN) {
        for (int i = tid; i < nb; i += nth) master[bnode[i] - nb0] = -1;
    } else if (A.model == RVE_MODEL_PER) {
        // slaves: x = x1 -> x0, y = y1 -> y0 (corners -> (x0,y0)); match by the other coordinate: the first node of
        // the left (bottom) side, in (coordinate, id) order, at or above key - tol7, which must lie within tol7
        for (int i = tid; i < nb; i += nth) {
            const int n = bnode[i] - nb0;
            const double x = nxy[2 * n];
            double y = nxy[2 * n + 1];
            const bool onr = fabs(x - (x0 + Lx)) < tol, ont = fabs(y - (y0 + Ly)) < tol;
            int m = n;
            if (onr) {
                double bc = 1e300; int bm = -1;
                for (int j = 0; j < nb; ++j) {
                    const int q = bnode[j] - nb0;
                    if (fabs(nxy[2 * q] - x0) < tol) {
                        const double c = nxy[2 * q + 1];
                        if (c >= y - tol7 && (c < bc || (c == bc && q < bm))) { bc = c; bm = q; }
                    }
                }
                if (bm < 0 || fabs(bc - y) > tol7) { atomicMax(&s_err, DE_PER_LR); continue; }
                m = bm; y = nxy[2 * m + 1];
            }
            if (ont) {
                const double xm = nxy[2 * m];
                double bc = 1e300; int bm = -1;
                for (int j = 0; j < nb; ++j) {
                    const int q = bnode[j] - nb0;
                    if (fabs(nxy[2 * q + 1] - y0) < tol) {
                        const double c = nxy[2 * q];
                        if (c >= xm - tol7 && (c < bc || (c == bc && q < bm))) { bc = c; bm = q; }
                    }
                }
                if (bm < 0 || fabs(bc - xm) > tol7) { atomicMax(&s_err, DE_PER_BT); continue; }
                m = bm;
            }
            master[n] = m;
        }
        __syncthreads();
        for (int i = tid; i < nb; i += nth) {
            const int m = master[bnode[i] - nb0];
            if (m >= 0 && master[m] != m) atomicMax(&s_err, DE_PER_CHAIN);
        }
    }
    __syncthreads();
    if (s_err) { if (tid == 0) A.sys_cnt[8 * s + DC_ERR] = s_err; return; }
    // ---- active nodes ordered by the Morton rank of their level-1 cell, ties by node id
    const int ncx = A.ncx, ncy = A.ncy, ncells = ncx * ncy;
    int* cc = A.ccnt + (size_t)s * (ncells + 1);
    int* cf = A.cfill + (size_t)s * ncells;
    for (int c = tid; c < ncells; c += nth) { cc[c] = 0; cf[c] = 0; }
    if (tid == 0) cc[ncells] = 0;
    __syncthreads();
    const double invHx = ncx / Lx, invHy = ncy / Ly;
    for (int n = tid; n < nn; n += nth) {
        int cr = -1;
        if (master[n] == n) {
            int ci, cj; double ss, tt;
            coarse_cell(nxy[2 * n], nxy[2 * n + 1], x0, y0, invHx, invHy, ncx, ncy, ci, cj, ss, tt);
            cr = A.cellrank[cj * ncx + ci];
            atomicAdd(&cc[cr], 1);
        }
        ncell[n] = cr;
    }
    __syncthreads();
    const int na = ds_scan(cc, ncells, s_w);
    if (tid == 0) cc[ncells] = na;
    __syncthreads();
    for (int n = tid; n < nn; n += nth) {
        const int cr = ncell[n];
        if (cr >= 0) pnode[cc[cr] + atomicAdd(&cf[cr], 1)] = n;
    }
    __syncthreads();
    for (int c = tid; c < ncells; c += nth) ds_isort(pnode + cc[c], cc[c + 1] - cc[c]);
    __syncthreads();
    for (int r = tid; r < na; r += nth) { prow[pnode[r]] = r; adjc[r] = 0; cur[r] = 0; rowb[r] = 0; }
    if (tid == 0) adjc[na] = 0;
    __syncthreads();
    for (int n = tid; n < nn; n += nth) nrow0[n] = master[n] < 0 ? -1 : prow[master[n]];
    __syncthreads();
    // ---- elements around every (preliminary) row, ascending; rows that gather a mesh-boundary node
    for (int q = tid; q < 6 * nt; q += nth) {
        const int r = nrow0[en[q] - nb0];
        if (r >= 0) atomicAdd(&adjc[r], 1);
    }
    for (int i = tid; i < nb; i += nth) {
        const int r = nrow0[bnode[i] - nb0];
        if (r >= 0) rowb[r] = 1;
    }
    __syncthreads();
    const int nadj = ds_scan(adjc, na, s_w);
    if (tid == 0) adjc[na] = nadj;
    __syncthreads();
    for (int q = tid; q < 6 * nt; q += nth) {
        const int r = nrow0[en[q] - nb0];
        if (r >= 0) adj[adjc[r] + atomicAdd(&cur[r], 1)] = q / 6;
    }
    __syncthreads();
    // row CAPACITIES from the element fans: a vertex with d elements couples to 3d+1 nodes in a closed fan, 3d+3 in an
    // open one (rows that gather a mesh-boundary node), a mid-edge node to 9 or 6; inactive neighbours only shorten a
    // row.  The true columns and lengths come from k_sym_cols; SELL padding with these widths is +0.3..1 %
    for (int r = tid; r < na; r += nth) {
        const int d = adjc[r + 1] - adjc[r];
        ds_isort(adj + adjc[r], d);
        const int len = pnode[r] < nv ? 3 * d + (rowb[r] ? 3 : 1) : (d >= 2 ? 9 : 6);
        if (len > 63) atomicMax(&s_err, DE_ROW_LONG);
        len0[r] = len;
    }
    __syncthreads();
    if (s_err) { if (tid == 0) A.sys_cnt[8 * s + DC_ERR] = s_err; return; }
    // ---- windows of RVE_WIN preliminary rows, each stably sorted by descending capacity (4 windows per pass)
    // (stable counting sort on the 64 possible keys: per-warp histograms by __match_any, prefix over keys and warps)
    const int nwin = (na + RVE_WIN - 1) / RVE_WIN;
    const int grp = tid / RVE_WIN, i = tid % RVE_WIN, wvg = i >> 5, lane = tid & 31;
    for (int w0 = 0; w0 < nwin; w0 += nth / RVE_WIN) {
        const int w = w0 + grp;
        const int r0 = w * RVE_WIN;
        const int wn = w < nwin ? min(RVE_WIN, na - r0) : 0;
        const bool valid = i < wn;
        const int key = valid ? 63 - len0[r0 + i] : 64;
        for (int q = tid; q < 4 * 8 * 64; q += nth) (&s_hist[0][0][0])[q] = 0;
        __syncthreads();
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank == 0) s_hist[grp][wvg][key] = __popc(peers);
        __syncthreads();
        if (i < 64) {
            int tot = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) tot += s_hist[grp][q][i];
            s_tot[grp][i] = tot;
        }
        __syncthreads();
        if (i < 32) {            // exclusive prefix over the 64 keys: two per lane
            const int t0k = s_tot[grp][2 * i], t1k = s_tot[grp][2 * i + 1];
            int inc = t0k + t1k;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            s_pre[grp][2 * i] = inc - t0k - t1k; s_pre[grp][2 * i + 1] = inc - t1k;
        }
        __syncthreads();
        if (valid) {
            int pos = s_pre[grp][key] + rank;
            for (int q = 0; q < wvg; ++q) pos += s_hist[grp][q][key];
            newid[r0 + i] = r0 + pos; oldid[r0 + pos] = r0 + i;
        }
        __syncthreads();
    }
    int* rnl = A.row_node_loc + nb0; int* nrl = A.node_row_loc + nb0;
    int* lrp = A.loc_rowptr + nb0 + s; int* lap = A.loc_adjptr + nb0 + s; int* lsw = A.loc_slw + nb0 / 32 + 2 * s;
    for (int r = tid; r < na; r += nth) {
        const int o = oldid[r];
        rnl[r] = pnode[o]; lrp[r] = len0[o]; lap[r] = adjc[o + 1] - adjc[o];
    }
    for (int n = tid; n < nn; n += nth) nrl[n] = nrow0[n] < 0 ? -1 : newid[nrow0[n]];
    __syncthreads();
    const int nsl = (na + 31) / 32;
    for (int sl = tid; sl < nsl; sl += nth) {
        int width = 0;
        for (int r = 32 * sl; r < min(na, 32 * sl + 32); ++r) width = max(width, lrp[r]);
        lsw[sl] = width;
    }
    __syncthreads();
    const int nblk = ds_scan(lrp, na, s_w);
    ds_scan(lap, na, s_w);
    const int ngrp = ds_scan(lsw, nsl, s_w);
    if (tid == 0) {
        lrp[na] = nblk; lap[na] = nadj; lsw[nsl] = ngrp;
        A.sys_cnt[8 * s + DC_NA] = na; A.sys_cnt[8 * s + DC_NBLK] = nblk; A.sys_cnt[8 * s + DC_NGRP] = ngrp;
        A.sys_cnt[8 * s + DC_NADJ] = nadj;
    }
}

struct DsFinish {
    const int* node_base; const int* row_base; const int* win_base; const int* slice_base;
    const unsigned* blk_base; const unsigned* grp_base; const unsigned* adjf_base;   // [n_sys+1]
    const int* row_node_loc; const int* node_row_loc; const int* loc_rowptr; const int* loc_adjptr; const int* loc_slw;
    const int* oldid; const int* adjc; const int* adj;
    int* node_row; int* row_node; unsigned* row_ptr; unsigned* adjf_ptr; int* adjf;
    int* win_sys; int* win_row0; int* win_n; int* win_slice0; unsigned* slice_ptr;
    int n_sys;
};

// (C) local numbering -> the global row / block / slice index spaces of the batch
__global__ void __launch_bounds__(DS_THREADS)
k_ds_finish(DsMesh M, DsFinish F) {
    const int s = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
    const int t0 = M.toff[s];
    const int nb0 = F.node_base[s], nn = F.node_base[s + 1] - nb0;
    const int rb0 = F.row_base[s], na = F.row_base[s + 1] - rb0;
    const int w0 = F.win_base[s], nwin = F.win_base[s + 1] - w0, sl0 = F.slice_base[s], nsl = F.slice_base[s + 1] - sl0;
    const unsigned kb0 = F.blk_base[s], gb0 = F.grp_base[s], ab0 = F.adjf_base[s];
    const int* nrl = F.node_row_loc + nb0; const int* rnl = F.row_node_loc + nb0;
    const int* lrp = F.loc_rowptr + nb0 + s; const int* lap = F.loc_adjptr + nb0 + s; const int* lsw = F.loc_slw + nb0 / 32 + 2 * s;
    const int* oldid = F.oldid + nb0; const int* adjc = F.adjc + nb0 + s; const int* adj = F.adj + 6 * (size_t)t0;
    for (int n = tid; n < nn; n += nth) F.node_row[nb0 + n] = nrl[n] < 0 ? -1 : rb0 + nrl[n];
    for (int r = tid; r < na; r += nth) {
        F.row_node[rb0 + r] = nb0 + rnl[r];
        F.row_ptr[rb0 + r] = kb0 + (unsigned)lrp[r];
        F.adjf_ptr[rb0 + r] = ab0 + (unsigned)lap[r];
        const int o = oldid[r];
        const int a0 = adjc[o], d = adjc[o + 1] - a0;
        for (int q = 0; q < d; ++q) F.adjf[(size_t)ab0 + lap[r] + q] = t0 + adj[a0 + q];
    }
    for (int w = tid; w < nwin; w += nth) {
        F.win_sys[w0 + w] = s; F.win_row0[w0 + w] = rb0 + w * RVE_WIN; F.win_n[w0 + w] = min(RVE_WIN, na - w * RVE_WIN);
        F.win_slice0[w0 + w] = sl0 + w * (RVE_WIN / 32);
    }
    for (int sl = tid; sl < nsl; sl += nth) F.slice_ptr[sl0 + sl] = gb0 + (unsigned)lsw[sl];
    if (s == F.n_sys - 1 && tid == 0) {
        F.row_ptr[rb0 + na] = F.blk_base[s + 1]; F.adjf_ptr[rb0 + na] = F.adjf_base[s + 1];
        F.slice_ptr[sl0 + nsl] = F.grp_base[s + 1];
    }
}

// ---------------------------------------------------------------------------------------------
// (D) Vref sample points: element and barycentric coordinates of the 4 nside + 1 nodes of the boundary mesh
//     (counter-clockwise from (x0,y0), centre last).  Candidates = elements whose bounding box touches the
//     outline of the Vref box or its centre; per point a warp takes the element with the largest minimal
//     barycentric coordinate (ties within 1e-12: lowest element id).  Replaces the bbox-tree point location
//     of usol.interpolate(...) (simulation_snapshots.py:56-58).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_ds_vref(DsMesh M, const double* __restrict__ box, const double* __restrict__ vbox, int nside, int* __restrict__ cand /*[NT]*/,
          int* __restrict__ samp_elem, double* __restrict__ samp_lam, int* __restrict__ sys_cnt) {
    __shared__ int s_n;
    const int s = blockIdx.x, tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, wv = tid >> 5, nw = nth >> 5;
    if (sys_cnt[8 * s + DC_ERR]) return;
    const int v0 = M.voff[s], t0 = M.toff[s], nt = M.toff[s + 1] - t0;
    const double* xy = M.xy + 2 * (size_t)v0;
    const int* tri = M.tri + 3 * (size_t)t0;
    int* cd = cand + t0;
    const double* vb = vbox ? vbox : box + 4 * s;
    const double Lx = box[4 * s + 2], Ly = box[4 * s + 3];
    const double e = 1e-9 * fmax(Lx, Ly), cxv = vb[0] + 0.5 * vb[2], cyv = vb[1] + 0.5 * vb[3];
    if (tid == 0) s_n = 0;
    __syncthreads();
    for (int t = tid; t < nt; t += nth) {
        double xa = 1e300, xb = -1e300, ya = 1e300, yb = -1e300;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int v = tri[3 * t + k];
            xa = fmin(xa, xy[2 * v]); xb = fmax(xb, xy[2 * v]); ya = fmin(ya, xy[2 * v + 1]); yb = fmax(yb, xy[2 * v + 1]);
        }
        const bool hits = xa <= vb[0] + vb[2] + e && xb >= vb[0] - e && ya <= vb[1] + vb[3] + e && yb >= vb[1] - e;
        const bool inside = xa > vb[0] + e && xb < vb[0] + vb[2] - e && ya > vb[1] + e && yb < vb[1] + vb[3] - e;
        const bool centre = xa <= cxv + e && xb >= cxv - e && ya <= cyv + e && yb >= cyv - e;
        if ((hits && !inside) || centre) cd[atomicAdd(&s_n, 1)] = t;
    }
    __syncthreads();
    const int nc = s_n, nvref = 4 * nside + 1;
    for (int k = wv; k < nvref; k += nw) {
        double px, py;
        if (k == 4 * nside) { px = cxv; py = cyv; }
        else {
            const int side = k / nside; const double f = (double)(k % nside) / nside;
            if (side == 0) { px = vb[0] + vb[2] * f; py = vb[1]; }
            else if (side == 1) { px = vb[0] + vb[2]; py = vb[1] + vb[3] * f; }
            else if (side == 2) { px = vb[0] + vb[2] * (1 - f); py = vb[1] + vb[3]; }
            else { px = vb[0]; py = vb[1] + vb[3] * (1 - f); }
        }
        auto lam = [&](int t, double& l1, double& l2) -> double {
            const double* p0 = xy + 2 * tri[3 * t]; const double* p1 = xy + 2 * tri[3 * t + 1]; const double* p2 = xy + 2 * tri[3 * t + 2];
            const double twoA = (p1[0] - p0[0]) * (p2[1] - p0[1]) - (p1[1] - p0[1]) * (p2[0] - p0[0]);
            l1 = ((px - p0[0]) * (p2[1] - p0[1]) - (py - p0[1]) * (p2[0] - p0[0])) / twoA;
            l2 = ((p1[0] - p0[0]) * (py - p0[1]) - (p1[1] - p0[1]) * (px - p0[0])) / twoA;
            return fmin(1.0 - l1 - l2, fmin(l1, l2));
        };
        double best = -1e300;
        for (int c = lane; c < nc; c += 32) { double l1, l2; best = fmax(best, lam(cd[c], l1, l2)); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, o));
        int bt = 0x7fffffff;
        for (int c = lane; c < nc; c += 32) { double l1, l2; const int t = cd[c]; if (lam(t, l1, l2) >= best - 1e-12) bt = min(bt, t); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bt = min(bt, __shfl_xor_sync(0xffffffffu, bt, o));
        if (lane == 0) {
            if (nc == 0 || best < -1e-8) { atomicMax(&sys_cnt[8 * s + DC_ERR], DE_VREF_OUTSIDE); }
            else {
                double l1, l2; lam(bt, l1, l2);
                samp_elem[(size_t)s * nvref + k] = t0 + bt;
                samp_lam[2 * ((size_t)s * nvref + k)] = l1; samp_lam[2 * ((size_t)s * nvref + k) + 1] = l2;
            }
        }
    }
}
