// rve_symbolic.hpp — CPU restatement of the symbolic phase, used ONLY by the test hook rve_host_symbolic
// (the product runs this phase on the GPU: rve_devsym.cuh + k_sym_cols / k_sym_window / k_sym_coarse).
//
// For one RVE mesh: P2 numbering (vertices, then edges sorted by (va,vb) — the same order
// numpy.unique gives the oracle), boundary detection, the constraint map of the BC model
// (Dirichlet elimination for lin/dnn, periodic master map for per, identity for MR), active
// block rows ordered by level-1 coarse cell, node-block CSR pattern, windows / SELL slices / halo and
// restriction lists, Vref sample location — the device layout, checked for consistency by sym_selfcheck.
// The CPU tests compare it with the oracle's pattern, the GPU tests compare the device numbering with it.
// Mirrors what dolfin's DofMap / multiphenics' BlockDofMap / PeriodicBoundary build
// behind listMultiscaleModels[model](mesh, ...) (core/multiscale/micro_model.py:56-57).
// Also holds the shared constants (RVE_WIN, RowInfo), the level-1 grid rule (pick_nc) and parallel_for.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rve_b200.h"
#include "rve_math.cuh"

#define RVE_WIN 256   // rows per window = rows per CTA of the PCG kernels

// per active row: position in its level-1 coarse cell (fp32: preconditioner only) and the
// window-local slots of the 4 cell corners (0xffff = corner not in the coarse space)
struct RowInfo { float s, t; uint16_t slot[4]; };

struct SysSym {
    int nv = 0, nt = 0, ne = 0, nn = 0, na = 0, nb = 0, nbe = 0;
    double box[4];                   // x0,y0,Lx,Ly (mesh bounding box)
    std::vector<int> tri;            // [nt][3] CCW
    std::vector<int> elem_node;      // [nt][6] local P2 node ids
    std::vector<unsigned char> reg;  // [nt]
    std::vector<double> node_xy;     // [nn][2]
    std::vector<int> node_va, node_vb;
    std::vector<int> node_row;       // [nn] local active row or -1
    std::vector<int> node_bnd;       // [nn] index into the boundary-node list or -1
    std::vector<int> row_node;       // [na] master node of the row
    std::vector<unsigned> row_ptr;   // [na+1] local block offsets (CSR in the final numbering, host only)
    std::vector<int> bcol;           // local row ids
    // device layout: windows of <= RVE_WIN rows, SELL-32 slices, window-local 16-bit columns
    int nwin = 0, max_tile = 0, max_cn = 0;
    std::vector<int> win_n;          // [nwin] rows of the window
    std::vector<int> win_slice0;     // [nwin+1] first slice
    std::vector<int> win_halo0;      // [nwin+1] offset into halo
    std::vector<int> win_cn0;        // [nwin+1] offset into cn_node
    std::vector<unsigned> slice_ptr; // [nslices+1] offsets in 32-entry groups
    std::vector<uint16_t> scol;      // [groups][32] window-local column (own rows, then halo)
    std::vector<int> halo;           // local row ids gathered by each window
    std::vector<int> cn_node;        // level-1 grid nodes touched by each window
    std::vector<unsigned> cn_beg;    // [sum ncn + 1] offsets into cn_ent
    std::vector<uint16_t> cn_ent;    // (row_in_window << 2) | corner
    std::vector<RowInfo> rowinfo;    // [na] bilinear weights + window-local coarse slots
    size_t nblocks = 0;              // number of 2x2 blocks of the system
    std::vector<int> bnd_node;       // [nb]
    std::vector<double> bnd_s;       // [nb][2] Vref perimeter parameters of the end vertices
    std::vector<int> bedge;          // [nbe][3] boundary edges (a,b,m), outward normal to the right
    std::vector<double> bedge_nl;    // [nbe][2] normal * length
    std::vector<int> samp_elem;      // [nvref] local element id
    std::vector<double> samp_lam;    // [nvref][2] (l1,l2)
    double max_ext = 0;              // largest element bounding-box side
    double mean_ext = 0;             // mean element bounding-box side
    std::string err;
};

struct SymParams {
    int model;
    int ncx, ncy;        // level-1 grid
    int vref_nb;         // 0: none
    bool vref_box_given;
    double vref_box[4];
};

// pass 1: everything that does not depend on the level-1 grid
static void sym_pass1(SysSym& S, const double* xy, const int* tri, const int* region, int nv, int nt) {
    // S may be a re-used object of an earlier batch (its vectors keep their capacity): reset the appended lists
    S.err.clear(); S.bedge.clear(); S.bedge_nl.clear(); S.bnd_node.clear();
    S.nv = nv; S.nt = nt;
    S.tri.assign(tri, tri + 3 * (size_t)nt);
    int rmin = 1 << 30;
    for (int t = 0; t < nt; ++t) rmin = std::min(rmin, region[t]);
    S.reg.resize(nt);
    double xmin = 1e300, ymin = 1e300, xmax = -1e300, ymax = -1e300;
    for (int v = 0; v < nv; ++v) {
        xmin = std::min(xmin, xy[2 * v]); xmax = std::max(xmax, xy[2 * v]);
        ymin = std::min(ymin, xy[2 * v + 1]); ymax = std::max(ymax, xy[2 * v + 1]);
    }
    S.box[0] = xmin; S.box[1] = ymin; S.box[2] = xmax - xmin; S.box[3] = ymax - ymin;
    double ext = 0, ext_sum = 0;
    for (int t = 0; t < nt; ++t) {
        int r = region[t] - rmin;
        if (r < 0 || r > 3) { S.err = "region tag out of range (need max-min <= 3)"; return; }
        S.reg[t] = (unsigned char)r;
        int* T = &S.tri[3 * t];
        for (int k = 0; k < 3; ++k)
            if (T[k] < 0 || T[k] >= nv) { S.err = "triangle vertex id out of range"; return; }
        const double *p0 = xy + 2 * T[0], *p1 = xy + 2 * T[1], *p2 = xy + 2 * T[2];
        double a2 = (p1[0] - p0[0]) * (p2[1] - p0[1]) - (p1[1] - p0[1]) * (p2[0] - p0[0]);
        if (a2 == 0.0) { S.err = "degenerate triangle"; return; }
        if (a2 < 0) std::swap(T[1], T[2]);
        double ex = std::max({p0[0], p1[0], p2[0]}) - std::min({p0[0], p1[0], p2[0]});
        double ey = std::max({p0[1], p1[1], p2[1]}) - std::min({p0[1], p1[1], p2[1]});
        ext = std::max(ext, std::max(ex, ey));
        ext_sum += std::max(ex, ey);
    }
    S.max_ext = ext;
    S.mean_ext = nt > 0 ? ext_sum / nt : 0.0;
    // edges: bucket the 3*nt half-edges by their smaller vertex (counting sort), then order each
    // bucket (<= ~10 entries) by the larger vertex: same numbering as a global sort by (va,vb)
    std::vector<int> vptr(nv + 1, 0);
    for (int t = 0; t < nt; ++t)
        for (int k = 0; k < 3; ++k) {
            int a = S.tri[3 * t + k], b = S.tri[3 * t + (k + 1) % 3];
            vptr[std::min(a, b) + 1]++;
        }
    for (int v = 0; v < nv; ++v) vptr[v + 1] += vptr[v];
    std::vector<std::pair<int, int>> he(3 * (size_t)nt);          // (larger vertex, slot)
    {
        std::vector<int> vf(vptr.begin(), vptr.end() - 1);
        for (int t = 0; t < nt; ++t)
            for (int k = 0; k < 3; ++k) {
                int a = S.tri[3 * t + k], b = S.tri[3 * t + (k + 1) % 3];
                he[vf[std::min(a, b)]++] = {std::max(a, b), 3 * t + k};
            }
    }
    S.elem_node.resize(6 * (size_t)nt);
    std::vector<int> ecount;
    std::vector<uint64_t> ekeys;
    ecount.reserve(3 * (size_t)nt / 2 + nv); ekeys.reserve(3 * (size_t)nt / 2 + nv);
    for (int v = 0; v < nv; ++v) {
        std::pair<int, int>* b0 = he.data() + vptr[v];
        const int m = vptr[v + 1] - vptr[v];
        for (int i = 1; i < m; ++i) {
            const std::pair<int, int> x = b0[i];
            int j = i - 1;
            while (j >= 0 && b0[j] > x) { b0[j + 1] = b0[j]; --j; }
            b0[j + 1] = x;
        }
        for (int i = 0; i < m; ++i) {
            if (i == 0 || b0[i].first != b0[i - 1].first) {
                ekeys.push_back(((uint64_t)v << 32) | (uint64_t)b0[i].first); ecount.push_back(0);
            }
            ecount.back()++;
            const int t = b0[i].second / 3, k = b0[i].second % 3;
            S.elem_node[6 * t + 3 + k] = nv + (int)ekeys.size() - 1;
        }
    }
    S.ne = (int)ekeys.size();
    S.nn = nv + S.ne;
    for (int t = 0; t < nt; ++t)
        for (int k = 0; k < 3; ++k) S.elem_node[6 * t + k] = S.tri[3 * t + k];
    S.node_xy.resize(2 * (size_t)S.nn);
    S.node_va.resize(S.nn); S.node_vb.resize(S.nn);
    for (int v = 0; v < nv; ++v) {
        S.node_xy[2 * v] = xy[2 * v]; S.node_xy[2 * v + 1] = xy[2 * v + 1];
        S.node_va[v] = S.node_vb[v] = v;
    }
    for (int e = 0; e < S.ne; ++e) {
        int a = (int)(ekeys[e] >> 32), b = (int)(ekeys[e] & 0xffffffffu);
        S.node_xy[2 * (nv + e)] = 0.5 * (xy[2 * a] + xy[2 * b]);
        S.node_xy[2 * (nv + e) + 1] = 0.5 * (xy[2 * a + 1] + xy[2 * b + 1]);
        S.node_va[nv + e] = a; S.node_vb[nv + e] = b;
    }
    // boundary edges with outward normal (owning triangle is CCW -> outward is to the right of a->b)
    S.node_bnd.assign(S.nn, -1);
    for (int t = 0; t < nt; ++t)
        for (int k = 0; k < 3; ++k) {
            int m = S.elem_node[6 * t + 3 + k];
            if (ecount[m - nv] == 1) {
                int a = S.tri[3 * t + k], b = S.tri[3 * t + (k + 1) % 3];
                S.bedge.push_back(a); S.bedge.push_back(b); S.bedge.push_back(m);
                double dx = xy[2 * b] - xy[2 * a], dy = xy[2 * b + 1] - xy[2 * a + 1];
                S.bedge_nl.push_back(dy); S.bedge_nl.push_back(-dx);
                S.node_bnd[a] = S.node_bnd[b] = S.node_bnd[m] = 0;
            } else if (ecount[m - nv] != 2) {
                S.err = "non-manifold edge"; return;
            }
        }
    S.nbe = (int)S.bedge.size() / 3;
    for (int n = 0; n < S.nn; ++n)
        if (S.node_bnd[n] == 0) { S.node_bnd[n] = (int)S.bnd_node.size(); S.bnd_node.push_back(n); }
    S.nb = (int)S.bnd_node.size();
}

// Vref perimeter parameter of a point on the box boundary: node k of Vref sits at s = k.
static double perimeter_param(double x, double y, const double* vb, int nside) {
    double x0 = vb[0], y0 = vb[1], Lx = vb[2], Ly = vb[3];
    double tol = 1e-9 * std::max(Lx, Ly);
    if (std::fabs(y - y0) < tol && x < x0 + Lx - tol) return (x - x0) / Lx * nside;
    if (std::fabs(x - (x0 + Lx)) < tol && y < y0 + Ly - tol) return nside + (y - y0) / Ly * nside;
    if (std::fabs(y - (y0 + Ly)) < tol && x > x0 + tol) return 2 * nside + (x0 + Lx - x) / Lx * nside;
    if (std::fabs(x - x0) < tol) return 3 * nside + (y0 + Ly - y) / Ly * nside;
    return -1.0;
}

// pass 2: constraint map, rows ordered by level-1 cell, node-block pattern, Vref samples
static void sym_pass2(SysSym& S, const SymParams& P) {
    const int nn = S.nn, nt = S.nt;
    const double x0 = S.box[0], y0 = S.box[1], Lx = S.box[2], Ly = S.box[3];
    const double tol = 1e-9 * std::max(Lx, Ly);
    std::vector<int> master(nn);
    for (int n = 0; n < nn; ++n) master[n] = n;
    if (P.model == RVE_MODEL_LIN || P.model == RVE_MODEL_DNN) {
        for (int i = 0; i < S.nb; ++i) master[S.bnd_node[i]] = -1;
    } else if (P.model == RVE_MODEL_PER) {
        // slaves: x = x1 -> x0, y = y1 -> y0 (corners -> (x0,y0)); match by the other coordinate
        std::vector<std::pair<double, int>> left, bottom;
        for (int i = 0; i < S.nb; ++i) {
            int n = S.bnd_node[i];
            double x = S.node_xy[2 * n], y = S.node_xy[2 * n + 1];
            if (std::fabs(x - x0) < tol) left.push_back({y, n});
            if (std::fabs(y - y0) < tol) bottom.push_back({x, n});
        }
        std::sort(left.begin(), left.end());
        std::sort(bottom.begin(), bottom.end());
        auto find = [&](const std::vector<std::pair<double, int>>& L, double key) -> int {
            auto it = std::lower_bound(L.begin(), L.end(), std::make_pair(key - 1e-7 * std::max(Lx, Ly), -1));
            if (it == L.end() || std::fabs(it->first - key) > 1e-7 * std::max(Lx, Ly)) return -1;
            return it->second;
        };
        for (int i = 0; i < S.nb; ++i) {
            int n = S.bnd_node[i];
            double x = S.node_xy[2 * n], y = S.node_xy[2 * n + 1];
            bool onr = std::fabs(x - (x0 + Lx)) < tol, ont = std::fabs(y - (y0 + Ly)) < tol;
            int m = n;
            if (onr) { m = find(left, y); if (m < 0) { S.err = "periodic: left/right boundary nodes do not match"; return; } y = S.node_xy[2 * m + 1]; }
            if (ont) {
                double xm = S.node_xy[2 * m];
                m = find(bottom, xm);
                if (m < 0) { S.err = "periodic: bottom/top boundary nodes do not match"; return; }
            }
            master[n] = m;
        }
        for (int it = 0; it < 3; ++it)
            for (int n = 0; n < nn; ++n) if (master[n] >= 0) master[n] = master[master[n]];
    }
    // ---- active rows ordered along the Morton curve of the level-1 cells (compact row windows)
    const int ncx = P.ncx, ncy = P.ncy;
    const double invHx = ncx / Lx, invHy = ncy / Ly;
    auto spread = [](unsigned v) -> uint64_t {
        uint64_t x = v & 0xffffu;
        x = (x | (x << 8)) & 0x00ff00ffull; x = (x | (x << 4)) & 0x0f0f0f0full;
        x = (x | (x << 2)) & 0x33333333ull; x = (x | (x << 1)) & 0x55555555ull;
        return x;
    };
    // counting sort of the active nodes by the Morton rank of their cell (ties keep node order)
    std::vector<int> cellrank((size_t)ncx * ncy);
    {
        std::vector<std::pair<uint64_t, int>> ck((size_t)ncx * ncy);
        for (int cj = 0; cj < ncy; ++cj)
            for (int ci = 0; ci < ncx; ++ci) ck[(size_t)cj * ncx + ci] = {spread((unsigned)ci) | (spread((unsigned)cj) << 1), cj * ncx + ci};
        std::sort(ck.begin(), ck.end());
        for (size_t k = 0; k < ck.size(); ++k) cellrank[ck[k].second] = (int)k;
    }
    std::vector<int> ncell(nn, -1), cptr((size_t)ncx * ncy + 1, 0);
    for (int n = 0; n < nn; ++n)
        if (master[n] == n) {
            int ci, cj; double s, t;
            coarse_cell(S.node_xy[2 * n], S.node_xy[2 * n + 1], x0, y0, invHx, invHy, ncx, ncy, ci, cj, s, t);
            ncell[n] = cellrank[cj * ncx + ci];
            cptr[ncell[n] + 1]++;
        }
    for (size_t c = 0; c < (size_t)ncx * ncy; ++c) cptr[c + 1] += cptr[c];
    std::vector<std::pair<uint64_t, int>> keyed(cptr.back());
    for (int n = 0; n < nn; ++n)
        if (ncell[n] >= 0) keyed[cptr[ncell[n]]++] = {(uint64_t)ncell[n], n};
    S.na = (int)keyed.size();
    const int na = S.na;
    // preliminary numbering = Morton order; row lengths by a marker-array dedupe over the adjacent elements
    std::vector<int> prow(nn, -1), pnode(na);
    for (int r = 0; r < na; ++r) { pnode[r] = keyed[r].second; prow[keyed[r].second] = r; }
    std::vector<int> nrow0(nn);
    for (int n = 0; n < nn; ++n) nrow0[n] = master[n] < 0 ? -1 : prow[master[n]];
    std::vector<int> erow(6 * (size_t)nt);                 // element -> preliminary rows
    std::vector<int> adj_ptr(na + 1, 0);
    for (int t = 0; t < nt; ++t)
        for (int k = 0; k < 6; ++k) {
            const int r = nrow0[S.elem_node[6 * t + k]];
            erow[6 * (size_t)t + k] = r;
            if (r >= 0) adj_ptr[r + 1]++;
        }
    for (int r = 0; r < na; ++r) adj_ptr[r + 1] += adj_ptr[r];
    std::vector<int> adj(adj_ptr[na]), fill(adj_ptr.begin(), adj_ptr.end() - 1);
    for (int t = 0; t < nt; ++t)
        for (int k = 0; k < 6; ++k) {
            const int r = erow[6 * (size_t)t + k];
            if (r >= 0) adj[fill[r]++] = t;
        }
    std::vector<int> mark(na, -1), len0(na), cap0(na);
    {
        // row CAPACITIES from the element fans alone (no dedupe; the GPU finds the true columns and lengths): a vertex
        // with d elements around it couples to 3d+1 nodes in a closed fan and to 3d+3 in an open one (rows that gather a
        // mesh-boundary node), a mid-edge node to 9 (two elements) or 6 (one).  Inactive neighbours only shorten a row.
        // Measured SELL padding with these widths: +0.3..1 % over the exact ones.  The window sort below always uses
        // the capacities, so that this numbering is the one of the device symbolic phase (rve_devsym.cuh).
        std::vector<unsigned char> rowb(na, 0);
        for (int i = 0; i < S.nb; ++i) {
            const int r = nrow0[S.bnd_node[i]];
            if (r >= 0) rowb[r] = 1;
        }
        for (int r = 0; r < na; ++r) {
            const int d = adj_ptr[r + 1] - adj_ptr[r];
            cap0[r] = pnode[r] < S.nv ? 3 * d + (rowb[r] ? 3 : 1) : (d >= 2 ? 9 : 6);
        }
    }
    for (int r = 0; r < na; ++r) {       // exact row lengths by a marker-array dedupe over the adjacent elements
        int cnt = 0;
        for (int i = adj_ptr[r]; i < adj_ptr[r + 1]; ++i) {
            const int* er = &erow[6 * (size_t)adj[i]];
            for (int k = 0; k < 6; ++k) {
                const int c = er[k];
                if (c >= 0 && mark[c] != r) { mark[c] = r; ++cnt; }
            }
        }
        len0[r] = cnt;
    }
    // windows of RVE_WIN consecutive rows; inside a window rows are sorted by descending length so
    // that the 32-row SELL slices are nearly uniform (vertex rows ~19 blocks, mid-edge rows ~9)
    const int nwin = (na + RVE_WIN - 1) / RVE_WIN;
    S.nwin = nwin;
    std::vector<int> newid(na), oldid(na);
    {
        int bucket[64];
        for (int w = 0; w < nwin; ++w) {                    // stable counting sort by length (lengths < 64)
            const int r0 = w * RVE_WIN, r1 = std::min(na, r0 + RVE_WIN);
            for (int k = 0; k < 64; ++k) bucket[k] = 0;
            for (int r = r0; r < r1; ++r) bucket[63 - std::min(cap0[r], 63)]++;
            int acc = 0;
            for (int k = 0; k < 64; ++k) { const int c = bucket[k]; bucket[k] = acc; acc += c; }
            for (int r = r0; r < r1; ++r) {
                const int pos = r0 + bucket[63 - std::min(cap0[r], 63)]++;
                newid[r] = pos; oldid[pos] = r;
            }
        }
    }
    S.row_node.resize(na);
    for (int r = 0; r < na; ++r) S.row_node[r] = pnode[oldid[r]];
    S.node_row.resize(nn);
    for (int n = 0; n < nn; ++n) S.node_row[n] = nrow0[n] < 0 ? -1 : newid[nrow0[n]];
    // CSR row pointer and block columns in the final numbering
    S.row_ptr.assign(na + 1, 0);
    for (int r = 0; r < na; ++r) S.row_ptr[r + 1] = S.row_ptr[r] + (unsigned)len0[oldid[r]];
    S.nblocks = S.row_ptr[na];
    S.bcol.resize(S.row_ptr[na]);
    std::fill(mark.begin(), mark.end(), -1);
    for (int r = 0; r < na; ++r) {
        const int o = oldid[r];
        int* dst = &S.bcol[S.row_ptr[r]];
        int cnt = 0;
        for (int i = adj_ptr[o]; i < adj_ptr[o + 1]; ++i) {
            const int* er = &erow[6 * (size_t)adj[i]];
            for (int k = 0; k < 6; ++k) {
                const int c = er[k];
                if (c >= 0 && mark[c] != r) { mark[c] = r; dst[cnt++] = newid[c]; }
            }
        }
    }
    // per window: halo list, SELL-32 slices with window-local 16-bit columns, coarse-node lists
    S.win_n.assign(nwin, 0); S.win_slice0.assign(nwin + 1, 0); S.win_halo0.assign(nwin + 1, 0); S.win_cn0.assign(nwin + 1, 0);
    S.slice_ptr.assign(1, 0u);
    S.scol.clear(); S.halo.clear(); S.cn_node.clear(); S.cn_beg.assign(1, 0u); S.cn_ent.clear();
    S.scol.reserve((size_t)S.bcol.size() + S.bcol.size() / 6);
    S.rowinfo.resize(na);
    S.max_tile = 0; S.max_cn = 0;
    const int nxg = ncx + 1;
    auto cnode = [&](int i, int j) -> int {   // same validity rules as grid_node() on the device
        if (P.model == RVE_MODEL_PER) { if (i >= ncx) i -= ncx; if (j >= ncy) j -= ncy; return j * nxg + i; }   // i in [0, ncx], j in [0, ncy]
        const int lo = (P.model == RVE_MODEL_MR) ? 0 : 1, hix = (P.model == RVE_MODEL_MR) ? ncx : ncx - 1,
                  hiy = (P.model == RVE_MODEL_MR) ? ncy : ncy - 1;
        if (i < lo || i > hix || j < lo || j > hiy) return -1;
        return j * nxg + i;
    };
    std::vector<int> hal, cns, hpos(na, -1), cpos((size_t)nxg * (ncy + 1), -1), ccount, cfill;
    std::vector<int> corner(4 * (size_t)RVE_WIN);
    for (int w = 0; w < nwin; ++w) {
        const int r0 = w * RVE_WIN, r1 = std::min(na, r0 + RVE_WIN), wn = r1 - r0;
        S.win_n[w] = wn;
        const int nsl = (wn + 31) / 32;
        hal.clear();
        for (unsigned k = S.row_ptr[r0]; k < S.row_ptr[r1]; ++k) {
            const int c = S.bcol[k];
            if ((c < r0 || c >= r1) && hpos[c] < 0) { hpos[c] = 0; hal.push_back(c); }
        }
        std::sort(hal.begin(), hal.end());
        for (size_t k = 0; k < hal.size(); ++k) hpos[hal[k]] = wn + (int)k;
        if (wn + (int)hal.size() > 65535) { S.err = "window halo too large for 16-bit local columns"; return; }
        S.max_tile = std::max(S.max_tile, wn + (int)hal.size());
        S.halo.insert(S.halo.end(), hal.begin(), hal.end());
        S.win_halo0[w + 1] = (int)S.halo.size();
        for (int sl = 0; sl < nsl; ++sl) {
            const int a0 = r0 + 32 * sl, a1 = std::min(r1, a0 + 32);
            int width = 0;
            for (int r = a0; r < a1; ++r) width = std::max(width, (int)(S.row_ptr[r + 1] - S.row_ptr[r]));
            const size_t base = S.scol.size();
            S.scol.resize(base + (size_t)width * 32);
            uint16_t* sc = &S.scol[base];
            for (int lane = 0; lane < 32; ++lane) {
                const int r = a0 + lane;
                const int len = r < a1 ? (int)(S.row_ptr[r + 1] - S.row_ptr[r]) : 0;
                const uint16_t self = (uint16_t)(r < a1 ? r - r0 : 0);
                const int* cols = r < a1 ? &S.bcol[S.row_ptr[r]] : nullptr;
                for (int j = 0; j < len; ++j) {
                    const int c = cols[j];
                    sc[(size_t)j * 32 + lane] = (uint16_t)((c >= r0 && c < r1) ? c - r0 : hpos[c]);
                }
                for (int j = len; j < width; ++j) sc[(size_t)j * 32 + lane] = self;   // padding: zero block
            }
            S.slice_ptr.push_back(S.slice_ptr.back() + (unsigned)width);
        }
        for (size_t k = 0; k < hal.size(); ++k) hpos[hal[k]] = -1;
        S.win_slice0[w + 1] = S.win_slice0[w] + nsl;
        // coarse nodes touched by the window and the (row, corner) lists of each of them
        cns.clear();
        for (int r = r0; r < r1; ++r) {
            const int n = S.row_node[r];
            int ci, cj; double s, t;
            coarse_cell(S.node_xy[2 * n], S.node_xy[2 * n + 1], x0, y0, invHx, invHy, ncx, ncy, ci, cj, s, t);
            S.rowinfo[r].s = (float)s; S.rowinfo[r].t = (float)t;
            for (int c = 0; c < 4; ++c) {
                const int nd = cnode(ci + (c & 1), cj + (c >> 1));
                corner[4 * (size_t)(r - r0) + c] = nd;
                if (nd >= 0 && cpos[nd] < 0) { cpos[nd] = 0; cns.push_back(nd); }
            }
        }
        std::sort(cns.begin(), cns.end());
        const int ncn = (int)cns.size();
        for (int j = 0; j < ncn; ++j) cpos[cns[j]] = j;
        S.max_cn = std::max(S.max_cn, ncn);
        S.cn_node.insert(S.cn_node.end(), cns.begin(), cns.end());
        ccount.assign(ncn + 1, 0);
        for (int i = 0; i < 4 * wn; ++i) if (corner[i] >= 0) ccount[cpos[corner[i]] + 1]++;
        for (int j = 0; j < ncn; ++j) ccount[j + 1] += ccount[j];
        const size_t ebase = S.cn_ent.size();
        S.cn_ent.resize(ebase + ccount[ncn]);
        cfill.assign(ccount.begin(), ccount.end() - 1);
        for (int rl = 0; rl < wn; ++rl)
            for (int c = 0; c < 4; ++c) {
                const int nd = corner[4 * (size_t)rl + c];
                uint16_t slot = 0xffffu;
                if (nd >= 0) {
                    slot = (uint16_t)cpos[nd];
                    S.cn_ent[ebase + cfill[slot]++] = (uint16_t)((rl << 2) | c);
                }
                S.rowinfo[r0 + rl].slot[c] = slot;
            }
        for (int j = 0; j < ncn; ++j) { S.cn_beg.push_back((unsigned)(ebase + ccount[j + 1])); cpos[cns[j]] = -1; }
        S.win_cn0[w + 1] = (int)S.cn_node.size();
    }
    // Dirichlet boundary parameters / Vref samples
    double vb[4];
    for (int k = 0; k < 4; ++k) vb[k] = P.vref_box_given ? P.vref_box[k] : S.box[k];
    const int nside = P.vref_nb > 0 ? P.vref_nb - 1 : 0;
    if (P.model == RVE_MODEL_DNN) {
        if (nside == 0) { S.err = "dnn model needs vref_nb > 0"; return; }
        S.bnd_s.resize(2 * (size_t)S.nb);
        for (int i = 0; i < S.nb; ++i) {
            int n = S.bnd_node[i];
            int a = S.node_va[n], b = S.node_vb[n];
            double sa = perimeter_param(S.node_xy[2 * a], S.node_xy[2 * a + 1], vb, nside);
            double sb = perimeter_param(S.node_xy[2 * b], S.node_xy[2 * b + 1], vb, nside);
            if (sa < 0 || sb < 0) { S.err = "dnn: mesh boundary does not coincide with the Vref box"; return; }
            S.bnd_s[2 * i] = sa; S.bnd_s[2 * i + 1] = sb;
        }
    }
    if (nside > 0) {
        const int nvref = 4 * nside + 1;
        S.samp_elem.assign(nvref, -1);
        S.samp_lam.assign(2 * (size_t)nvref, 0.0);
        // bin elements on a uniform grid for point location
        const int gb = std::max(1, (int)std::sqrt((double)nt / 4.0));
        std::vector<int> bptr((size_t)gb * gb + 1, 0), blist;
        auto cellrange = [&](int t, int& i0, int& i1, int& j0, int& j1) {
            double xa = 1e300, xb = -1e300, ya = 1e300, yb = -1e300;
            for (int k = 0; k < 3; ++k) {
                int v = S.tri[3 * t + k];
                xa = std::min(xa, S.node_xy[2 * v]); xb = std::max(xb, S.node_xy[2 * v]);
                ya = std::min(ya, S.node_xy[2 * v + 1]); yb = std::max(yb, S.node_xy[2 * v + 1]);
            }
            i0 = std::min(std::max((int)((xa - x0) / Lx * gb - 1e-9), 0), gb - 1);
            i1 = std::min(std::max((int)((xb - x0) / Lx * gb + 1e-9), 0), gb - 1);
            j0 = std::min(std::max((int)((ya - y0) / Ly * gb - 1e-9), 0), gb - 1);
            j1 = std::min(std::max((int)((yb - y0) / Ly * gb + 1e-9), 0), gb - 1);
            // only elements that can hold a sample point are binned: those touching the outline of the
            // Vref box, or its centre
            const double e = 1e-9 * std::max(Lx, Ly), cxv = vb[0] + 0.5 * vb[2], cyv = vb[1] + 0.5 * vb[3];
            const bool hits = xa <= vb[0] + vb[2] + e && xb >= vb[0] - e && ya <= vb[1] + vb[3] + e && yb >= vb[1] - e;
            const bool inside = xa > vb[0] + e && xb < vb[0] + vb[2] - e && ya > vb[1] + e && yb < vb[1] + vb[3] - e;
            const bool centre = xa <= cxv + e && xb >= cxv - e && ya <= cyv + e && yb >= cyv - e;
            if (!((hits && !inside) || centre)) { i0 = 1; i1 = 0; }
        };
        for (int pass = 0; pass < 2; ++pass) {
            std::vector<int> f(bptr.begin(), bptr.end() - 1);
            for (int t = 0; t < nt; ++t) {
                int i0, i1, j0, j1;
                cellrange(t, i0, i1, j0, j1);
                for (int j = j0; j <= j1; ++j)
                    for (int i = i0; i <= i1; ++i) {
                        if (pass == 0) bptr[(size_t)j * gb + i + 1]++;
                        else blist[f[(size_t)j * gb + i]++] = t;
                    }
            }
            if (pass == 0) {
                for (size_t c = 0; c < (size_t)gb * gb; ++c) bptr[c + 1] += bptr[c];
                blist.resize(bptr.back());
            }
        }
        for (int k = 0; k < nvref; ++k) {
            double px, py;
            if (k == 4 * nside) { px = vb[0] + 0.5 * vb[2]; py = vb[1] + 0.5 * vb[3]; }
            else {
                int side = k / nside; double f = (double)(k % nside) / nside;
                if (side == 0) { px = vb[0] + vb[2] * f; py = vb[1]; }
                else if (side == 1) { px = vb[0] + vb[2]; py = vb[1] + vb[3] * f; }
                else if (side == 2) { px = vb[0] + vb[2] * (1 - f); py = vb[1] + vb[3]; }
                else { px = vb[0]; py = vb[1] + vb[3] * (1 - f); }
            }
            int bi = std::min(std::max((int)((px - x0) / Lx * gb), 0), gb - 1);
            int bj = std::min(std::max((int)((py - y0) / Ly * gb), 0), gb - 1);
            double best = -1e300; int bt = -1; double bl1 = 0, bl2 = 0;
            for (int dj = -1; dj <= 1; ++dj)
                for (int di = -1; di <= 1; ++di) {
                    int i = bi + di, j = bj + dj;
                    if (i < 0 || j < 0 || i >= gb || j >= gb) continue;
                    for (int q = bptr[(size_t)j * gb + i]; q < bptr[(size_t)j * gb + i + 1]; ++q) {
                        int t = blist[q];
                        const double* p0 = &S.node_xy[2 * S.tri[3 * t]];
                        const double* p1 = &S.node_xy[2 * S.tri[3 * t + 1]];
                        const double* p2 = &S.node_xy[2 * S.tri[3 * t + 2]];
                        double twoA = (p1[0] - p0[0]) * (p2[1] - p0[1]) - (p1[1] - p0[1]) * (p2[0] - p0[0]);
                        double l1 = ((px - p0[0]) * (p2[1] - p0[1]) - (py - p0[1]) * (p2[0] - p0[0])) / twoA;
                        double l2 = ((p1[0] - p0[0]) * (py - p0[1]) - (p1[1] - p0[1]) * (px - p0[0])) / twoA;
                        double l0 = 1.0 - l1 - l2;
                        double mn = std::min(l0, std::min(l1, l2));
                        // deterministic tie-break: best min-lambda, then lowest element id
                        if (mn > best + 1e-12) { best = mn; bt = t; bl1 = l1; bl2 = l2; }
                    }
                }
            if (bt < 0 || best < -1e-8) { S.err = "Vref sample point outside the mesh"; return; }
            S.samp_elem[k] = bt; S.samp_lam[2 * k] = bl1; S.samp_lam[2 * k + 1] = bl2;
        }
    }
}

// Consistency check of the device layout against the host CSR pattern (used by the CPU test hook):
// every CSR block must sit in the SELL slot of its row with the right window-local column, padding
// must point at the row itself, halo lists must be sorted and outside their window, and every
// (row, corner) pair must appear exactly once in the coarse-node lists.  Returns "" when consistent.
static std::string sym_selfcheck(const SysSym& S) {
    const int na = S.na;
    if (S.nwin != (na + RVE_WIN - 1) / RVE_WIN) return "window count";
    size_t npairs = 0;
    for (int w = 0; w < S.nwin; ++w) {
        const int r0 = w * RVE_WIN, wn = S.win_n[w];
        if (wn != std::min(RVE_WIN, na - r0)) return "window size";
        const int h0 = S.win_halo0[w], h1 = S.win_halo0[w + 1];
        for (int k = h0; k < h1; ++k) {
            if (S.halo[k] >= r0 && S.halo[k] < r0 + wn) return "halo inside window";
            if (k > h0 && S.halo[k] <= S.halo[k - 1]) return "halo not sorted";
        }
        for (int rl = 0; rl < wn; ++rl) {
            const int r = r0 + rl, lane = rl & 31, slice = S.win_slice0[w] + (rl >> 5);
            const unsigned g0 = S.slice_ptr[slice], g1 = S.slice_ptr[slice + 1];
            const int len = (int)(S.row_ptr[r + 1] - S.row_ptr[r]);
            if ((int)(g1 - g0) < len) return "slice narrower than a row";
            for (unsigned g = g0; g < g1; ++g) {
                const int lc = S.scol[(size_t)g * 32 + lane];
                const int j = (int)(g - g0);
                if (j >= len) { if (lc != rl) return "padding column"; continue; }
                const int c = lc < wn ? r0 + lc : (lc - wn < h1 - h0 ? S.halo[h0 + lc - wn] : -1);
                if (c != S.bcol[S.row_ptr[r] + j]) return "SELL column mismatch";
            }
        }
        const int c0 = S.win_cn0[w], c1 = S.win_cn0[w + 1];
        std::vector<int> seen(4 * (size_t)wn, 0);
        for (int j = c0; j < c1; ++j) {
            if (j > c0 && S.cn_node[j] <= S.cn_node[j - 1]) return "coarse nodes not sorted";
            for (unsigned e = S.cn_beg[j]; e < S.cn_beg[j + 1]; ++e) {
                const int rl = S.cn_ent[e] >> 2, c = S.cn_ent[e] & 3;
                if (rl >= wn) return "restriction entry row";
                if (S.rowinfo[r0 + rl].slot[c] != j - c0) return "restriction entry slot";
                seen[4 * (size_t)rl + c]++; ++npairs;
            }
        }
        for (int rl = 0; rl < wn; ++rl)
            for (int c = 0; c < 4; ++c) {
                const bool valid = S.rowinfo[r0 + rl].slot[c] != 0xffff;
                if (seen[4 * (size_t)rl + c] != (valid ? 1 : 0)) return "restriction pair count";
            }
    }
    if (npairs != S.cn_ent.size()) return "restriction entry total";
    return "";
}

// level-1 grid size: largest n = m*2^j (m in {2,3}, j>=0), n >= 2, whose cell H = L/n is at least
// RVE_CELL_FACTOR mean element extents.  Elements larger than a cell are allowed: the Galerkin kernel
// clamps the few nodes that fall outside the element's 3x3 coarse-node window (k_galerkin1), which keeps
// the coarse operator a sum of SPSD element terms.  Finer grids cut the PCG iterations (full meshes:
// n=32 -> 117, n=48 -> 80, n=64 -> 63 iterations) at a growing V-cycle cost; 1.1 balances the two
// (lcar = 2/30: n = 24 on the 2x2 'reduced' box, n = 64 on the 6x6 'full' box).
#define RVE_CELL_FACTOR 1.1
static int pick_nc(double L, double mean_ext) {
    int best = 2;
    double lim = L / (RVE_CELL_FACTOR * mean_ext * 1.0000001);
    for (int m = 2; m <= 3; ++m)
        for (int n = m; n <= 4096; n *= 2)
            if ((double)n <= lim) best = std::max(best, n);
    return best;
}

template <class F>
static void parallel_for(int n, F f) {
    unsigned hw = std::thread::hardware_concurrency();
    // leave two cores to the thread that is feeding kernels of the previous batch to the GPU
    int nth = (int)std::min<unsigned>(hw > 3 ? hw - 2 : (hw ? hw : 4), 64);
    if (const char* e = std::getenv("RVE_HOST_THREADS")) { const int v = std::atoi(e); if (v > 0) nth = v; }
    nth = std::min(nth, std::max(1, n / 4));
    if (nth <= 1) { for (int i = 0; i < n; ++i) f(i); return; }
    std::vector<std::thread> th;
    for (int w = 0; w < nth; ++w)
        th.emplace_back([=]() { for (int i = w; i < n; i += nth) f(i); });
    for (auto& t : th) t.join();
}
