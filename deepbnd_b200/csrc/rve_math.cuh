// rve_math.cuh — P2 (quadratic, straight-edged) triangle arithmetic shared by host and device.
//
// Restates what FFC generates for the reference's forms
//   a = inner(sigma(u), eps(v)) dx,  sigma(u) = lam* div(u) I + 2 mu sym(grad u)
// (core/multiscale/micro_model_dnn.py:53, micro_model.py:49) on "CG 2" vector elements
// (others['polyorder'] = 2, micro_model.py:29).  Local node order: v0 v1 v2 m01 m12 m20.
#pragma once

#if defined(__CUDACC__)
#define RVE_HD __host__ __device__ __forceinline__
#else
#define RVE_HD inline
#endif

struct TriGeom {
    double gl[3][2];  // gradients of the barycentric coordinates
    double area;
};

RVE_HD void tri_geom(double x0, double y0, double x1, double y1, double x2, double y2, TriGeom& g) {
    double twoA = (x1 - x0) * (y2 - y0) - (y1 - y0) * (x2 - x0);
    double inv = 1.0 / twoA;
    g.area = 0.5 * twoA;
    g.gl[0][0] = (y1 - y2) * inv; g.gl[0][1] = (x2 - x1) * inv;
    g.gl[1][0] = (y2 - y0) * inv; g.gl[1][1] = (x0 - x2) * inv;
    g.gl[2][0] = (y0 - y1) * inv; g.gl[2][1] = (x1 - x0) * inv;
}

// Gradient of P2 shape function `a` at the q-th point of the edge-midpoint rule
// (points (1/2,1/2,0), (0,1/2,1/2), (1/2,0,1/2), weights area/3; exact for degree 2).
RVE_HD void p2_grad(const TriGeom& g, int q, int a, double& gx, double& gy) {
    // barycentric coordinates of the quadrature point
    double l0 = (q == 1) ? 0.0 : 0.5;
    double l1 = (q == 2) ? 0.0 : 0.5;
    double l2 = (q == 0) ? 0.0 : 0.5;
    double c0, c1, c2;  // grad = c0*gl0 + c1*gl1 + c2*gl2
    switch (a) {
        case 0: c0 = 4.0 * l0 - 1.0; c1 = 0.0; c2 = 0.0; break;
        case 1: c0 = 0.0; c1 = 4.0 * l1 - 1.0; c2 = 0.0; break;
        case 2: c0 = 0.0; c1 = 0.0; c2 = 4.0 * l2 - 1.0; break;
        case 3: c0 = 4.0 * l1; c1 = 4.0 * l0; c2 = 0.0; break;
        case 4: c0 = 0.0; c1 = 4.0 * l2; c2 = 4.0 * l1; break;
        default: c0 = 4.0 * l2; c1 = 0.0; c2 = 4.0 * l0; break;
    }
    gx = c0 * g.gl[0][0] + c1 * g.gl[1][0] + c2 * g.gl[2][0];
    gy = c0 * g.gl[0][1] + c1 * g.gl[1][1] + c2 * g.gl[2][1];
}

// 2x2 stiffness block K_ab[c][d] (row dof = (a,c), column dof = (b,d)):
//   G[c][d] = int d_c phi_a d_d phi_b ;  K = lam G + mu G^T + mu tr(G) I
RVE_HD void p2_stiff_block(const TriGeom& g, int a, int b, double lam, double mu, double K[4]) {
    double G00 = 0, G01 = 0, G10 = 0, G11 = 0;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        double ax, ay, bx, by;
        p2_grad(g, q, a, ax, ay);
        p2_grad(g, q, b, bx, by);
        G00 += ax * bx; G01 += ax * by; G10 += ay * bx; G11 += ay * by;
    }
    double w = g.area * (1.0 / 3.0);
    G00 *= w; G01 *= w; G10 *= w; G11 *= w;
    double tr = G00 + G11;
    K[0] = lam * G00 + mu * G00 + mu * tr;
    K[1] = lam * G01 + mu * G10;
    K[2] = lam * G10 + mu * G01;
    K[3] = lam * G11 + mu * G11 + mu * tr;
}

// int_T grad(phi_a) dx  (closed form of the 3-point rule)
RVE_HD void p2_grad_int(const TriGeom& g, int a, double& gx, double& gy) {
    double w = g.area * (1.0 / 3.0);
    if (a < 3) {
        gx = w * g.gl[a][0]; gy = w * g.gl[a][1];
    } else {
        int i = a - 3, j = (a - 2) % 3;
        gx = 4.0 * w * (g.gl[i][0] + g.gl[j][0]);
        gy = 4.0 * w * (g.gl[i][1] + g.gl[j][1]);
    }
}

// P2 shape functions at barycentric (l0,l1,l2)
RVE_HD void p2_shape(double l0, double l1, double l2, double N[6]) {
    N[0] = l0 * (2.0 * l0 - 1.0); N[1] = l1 * (2.0 * l1 - 1.0); N[2] = l2 * (2.0 * l2 - 1.0);
    N[3] = 4.0 * l0 * l1; N[4] = 4.0 * l1 * l2; N[5] = 4.0 * l2 * l0;
}

// Bilinear transfer fine node -> level-1 structured grid.  Returns the cell (ci,cj) and the local
// coordinates (s,t) in [0,1]; corner (ci+di, cj+dj) has weight (di?s:1-s)*(dj?t:1-t).
RVE_HD void coarse_cell(double x, double y, double x0, double y0, double invHx, double invHy,
                        int ncx, int ncy, int& ci, int& cj, double& s, double& t) {
    double fx = (x - x0) * invHx, fy = (y - y0) * invHy;
    ci = (int)fx; cj = (int)fy;
    if (ci < 0) ci = 0;
    if (cj < 0) cj = 0;
    if (ci > ncx - 1) ci = ncx - 1;
    if (cj > ncy - 1) cj = ncy - 1;
    s = fx - ci; t = fy - cj;
    if (s < 0.0) s = 0.0;
    if (t < 0.0) t = 0.0;
    if (s > 1.0) s = 1.0;
    if (t > 1.0) t = 1.0;
}

// pair index 0..20 -> (a,b) with a <= b over the 6 local nodes
RVE_HD void pair_ab(int p, int& a, int& b) {
    // rows: a=0 -> 6 pairs, a=1 -> 5, ...
    int base = 0;
    a = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        int len = 6 - k;
        if (p >= base && p < base + len) { a = k; b = k + (p - base); }
        base += len;
    }
}

// Matrix-free element operator: f = K_e u for one P2 element and one right-hand side, K_e never formed.
// geo = {g0x, g0y, g1x, g1y, w*lam, w*mu} with g_i = grad lambda_i (g2 = -g0 - g1) and w = area/3.  With the
// gradient table of the edge-midpoint rule, grad u(q) = (v0-v2) (x) g0 + (v1-v2) (x) g1, v_i = sum_a c[q][a][i] u_a,
// and the nodal forces are the transposed combination of t_i = sigma(q) g_i (same form as p2_stiff_block:
// sigma = lam* tr(grad u) I + mu (grad u + grad u^T), micro_model_dnn.py:53).  u, f: [6][2] in the local order
// v0 v1 v2 m01 m12 m20.
RVE_HD void p2_apply(const double geo[6], const double u[12], double f[12]) {
    const double g0x = geo[0], g0y = geo[1], g1x = geo[2], g1y = geo[3], wl = geo[4], wm = geo[5];
    double d0[3][2], d1[3][2];     // (v0 - v2), (v1 - v2) at the three quadrature points
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const double u0 = u[c], u1 = u[2 + c], u2 = u[4 + c], u3 = u[6 + c], u4 = u[8 + c], u5 = u[10 + c];
        const double m = 2.0 * (u3 - u4 - u5) + u2;
        d0[0][c] = u0 + m;                          d1[0][c] = u1 + m;
        d0[1][c] = 2.0 * (u3 + u5 - u4) - u0 - u2;  d1[1][c] = u1 - u2;
        d0[2][c] = u0 - u2;                         d1[2][c] = 2.0 * (u3 + u4 - u5) - u1 - u2;
    }
    double t0[3][2], t1[3][2];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const double H00 = d0[q][0] * g0x + d1[q][0] * g1x, H01 = d0[q][0] * g0y + d1[q][0] * g1y;
        const double H10 = d0[q][1] * g0x + d1[q][1] * g1x, H11 = d0[q][1] * g0y + d1[q][1] * g1y;
        const double tr = wl * (H00 + H11);
        const double s00 = tr + 2.0 * wm * H00, s11 = tr + 2.0 * wm * H11, s01 = wm * (H01 + H10);
        t0[q][0] = s00 * g0x + s01 * g0y; t0[q][1] = s01 * g0x + s11 * g0y;
        t1[q][0] = s00 * g1x + s01 * g1y; t1[q][1] = s01 * g1x + s11 * g1y;
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const double s = t0[0][c] + t1[0][c];       // -t2 at q0
        f[c] = t0[0][c] - t0[1][c] + t0[2][c];
        f[2 + c] = t1[0][c] + t1[1][c] - t1[2][c];
        f[4 + c] = s - (t0[1][c] + t1[1][c]) - (t0[2][c] + t1[2][c]);
        f[6 + c] = 2.0 * (s + t0[1][c] + t1[2][c]);
        f[8 + c] = 2.0 * (t1[2][c] - s - t0[1][c]);
        f[10 + c] = 2.0 * (t0[1][c] - s - t1[2][c]);
    }
}

RVE_HD void p2_apply_geo(const TriGeom& g, double lam, double mu, double geo[6]) {
    const double w = g.area * (1.0 / 3.0);
    geo[0] = g.gl[0][0]; geo[1] = g.gl[0][1]; geo[2] = g.gl[1][0]; geo[3] = g.gl[1][1]; geo[4] = w * lam; geo[5] = w * mu;
}
