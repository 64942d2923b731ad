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
