/* rve_b200.h — C-ABI of the B200 batched RVE solver (librve_b200.so).
 *
 * The reference (felipefr/deepbnd) has no FFI: its hot path is Python calling dolfin /
 * multiphenics / PETSc LU one RVE at a time.  Each entry point below replaces a group of
 * reference calls for a whole BATCH of RVEs; the Python mirror classes in deepbnd_b200/
 * bind these through ctypes.
 *
 * Conventions: every pointer is a HOST pointer owned by the caller; the library owns all device
 * memory behind the opaque handle; every call returns 0 on success or a negative RVE_E_* code
 * (message via rve_last_error); no exceptions cross the boundary; one handle per GPU; calls on
 * one handle are not re-entrant, distinct handles may be driven from distinct threads/processes.
 * Arithmetic is fp64, indices int32.  Load cases use Voigt macro strains [e11, e22, 2*e12]
 * (engineering shear, micmacsfenics macro_strain).
 */
#ifndef RVE_B200_H
#define RVE_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rve_batch_s* rve_handle_t;

/* boundary-condition models: keys of mscm.listMultiscaleModels used at
 * core/multiscale/micro_model.py:31 and core/multiscale/micro_model_dnn.py:33 */
#define RVE_MODEL_LIN 0 /* 'lin': zero (or given) Dirichlet fluctuation on the whole boundary  */
#define RVE_MODEL_DNN 1 /* 'dnn': Dirichlet fluctuation predicted on Vref (bcs.hd5 u0,u1,u2)    */
#define RVE_MODEL_PER 2 /* 'per': periodic fluctuation + zero average                            */
#define RVE_MODEL_MR  3 /* 'MR' : minimally restrictive: int u dx = 0, int u (x) n ds = 0        */

#define RVE_PRECOND_JACOBI 0   /* 2x2 node-block Jacobi                                            */
#define RVE_PRECOND_TWOLEVEL 1 /* block Jacobi + structured-grid Galerkin V-cycle (additive)       */

#define RVE_E_ARG -1
#define RVE_E_CUDA -2
#define RVE_E_STATE -3
#define RVE_E_MESH -4
#define RVE_E_NOCONV -5 /* at least one system did not reach rtol within maxit (see status[])  */

#define RVE_MAX_RHS 4

/* replaces: object construction MicroModel.__init__ / MicroConstitutiveModelDNN.__init__
 * (micro_model.py:24-36, micro_model_dnn.py:26-38) — one handle holds a whole batch. */
int rve_create(int device, rve_handle_t* out);
int rve_destroy(rve_handle_t h);
/* h may be NULL: returns the message of the last failed rve_create on this thread. */
const char* rve_last_error(rve_handle_t h);

/* replaces: readMesh (micro_model.py:38-49, micro_model_dnn.py:40-48: mesh, bbox x0..y1, Lame
 * table via getLameInclusions fenics_utils.py:41-53 + myCoeff wrapper_expression.py:15-56) and
 * the symbolic half of mp.block_assemble / the [EXT] formulation's constraint set-up.
 *   n_sys                  number of RVEs (ragged)
 *   vert_off[n_sys+1]      prefix offsets into xy   (vertices of system s: vert_off[s]..vert_off[s+1])
 *   tri_off[n_sys+1]       prefix offsets into tri/region
 *   xy[.][2]               vertex coordinates (straight-edged triangles; P2 nodes are built inside)
 *   tri[.][3]              LOCAL vertex ids of each triangle (any orientation)
 *   region[.]              physical tag per triangle (0,1 reduced; 0..3 full); min is subtracted
 *   mat[4][2]              [lambda*, mu] per region-min(region)
 *   model                  RVE_MODEL_*
 *   vref_nb                Nb of degeneratedBoundaryRectangleMesh (21 in the reference), 0 = no Vref
 *   vref_box[4]            x0,y0,Lx,Ly of Vref (NULL: the mesh bounding box)
 * Vref DOF order is the library's own: node k = 0..4(Nb-1)-1 counter-clockwise from (x0,y0),
 * centre node last, DOF = 2*node + comp.
 * The raw arrays are copied to the GPU and the whole symbolic phase (P2 numbering, boundary detection, constraint
 * maps, row order, sparsity pattern, Vref point location) runs there; invalid meshes come back as RVE_E_MESH with
 * "system <s>: <reason>" in rve_last_error.  All systems of a batch must share the box size. */
int rve_set_batch(rve_handle_t h, int32_t n_sys, const int64_t* vert_off, const int64_t* tri_off,
                  const double* xy, const int32_t* tri, const int32_t* region, const double* mat,
                  int32_t model, int32_t vref_nb, const double* vref_box);

/* rve_set_batch for n_sys meshes that share ONE topology (triangles, region tags) and differ only in their vertex
 * coordinates xy[n_sys][nv][2] — the meshes deepbnd_b200/mesh/rve_morph.py derives from one template, the reference's
 * datasets vary only the inclusion radii (build_snapshots_param.py:21-41).  The symbolic phase runs once; per system only
 * the coordinates, the level-1 cell of every row and the restriction lists are rebuilt.  The boundary vertices must be
 * the same in every system (checked).  Results equal those of rve_set_batch on the same meshes up to the row order
 * (taken from system 0), i.e. to rounding. */
int rve_set_batch_shared(rve_handle_t h, int32_t n_sys, int32_t nv, int32_t nt, const double* xy, const int32_t* tri,
                         const int32_t* region, const double* mat, int32_t model, int32_t vref_nb, const double* vref_box);

/* rve_set_batch_shared with the coordinates generated ON THE DEVICE by radial morphing of one template mesh
 * (deepbnd_b200/mesh/rve_morph.py; the reference re-meshes every snapshot with gmsh, mesh_RVE.py:38-64): per RVE only
 * its inclusion parameters travel to the GPU.  xy_template[nv][2] is the template mesh; the moving nodes are described
 * by mov_node (vertex id), mov_incl (nearest inclusion), mov_d / mov_dir (distance and unit direction from its centre);
 * centre[n_incl][2], template radius r0 and influence radius R; prm[n_sys][n_incl][3] = (l, e, theta) of every inclusion
 * (circle: e = 1). */
int rve_set_batch_morphed(rve_handle_t h, int32_t n_sys, int32_t nv, int32_t nt, const double* xy_template, const int32_t* tri,
                          const int32_t* region, const double* mat, int32_t model, int32_t vref_nb, const double* vref_box,
                          int32_t n_mov, const int32_t* mov_node, const int32_t* mov_incl, const double* mov_d, const double* mov_dir,
                          int32_t n_incl, const double* centre, double r0, double R, const double* prm);

/* replaces: A = mp.block_assemble(a); bcs.apply(A) (micro_model.py:60-62, micro_model_dnn.py:65-67).
 * Numeric P2 stiffness assembly of every system into the SELL-32 node-block matrix + preconditioner set-up. */
int rve_assemble(rve_handle_t h);

/* parity-test accessors (not on the hot path) */
int rve_get_sizes(rve_handle_t h, int32_t sys, int64_t* n_nodes, int64_t* n_rows, int64_t* n_blocks,
                  int64_t* n_elems);
/* scalar CSR of system sys: 2*n_rows rows, 4*n_blocks entries, sorted columns. */
int rve_get_csr(rve_handle_t h, int32_t sys, int32_t* rowptr, int32_t* col, double* val);
/* for each active block row: the P2 node it stands for as a vertex pair (va==vb: vertex; else
 * the mid-edge node of edge (va,vb), va<vb). */
int rve_get_rowmap(rve_handle_t h, int32_t sys, int32_t* va, int32_t* vb);
/* same for every P2 node (node order = vertices, then edges sorted by (va,vb)). */
int rve_get_nodemap(rve_handle_t h, int32_t sys, int32_t* va, int32_t* vb);
/* assembled right-hand sides of the last rve_solve: rhs[n_rows][n_rhs][2] */
int rve_get_rhs(rve_handle_t h, int32_t sys, double* rhs);
/* full nodal fluctuation field of the last rve_solve: u[n_nodes][n_rhs][2] */
int rve_get_solution(rve_handle_t h, int32_t sys, double* u);
/* level-l coarse operator of the two-level preconditioner as dense (2G x 2G), G returned by
 * rve_get_coarse_dim; test accessor. */
int rve_get_coarse_dim(rve_handle_t h, int32_t level, int32_t* G, int32_t* n_levels);
int rve_get_coarse_dense(rve_handle_t h, int32_t sys, int32_t level, double* dense);

/* replaces: the load loop  Eps.assign; F = mp.block_assemble(f); bcs.apply(F); solver.solve(A,x,F)
 * (micro_model.py:68-82, micro_model_dnn.py:81-101) incl. the dnn pre-step
 * B = -Integral(outer(uD,n), Mref.ds)/4; uD += Bx; Eps = macro_strain(i) - B (micro_model_dnn.py:85-94).
 *   n_rhs      load cases solved together against the same matrix (1..3)
 *   eps        [n_sys][n_rhs][3] Voigt macro strains (before the dnn B correction)
 *   uD         [n_sys][n_rhs][2*nvref] Dirichlet fluctuation on Vref (dnn) or NULL (zero: 'lin')
 *   rtol       stop when ||r||_2 <= rtol ||b||_2 for every load of a system
 *   status     [n_sys] 0 = converged, 1 = maxit reached;   iters[n_sys] iteration counts */
int rve_solve(rve_handle_t h, int32_t n_rhs, const double* eps, const double* uD, double rtol,
              int32_t maxit, int32_t precond, int32_t* status, int32_t* iters);

/* replaces: sigma_hom = Integral(sig_mu, dy(0)+dy(1))/vol; Chom_[:,i] = sigma_hom.flatten()[[0,3,1]]
 * (micro_model_dnn.py:103-106).  Needs n_rhs == 3 with loads 0,1,2.  C[n_sys][3][3]. */
int rve_epilogue_tangent(rve_handle_t h, double* C);

/* replaces solve_snapshot's per-load post-processing (simulation_snapshots.py:50-60):
 * getAffineTransformationLocal (multiscale/misc.py:49-61), usol.interpolate onto Vref,
 * homogenise([0,1]) and homogenise([0,1,2,3]) (micro_model.py:85-96), flatten()[[0,3,2]].
 *   sol[n_sys][n_rhs][2*nvref], sigma[n_sys][n_rhs][3], a[n_sys][n_rhs][2], B[n_sys][n_rhs][2][2],
 *   sigmaT[n_sys][n_rhs][3]; any pointer may be NULL. */
int rve_epilogue_snapshot(rve_handle_t h, double* sol, double* sigma, double* a, double* B,
                          double* sigmaT);

/* replaces the per-load post-processing of MicroConstitutiveModelGen.computeHomogenisation
 * (core/multiscale/micro_model_gen.py:139-147): average stress (Voigt 00,11,01) over the regions {0,1}
 * ("L") and over all regions, average gradient of the fluctuation (2x2, row-major) over the same two
 * region sets, and the effective macro strain eps - B of each load ([e11, e22, 2 e12]).  Each array is
 * [n_sys][n_rhs][.]; any pointer may be NULL. */
int rve_epilogue_homogenisation(rve_handle_t h, double* sigmaL, double* sigmaT, double* gradL, double* gradT,
                                double* eps_eff);

/* solver options.  RVE_OPT_SYM_B: symmetrise the dnn correction B = 0.5 (B + B^T) like
 * micro_model_gen.py:122 (default 0: full B, micro_model_dnn.py:88). */
#define RVE_OPT_SYM_B 1
/* RVE_OPT_OPERATOR: how q = A p is applied inside the PCG.  RVE_OPERATOR_MATFREE (default): element by element
 * from 72 bytes per element, the stiffness matrix is never stored (rve_matfree.cuh);  RVE_OPERATOR_ASSEMBLED: the
 * SELL-32 node-block matrix written by the assembly kernel (34 bytes per 2x2 block) and the SpMV kernel.  Both give
 * the same operator to rounding; rve_get_csr always returns the assembled matrix.  Set it before rve_assemble. */
#define RVE_OPT_OPERATOR 2
#define RVE_OPERATOR_ASSEMBLED 0
#define RVE_OPERATOR_MATFREE 1
/* RVE_OPT_RESIDUAL_REPLACE = N > 0: every N PCG iterations the recurrence residual is replaced by the true residual
 * b - A x (one extra operator application) and its norms are recomputed, so that the stopping test cannot drift away
 * from the true residual on badly conditioned systems (high contrast); 0 (default) = never. */
#define RVE_OPT_RESIDUAL_REPLACE 3
int rve_set_option(rve_handle_t h, int32_t option, int32_t value);

/* replaces getAlphas_fast (creation_model/RB/RB_utils.py:211-213): Y = U M W[:nmax]^T on the
 * Vref samples of the last solve.  W[n_rhs][nmax][nh] (one basis per load), M[nh][nh],
 * translate[n_sys][n_rhs][2] or NULL: constant added to the samples first
 * (build_inout.py:36-51 translateSolution).  Y[n_sys][n_rhs][nmax]. */
int rve_project(rve_handle_t h, int32_t nmax, int32_t nh, const double* W, const double* M,
                const double* translate, double* Y);

/* CUDA-event phase timings of the last calls (ms) and traffic counters:
 *   t[0] H2D + symbolic phase up to the global numbering t[1] pattern / halo / restriction lists t[2] assemble t[3] precond set-up t[4] rhs t[5] pcg total
 *   t[6] spmv kernels total t[7] epilogue;   c[0] spmv launches c[1] algorithmic spmv bytes (B_spmv(k))
 *   c[2] pcg iterations launched c[3] sum of per-system iterations c[4] kernels launched since
 *   rve_create c[5] constraint post-processing ms c[6]/c[7] host->device / device->host bytes since
 *   rve_create; t has 8 entries, c has 8 entries */
int rve_timings(rve_handle_t h, double* t, double* c);

/* Operator micro-benchmark on the assembled batch: runs `reps` launches of the kernel RVE_OPT_OPERATOR selects
 * (k_apply or k_spmv; n_rhs right-hand sides) on a dedicated stream and returns the average launch time in ms and the
 * algorithmic bytes per launch. */
int rve_bench_spmv(rve_handle_t h, int32_t n_rhs, int32_t reps, double* ms_per_launch,
                   double* alg_bytes_per_launch);

/* parity-test accessor (not on the hot path): q = A p for the whole batch with the operator RVE_OPT_OPERATOR
 * selects; p, q: [total rows][n_rhs][2], rows of system s start at the sum of n_rows of the systems before it. */
int rve_apply_operator(rve_handle_t h, int32_t n_rhs, const double* p, double* q);

/* Host-only test hook (no CUDA call, not used by the product): CPU restatement of the symbolic phase of ONE system —
 * P2 numbering, constraint map, rows ordered by coarse cell, node-block pattern, Vref sample location.  The CPU
 * test-suite checks it against the oracle's pattern, the GPU tests check the device numbering against it.  Arrays are caller-allocated with the
 * given capacities (node_row: nv + 3*nt entries is always enough). */
int rve_host_symbolic(const double* xy, int32_t nv, const int32_t* tri, const int32_t* region, int32_t nt,
                      int32_t model, int32_t vref_nb, int32_t cap_rows, int32_t cap_blocks, int32_t* n_nodes,
                      int32_t* n_rows, int32_t* n_blocks, int32_t* nc, int32_t* rowptr, int32_t* bcol,
                      int32_t* row_va, int32_t* row_vb, int32_t* node_row, int32_t* samp_elem, double* samp_lam);

#ifdef __cplusplus
}
#endif
#endif
