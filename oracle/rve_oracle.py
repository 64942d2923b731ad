"""CPU oracle for the DeepBND RVE hot path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy/scipy restatement of the discrete problem that the
reference (felipefr/deepbnd) solves with FEniCS 2019.1 + multiphenics + PETSc LU.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may
import it; the product path (`deepbnd_b200/`) never does.

PARITY UNPINNED: the arithmetic of the reference lives in dolfin / multiphenics /
micmacsfenics, none of which is under /root/reference nor installable here, and
the reference's own tests hold no golden vector for the solve path (SURVEY.md
F2/F3/F6).  The oracle is therefore pinned only by (i) analytic known answers
(homogeneous material, patch tests, energy ordering), (ii) the reference's
material constants and tag conventions, (iii) `tests/test_sampling.py` goldens
(off-path), (iv) the reference's golden mesh volume fractions (tests/test_meshes.py:58-59), which the
mesher reproduces to 1e-15.  tools/export_goldens_fenics.py + tests/test_fenics_goldens.py turn this
into a pinned oracle on a machine that has FEniCS (the tests are skipped while the files are absent).

Reference call sites restated here (file:line under /root/reference):
  material table          core/elasticity/fenics_utils.py:41-53, core/elasticity/misc.py:25-33
  per-cell coefficient    core/fenics_tools/wrapper_expression.py:15-56
  stress law              core/multiscale/micro_model_dnn.py:53, micro_model.py:49
  K once, 3 RHS, LU       core/multiscale/micro_model.py:52-82
  homogenise              core/multiscale/micro_model.py:85-96
  tangent + dnn pre-step  core/multiscale/micro_model_dnn.py:50-113
  (a, B) affine data      core/multiscale/misc.py:49-61
  Integral                core/fenics_tools/misc.py:18-41
  snapshot outputs        creation_model/dataset/simulation_snapshots.py:42-64
  RB projection, mass     creation_model/RB/RB_utils.py:125-132,211-213
[EXT] micmacsfenics formulations: 'lin'/'dnn' = Dirichlet data on
the whole boundary (no zero-average block, switchable), 'per' = periodic DOF
identification + zero-average multiplier, 'MR' = zero-average multiplier (2) +
tensor multiplier on int u (x) n ds (4).

Linear solver: scipy.sparse.linalg.splu (sparse LU, like the reference's
PETScLUSolver), factor once, solve every RHS.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

# ----------------------------------------------------------------------------------------------
# material (core/elasticity/misc.py:25-33, fenics_utils.py:41-53)
# ----------------------------------------------------------------------------------------------

def eng2mu(nu, E):
    return E / (2.0 * (1.0 + nu))


def eng2lamb(nu, E):
    return nu * E / ((1.0 - 2.0 * nu) * (1.0 + nu))


def lame2lambPlane(lamb, mu):
    return (2.0 * mu * lamb) / (lamb + 2.0 * mu)


def eng2lambPlane(nu, E):
    return lame2lambPlane(eng2lamb(nu, E), eng2mu(nu, E))


def lame_table(nu1, E1, nu2, E2):
    """4x2 table [lambda*, mu] indexed by (region - min(region)); fenics_utils.py:42-51."""
    l1, m1 = eng2lambPlane(nu1, E1), eng2mu(nu1, E1)
    l2, m2 = eng2lambPlane(nu2, E2), eng2mu(nu2, E2)
    return np.array([[l1, m1], [l2, m2], [l1, m1], [l2, m2]])


def macro_strain(i):
    """[EXT] micmacsfenics macro_strain: unit Voigt strain i with engineering shear."""
    return [np.array([[1.0, 0.0], [0.0, 0.0]]),
            np.array([[0.0, 0.0], [0.0, 1.0]]),
            np.array([[0.0, 0.5], [0.5, 0.0]])][i]


# ----------------------------------------------------------------------------------------------
# P2 mesh topology
# ----------------------------------------------------------------------------------------------

class P2Mesh:
    """Straight-edged P2 triangle mesh.  Local node order: v0 v1 v2 m01 m12 m20."""

    def __init__(self, xy, tri, region):
        xy = np.asarray(xy, dtype=np.float64)
        tri = np.asarray(tri, dtype=np.int64)
        # force counter-clockwise orientation
        a = xy[tri[:, 1]] - xy[tri[:, 0]]
        b = xy[tri[:, 2]] - xy[tri[:, 0]]
        neg = (a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]) < 0
        tri = tri.copy()
        tri[neg, 1], tri[neg, 2] = tri[neg, 2].copy(), tri[neg, 1].copy()
        self.nv = xy.shape[0]
        self.nt = tri.shape[0]
        self.tri = tri
        self.region = np.asarray(region, dtype=np.int64) - int(np.min(region))
        e = np.stack([tri[:, [0, 1]], tri[:, [1, 2]], tri[:, [2, 0]]], axis=1)  # nt,3,2
        es = np.sort(e.reshape(-1, 2), axis=1)
        key = es[:, 0] * self.nv + es[:, 1]
        ukey, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
        self.ne = ukey.shape[0]
        self.edges = np.stack([ukey // self.nv, ukey % self.nv], axis=1)
        self.edge_count = cnt
        mid = 0.5 * (xy[self.edges[:, 0]] + xy[self.edges[:, 1]])
        self.xy = np.vstack([xy, mid])
        self.nn = self.nv + self.ne
        self.cells = np.hstack([tri, self.nv + inv.reshape(-1, 3)])  # nt,6
        bedge = np.where(cnt == 1)[0]
        self.bnd_edges = bedge
        bn = np.zeros(self.nn, dtype=bool)
        bn[self.edges[bedge].ravel()] = True
        bn[self.nv + bedge] = True
        self.bnd_nodes = np.where(bn)[0]
        self.x0, self.y0 = self.xy.min(axis=0)
        self.x1, self.y1 = self.xy.max(axis=0)
        # geometry
        p0, p1, p2 = xy[tri[:, 0]], xy[tri[:, 1]], xy[tri[:, 2]]
        self.area = 0.5 * ((p1[:, 0] - p0[:, 0]) * (p2[:, 1] - p0[:, 1])
                           - (p1[:, 1] - p0[:, 1]) * (p2[:, 0] - p0[:, 0]))
        twoA = 2.0 * self.area
        gl = np.empty((self.nt, 3, 2))
        gl[:, 0, 0] = (p1[:, 1] - p2[:, 1]) / twoA
        gl[:, 0, 1] = (p2[:, 0] - p1[:, 0]) / twoA
        gl[:, 1, 0] = (p2[:, 1] - p0[:, 1]) / twoA
        gl[:, 1, 1] = (p0[:, 0] - p2[:, 0]) / twoA
        gl[:, 2, 0] = (p0[:, 1] - p1[:, 1]) / twoA
        gl[:, 2, 1] = (p1[:, 0] - p0[:, 0]) / twoA
        self.gradlam = gl


_QL = np.array([[0.5, 0.5, 0.0], [0.0, 0.5, 0.5], [0.5, 0.0, 0.5]])  # edge-midpoint rule, exact deg 2


def p2_shape(lam):
    """P2 shape functions at barycentric points lam (...,3) -> (...,6)."""
    l0, l1, l2 = lam[..., 0], lam[..., 1], lam[..., 2]
    return np.stack([l0 * (2 * l0 - 1), l1 * (2 * l1 - 1), l2 * (2 * l2 - 1),
                     4 * l0 * l1, 4 * l1 * l2, 4 * l2 * l0], axis=-1)


def p2_grad(mesh, lam):
    """Gradients of the 6 shape functions at one barycentric point: (nt,6,2)."""
    gl = mesh.gradlam
    l0, l1, l2 = lam
    g = np.empty((mesh.nt, 6, 2))
    g[:, 0] = (4 * l0 - 1) * gl[:, 0]
    g[:, 1] = (4 * l1 - 1) * gl[:, 1]
    g[:, 2] = (4 * l2 - 1) * gl[:, 2]
    g[:, 3] = 4 * (l1 * gl[:, 0] + l0 * gl[:, 1])
    g[:, 4] = 4 * (l2 * gl[:, 1] + l1 * gl[:, 2])
    g[:, 5] = 4 * (l0 * gl[:, 2] + l2 * gl[:, 0])
    return g


def voigt_D(lam, mu):
    D = np.zeros(lam.shape + (3, 3))
    D[..., 0, 0] = D[..., 1, 1] = lam + 2 * mu
    D[..., 0, 1] = D[..., 1, 0] = lam
    D[..., 2, 2] = mu
    return D


def element_B(mesh, lam):
    """Voigt strain-displacement matrix (nt,3,12), dof order 2*a+c."""
    g = p2_grad(mesh, lam)
    B = np.zeros((mesh.nt, 3, 12))
    B[:, 0, 0::2] = g[:, :, 0]
    B[:, 1, 1::2] = g[:, :, 1]
    B[:, 2, 0::2] = g[:, :, 1]
    B[:, 2, 1::2] = g[:, :, 0]
    return B


def assemble(mesh, mat):
    """Global stiffness (2nn x 2nn CSR) and the per-element data reused by the loads."""
    D = voigt_D(mat[mesh.region, 0], mat[mesh.region, 1])          # nt,3,3
    Ke = np.zeros((mesh.nt, 12, 12))
    BtD_int = np.zeros((mesh.nt, 12, 3))                           # int B^T D
    Bint = np.zeros((mesh.nt, 3, 12))                              # int B
    for q in range(3):
        B = element_B(mesh, _QL[q])
        w = (mesh.area / 3.0)[:, None, None]
        DB = np.einsum('tij,tjk->tik', D, B)
        Ke += w * np.einsum('tji,tjk->tik', B, DB)
        BtD_int += w * np.einsum('tji,tjk->tik', B, D)
        Bint += w * B
    dofs = np.stack([2 * mesh.cells, 2 * mesh.cells + 1], axis=2).reshape(mesh.nt, 12)
    I = np.repeat(dofs, 12, axis=1).ravel()
    J = np.tile(dofs, (1, 12)).ravel()
    K = sp.coo_matrix((Ke.ravel(), (I, J)), shape=(2 * mesh.nn, 2 * mesh.nn)).tocsr()
    K.sum_duplicates()
    K.sort_indices()
    return K, dict(D=D, BtD_int=BtD_int, Bint=Bint, dofs=dofs, Ke=Ke)


def load_vector(mesh, ed, eps_voigt):
    """F = - int B^T D eps  (micro_model.py:72-74 with [EXT] f = -int Eps : sigma(v))."""
    fe = -np.einsum('tik,k->ti', ed['BtD_int'], eps_voigt)
    F = np.zeros(2 * mesh.nn)
    np.add.at(F, ed['dofs'].ravel(), fe.ravel())
    return F


def strain2voigt(e):
    return np.array([e[0, 0], e[1, 1], e[0, 1] + e[1, 0]])


# ----------------------------------------------------------------------------------------------
# constraint operators
# ----------------------------------------------------------------------------------------------

def mean_constraint(mesh):
    """C (2 x 2nn): int u_c dx.  P2: vertex functions integrate to 0, mid-edge ones to A/3."""
    w = np.zeros(mesh.nn)
    np.add.at(w, mesh.cells[:, 3:].ravel(), np.repeat(mesh.area / 3.0, 3))
    C = np.zeros((2, 2 * mesh.nn))
    C[0, 0::2] = w
    C[1, 1::2] = w
    return C


def boundary_edges_oriented(mesh):
    """Boundary edges as (va, vb, mid, length, outward normal) using the owning CCW triangle."""
    out = []
    bset = set(mesh.bnd_edges.tolist())
    for t in range(mesh.nt):
        for k, (i, j) in enumerate(((0, 1), (1, 2), (2, 0))):
            m = mesh.cells[t, 3 + k]
            if (m - mesh.nv) in bset:
                a, b = mesh.cells[t, i], mesh.cells[t, j]
                d = mesh.xy[b] - mesh.xy[a]
                L = np.hypot(*d)
                n = np.array([d[1], -d[0]]) / L        # CCW triangle -> outward is to the right
                out.append((a, b, m, L, n))
    return out


def outer_constraint(mesh):
    """C (4 x 2nn): int u_i n_j ds, row 2*i+j.  Simpson on P2 edge: (L/6)(u_a + 4 u_m + u_b)."""
    C = np.zeros((4, 2 * mesh.nn))
    for a, b, m, L, n in boundary_edges_oriented(mesh):
        for i in range(2):
            for j in range(2):
                C[2 * i + j, 2 * a + i] += L / 6.0 * n[j]
                C[2 * i + j, 2 * b + i] += L / 6.0 * n[j]
                C[2 * i + j, 2 * m + i] += 4.0 * L / 6.0 * n[j]
    return C


def periodic_map(mesh, tol=1e-9):
    """node -> master node for the periodic identification (x1->x0, y1->y0, corners -> (x0,y0))."""
    xy = mesh.xy
    master = np.arange(mesh.nn)
    bn = mesh.bnd_nodes
    L = xy[bn]
    on_l = np.abs(L[:, 0] - mesh.x0) < tol
    on_r = np.abs(L[:, 0] - mesh.x1) < tol
    on_b = np.abs(L[:, 1] - mesh.y0) < tol
    on_t = np.abs(L[:, 1] - mesh.y1) < tol
    left = bn[on_l]
    bottom = bn[on_b]
    for s in bn[on_r]:
        d = np.abs(xy[left, 1] - xy[s, 1])
        k = np.argmin(d)
        assert d[k] < 1e-7, "periodic mesh mismatch (left/right)"
        master[s] = left[k]
    for s in bn[on_t]:
        d = np.abs(xy[bottom, 0] - xy[s, 0])
        k = np.argmin(d)
        assert d[k] < 1e-7, "periodic mesh mismatch (bottom/top)"
        master[s] = bottom[k] if master[s] == s else master[s]
    # resolve chains (top-right corner -> left-top -> ... -> bottom-left)
    for _ in range(3):
        master = master[master]
    for s in bn[on_t]:
        # top nodes that were first mapped to the left side (corners) still need y-wrap
        m = master[s]
        if abs(xy[m, 1] - mesh.y1) < tol:
            d = np.abs(xy[bottom, 0] - xy[m, 0])
            master[s] = master[bottom[np.argmin(d)]]
    for _ in range(3):
        master = master[master]
    return master


# ----------------------------------------------------------------------------------------------
# Vref: the boundary reference space (core/mesh/degenerated_rectangle_mesh.py:3-35)
# ----------------------------------------------------------------------------------------------

def vref_points(x0=-1.0, y0=-1.0, Lx=2.0, Ly=2.0, Nb=21):
    """Our fixed Vref node order: 4*(Nb-1) boundary nodes CCW from (x0,y0), then the centre.
    DOF = 2*node + comp.  (dolfin's own ordering is not recorded anywhere)"""
    n = Nb - 1
    s = np.arange(n) / n
    pts = np.vstack([np.stack([x0 + Lx * s, np.full(n, y0)], 1),
                     np.stack([np.full(n, x0 + Lx), y0 + Ly * s], 1),
                     np.stack([x0 + Lx * (1 - s), np.full(n, y0 + Ly)], 1),
                     np.stack([np.full(n, x0), y0 + Ly * (1 - s)], 1),
                     [[x0 + 0.5 * Lx, y0 + 0.5 * Ly]]])
    return pts


def vref_mass(pts):
    """Boundary (ds) mass matrix of the P1 vector space on Vref (RB_utils.py:125-132 with
    dotProduct = inner(u,v)*ds, build_inout.py:34): cyclic tridiagonal, centre rows zero."""
    nb = pts.shape[0] - 1
    M = np.zeros((2 * (nb + 1), 2 * (nb + 1)))
    for k in range(nb):
        k2 = (k + 1) % nb
        L = np.hypot(*(pts[k2] - pts[k]))
        for c in range(2):
            M[2 * k + c, 2 * k + c] += L / 3
            M[2 * k2 + c, 2 * k2 + c] += L / 3
            M[2 * k + c, 2 * k2 + c] += L / 6
            M[2 * k2 + c, 2 * k + c] += L / 6
    return M


def vref_B_correction(uD, pts):
    """B = -(1/vol) oint uD (x) n ds on the Vref boundary (micro_model_dnn.py:88), exact
    trapezoid for P1; vol = Lx*Ly (volMref = 4.0 at micro_model_dnn.py:77)."""
    nb = pts.shape[0] - 1
    u = uD.reshape(-1, 2)
    B = np.zeros((2, 2))
    for k in range(nb):
        k2 = (k + 1) % nb
        d = pts[k2] - pts[k]
        L = np.hypot(*d)
        n = np.array([d[1], -d[0]]) / L
        B += L * 0.5 * np.outer(u[k] + u[k2], n)
    vol = (pts[:-1, 0].max() - pts[:-1, 0].min()) * (pts[:-1, 1].max() - pts[:-1, 1].min())
    return -B / vol


def vref_eval_on_boundary(uD, pts, xq):
    """Evaluate the P1 boundary trace of uD (Vref order) at points xq lying on the box boundary."""
    nb = pts.shape[0] - 1
    n = nb // 4
    x0, y0 = pts[0]
    Lx = pts[n, 0] - x0
    Ly = pts[2 * n, 1] - y0
    u = uD.reshape(-1, 2)
    out = np.zeros((xq.shape[0], 2))
    for q, (x, y) in enumerate(xq):
        tol = 1e-9
        if abs(y - y0) < tol and x < x0 + Lx - tol:
            s = (x - x0) / Lx * n
        elif abs(x - (x0 + Lx)) < tol and y < y0 + Ly - tol:
            s = n + (y - y0) / Ly * n
        elif abs(y - (y0 + Ly)) < tol and x > x0 + tol:
            s = 2 * n + (x0 + Lx - x) / Lx * n
        else:
            s = 3 * n + (y0 + Ly - y) / Ly * n
        k = int(np.floor(s + 1e-12))
        t = s - k
        k = k % nb
        out[q] = (1 - t) * u[k] + t * u[(k + 1) % nb]
    return out


# ----------------------------------------------------------------------------------------------
# The micro model
# ----------------------------------------------------------------------------------------------

class OracleRVE:
    """One RVE = one object, like MicroModel / MicroConstitutiveModelDNN."""

    def __init__(self, xy, tri, region, param=(0.3, 10.0, 0.3, 1.0), model='lin',
                 zero_average_for_dirichlet=False, dirichlet_semantics='p1_vertex'):
        # the two [EXT] switches a FEniCS export settles (tools/export_goldens_fenics.py): the zero-average block
        # for lin/dnn, and how the foreign-mesh P1 function uD becomes Dirichlet data on the P2 boundary nodes
        # ('p1_vertex': vertex values, mid-edge = mean of the end vertices, dolfin's restriction of a P1 coefficient
        # to the RVE cell; 'p2_nodal': uD evaluated at every P2 boundary node)
        assert dirichlet_semantics in ('p1_vertex', 'p2_nodal')
        self.dir_sem = dirichlet_semantics
        self.mesh = P2Mesh(xy, tri, region)
        self.mat = lame_table(*param)
        self.model = model
        self.zero_avg_dir = zero_average_for_dirichlet
        self.K, self.ed = assemble(self.mesh, self.mat)
        self.sol = [None, None, None]
        self.eps = [None, None, None]
        self._lu = None

    # ---- system construction per model --------------------------------------------------------
    def _factor(self):
        m = self.mesh
        N = 2 * m.nn
        if self.model in ('lin', 'dnn'):
            bd = np.zeros(N, dtype=bool)
            bd[2 * m.bnd_nodes] = True
            bd[2 * m.bnd_nodes + 1] = True
            self.free = np.where(~bd)[0]
            self.bdofs = np.where(bd)[0]
            Kff = self.K[self.free][:, self.free].tocsc()
            if self.zero_avg_dir:
                C = sp.csr_matrix(mean_constraint(m)[:, self.free])
                Kff = sp.bmat([[Kff, C.T], [C, None]]).tocsc()
            self._lu = spla.splu(Kff)
        elif self.model == 'per':
            master = periodic_map(m)
            um = np.unique(master)
            idx = -np.ones(m.nn, dtype=np.int64)
            idx[um] = np.arange(um.size)
            col = idx[master]
            P = sp.coo_matrix((np.ones(N), (np.arange(N), np.stack([2 * col, 2 * col + 1], 1).ravel())),
                              shape=(N, 2 * um.size)).tocsr()
            self.P = P
            Kp = (P.T @ self.K @ P)
            C = sp.csr_matrix(mean_constraint(m) @ P)
            self._lu = spla.splu(sp.bmat([[Kp, C.T], [C, None]]).tocsc())
        elif self.model == 'MR':
            C = sp.csr_matrix(np.vstack([mean_constraint(m), -outer_constraint(m)]))
            self.C = C
            self._lu = spla.splu(sp.bmat([[self.K, C.T], [C, None]]).tocsc())
        else:
            raise ValueError(self.model)

    def solve_load(self, i, eps=None, g_bnd=None):
        """Solve load case i.  eps: 2x2 macro strain (default macro_strain(i));
        g_bnd: (n_bnd_nodes,2) Dirichlet data at mesh.bnd_nodes for lin/dnn."""
        if self._lu is None:
            self._factor()
        m = self.mesh
        N = 2 * m.nn
        eps = macro_strain(i) if eps is None else eps
        self.eps[i] = eps
        F = load_vector(m, self.ed, strain2voigt(eps))
        u = np.zeros(N)
        if self.model in ('lin', 'dnn'):
            if g_bnd is not None:
                u[2 * m.bnd_nodes] = g_bnd[:, 0]
                u[2 * m.bnd_nodes + 1] = g_bnd[:, 1]
            rhs = (F - self.K @ u)[self.free]
            if self.zero_avg_dir:
                Cfull = mean_constraint(m)
                rhs = np.concatenate([rhs, -Cfull @ u])
            x = self._lu.solve(rhs)
            u[self.free] = x[:self.free.size]
        elif self.model == 'per':
            rhs = np.concatenate([self.P.T @ F, np.zeros(2)])
            x = self._lu.solve(rhs)
            u = self.P @ x[:-2]
        elif self.model == 'MR':
            rhs = np.concatenate([F, np.zeros(6)])
            x = self._lu.solve(rhs)
            u = x[:N]
        self.sol[i] = u
        return u

    def compute(self):
        """MicroModel.compute (micro_model.py:52-82): three unit macro strains."""
        for i in range(3):
            self.solve_load(i)

    # ---- dnn / lin Dirichlet data (micro_model_dnn.py:76-94) ----------------------------------
    def dirichlet_from_vref(self, uD, pts):
        """B-corrected Dirichlet data at the RVE boundary nodes + corrected macro strain shift.
        Returns (g_bnd, B).  Vertex values = P1 trace of uD at the vertex, mid-edge values =
        mean of the edge's end vertices ([EXT] dolfin restriction semantics)."""
        m = self.mesh
        B = vref_B_correction(uD, pts)
        uDc = uD.reshape(-1, 2) + pts @ B.T               # uD += B x   (micro_model_dnn.py:89-91)
        g = np.zeros((m.nn, 2))
        isb = np.zeros(m.nn, dtype=bool)
        isb[m.bnd_nodes] = True
        bv = m.bnd_nodes[m.bnd_nodes < m.nv]
        g[bv] = vref_eval_on_boundary(uDc.ravel(), pts, m.xy[bv])
        be = m.bnd_edges
        if self.dir_sem == 'p1_vertex':
            g[m.nv + be] = 0.5 * (g[m.edges[be, 0]] + g[m.edges[be, 1]])
        else:
            g[m.nv + be] = vref_eval_on_boundary(uDc.ravel(), pts, m.xy[m.nv + be])
        return g[m.bnd_nodes], B

    def compute_tangent(self, uD_list=None, pts=None):
        """MicroConstitutiveModelDNN.computeTangent (micro_model_dnn.py:50-113)."""
        C = np.zeros((3, 3))
        for i in range(3):
            B = np.zeros((2, 2))
            g = None
            if self.model in ('lin', 'dnn') and uD_list is not None:
                g, B = self.dirichlet_from_vref(np.asarray(uD_list[i], dtype=float), pts)
            self.solve_load(i, eps=macro_strain(i) - B, g_bnd=g)
            C[:, i] = self.homogenise([0, 1], i).flatten()[[0, 3, 1]]
        return C

    # ---- post-processing ------------------------------------------------------------------------
    def _elem_grad_int(self, u):
        """int_T grad(u) dx per element: (nt,2,2) [i,j] = int d u_i / d x_j."""
        m = self.mesh
        G = np.zeros((m.nt, 2, 2))
        ue = u[self.ed['dofs']].reshape(m.nt, 6, 2)
        for q in range(3):
            g = p2_grad(m, _QL[q])
            G += (m.area / 3.0)[:, None, None] * np.einsum('tai,taj->tij', ue, g)
        return G

    def homogenise(self, regions, i):
        """Region-averaged stress of (Eps.y + u~) (micro_model.py:85-96); returns 2x2."""
        m = self.mesh
        sel = np.isin(m.region, regions)
        vol = m.area[sel].sum()
        G = self._elem_grad_int(self.sol[i])
        lam = self.mat[m.region, 0]
        mu = self.mat[m.region, 1]
        H = G + m.area[:, None, None] * self.eps[i][None]
        tr = H[:, 0, 0] + H[:, 1, 1]
        S = lam[:, None, None] * tr[:, None, None] * np.eye(2)[None] + mu[:, None, None] * (H + H.transpose(0, 2, 1))
        return S[sel].sum(axis=0) / vol

    def affine_local(self, i, regions=(0, 1)):
        """getAffineTransformationLocal (multiscale/misc.py:49-61): a = -u_avg + G_avg yG, B = -G_avg."""
        m = self.mesh
        sel = np.isin(m.region, regions)
        vol = m.area[sel].sum()
        cent = m.xy[m.tri].mean(axis=1)
        yG = (cent[sel] * m.area[sel, None]).sum(axis=0) / vol
        u = self.sol[i]
        ue = u[self.ed['dofs']].reshape(m.nt, 6, 2)
        uint = (m.area / 3.0)[:, None] * ue[:, 3:, :].sum(axis=1)
        uavg = uint[sel].sum(axis=0) / vol
        Gavg = self._elem_grad_int(u)[sel].sum(axis=0) / vol
        return -uavg + Gavg @ yG, -Gavg

    def compute_homogenisation(self, uD_list=None, pts=None, domainL=(0, 1)):
        """MicroConstitutiveModelGen.computeHomogenisation (micro_model_gen.py:63-154): sigma, eps over the RVE
        and over domainL (Voigt, elasticity/misc.py:58-62), tangent = sigma eps^-1.  dnn: B symmetrised (:122)."""
        m = self.mesh
        Hom = {k: np.zeros((3, 3)) for k in ("sigma", "sigmaL", "eps", "epsL", "tangent", "tangentL")}
        for i in range(3):
            B = np.zeros((2, 2))
            g = None
            if self.model == 'dnn' and uD_list is not None:
                uD = np.asarray(uD_list[i], dtype=float)
                B = vref_B_correction(uD, pts)
                B = 0.5 * (B + B.T)
                uDc = uD.reshape(-1, 2) + pts @ B.T
                gn = np.zeros((m.nn, 2))
                bv = m.bnd_nodes[m.bnd_nodes < m.nv]
                gn[bv] = vref_eval_on_boundary(uDc.ravel(), pts, m.xy[bv])
                be = m.bnd_edges
                gn[m.nv + be] = 0.5 * (gn[m.edges[be, 0]] + gn[m.edges[be, 1]])
                g = gn[m.bnd_nodes]
            eps = macro_strain(i) - B
            self.solve_load(i, eps=eps, g_bnd=g)
            G = self._elem_grad_int(self.sol[i])
            for key, regions in (("L", list(domainL)), ("", [0, 1, 2, 3])):
                sel = np.isin(m.region, regions)
                vol = m.area[sel].sum()
                s = self.homogenise(regions, i)
                e = eps + 0.5 * (G[sel].sum(0) + G[sel].sum(0).T) / vol
                Hom["sigma" + key][:, i] = [s[0, 0], s[1, 1], s[0, 1]]
                Hom["eps" + key][:, i] = [e[0, 0], e[1, 1], 2 * e[0, 1]]
        Hom["tangent"] = Hom["sigma"] @ np.linalg.inv(Hom["eps"])
        Hom["tangentL"] = Hom["sigmaL"] @ np.linalg.inv(Hom["epsL"])
        return Hom

    def eval_at(self, u, pts):
        """P2 evaluation of u at arbitrary points (usol.interpolate, simulation_snapshots.py:56)."""
        m = self.mesh
        out = np.zeros((pts.shape[0], 2))
        p0 = m.xy[m.tri[:, 0]]
        gl = m.gradlam
        for q, p in enumerate(pts):
            d = p[None] - p0
            l1 = gl[:, 1, 0] * d[:, 0] + gl[:, 1, 1] * d[:, 1]
            l2 = gl[:, 2, 0] * d[:, 0] + gl[:, 2, 1] * d[:, 1]
            l0 = 1 - l1 - l2
            t = np.argmax(np.minimum(np.minimum(l0, l1), l2))
            lam = np.array([l0[t], l1[t], l2[t]])
            assert lam.min() > -1e-9, "point outside mesh"
            N = p2_shape(lam)
            out[q] = N @ u[self.ed['dofs'][t]].reshape(6, 2)
        return out

    def snapshot(self, pts):
        """solve_snapshot (simulation_snapshots.py:42-64): dict of per-load outputs for 'S' (2), 'A' (0)."""
        out = {}
        for j, lab in ((2, 'S'), (0, 'A')):
            if self.sol[j] is None:
                self.solve_load(j)
            a, B = self.affine_local(j, (0, 1))
            out['a_' + lab] = a
            out['B_' + lab] = B
            out['solutions_' + lab] = self.eval_at(self.sol[j], pts).ravel()
            out['sigma_' + lab] = self.homogenise([0, 1], j).flatten()[[0, 3, 2]]
            out['sigmaTotal_' + lab] = self.homogenise([0, 1, 2, 3], j).flatten()[[0, 3, 2]]
        return out


def rb_project(U, W, M, Nmax=None):
    """getAlphas_fast (RB_utils.py:211-213): Y = U M W[:N]^T."""
    Nmax = W.shape[0] if Nmax is None else Nmax
    return U @ M @ W[:Nmax].T
