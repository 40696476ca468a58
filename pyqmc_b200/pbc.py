"""Periodic boundary conditions: duck-typed pyscf ``Cell`` stand-ins and the host-side tables the
device needs (lattice images and cutoffs of the orbital evaluator, minimal-image shifts, Ewald
reciprocal points).

Reference statements (relative to /root/reference):
  * ``enforce_pbc``                       pyqmc/pbc/pbc.py:17-49 (``coord.wrap_into_cell``)
  * supercell construction / k-points    pyqmc/pbc/supercell.py:18-75
  * twist -> primitive k-point indices    pyqmc/pbc/twists.py:34-65
  * image list, per-shell cutoffs, phases pyqmc/wf/numba/pbcgto.py:518-621 (``max_Ls``,
    ``PeriodicAtomicOrbitalEvaluator.__init__``), ``_estimate_rcut`` 672-695
  * minimal-image shift table             pyqmc/configurations/distance.py:83-121
  * Ewald set-up                          pyqmc/observables/ewald.py:93-200, 356-379

pyscf is a third-party dependency of the reference (``pyscf>=2.8.0,<3.0.0``, pyproject.toml:12) that
is neither vendored under /root/reference nor installable here.  Two of its functions feed TABLES
into this path -- ``Cell.get_lattice_Ls`` (the candidate list of lattice images) and
``pyscf.pbc.gto.cell.estimate_rcut`` (the starting radius of the cutoff estimate).  They are
restated below from their published algorithm (``Cell.get_lattice_Ls`` / ``_estimate_rcut`` of
pyscf/pbc/gto/cell.py, 2.8 series); PARITY UNPINNED at that boundary.  Everything the reference
itself computes from those tables (sorting, ``max_Ls`` cutoffs, phases) is pinned against the
reference run in-container with the same stand-ins (tests/golden/make_golden.py).
"""
import numpy as np

from . import basis as _basis
from .systems import Mol


from .coord import wrap_into_cell as enforce_pbc  # noqa: E402,F401  (pbc.py:17-49; kept under the reference's name)


class Cell(Mol):
    """Minimal pyscf-``Cell`` look-alike (Bohr units): a ``Mol`` with lattice vectors ``a``."""

    dimension = 3

    def __init__(self, atoms, basis, ecp, nelec, charges, a):
        super().__init__(atoms, basis, ecp, nelec, charges)
        self.a = np.asarray(a, dtype=float).tolist()
        self._shells = None

    def lattice_vectors(self):
        return np.asarray(self.a, dtype=float)

    def reciprocal_vectors(self):
        return 2 * np.pi * np.linalg.inv(self.lattice_vectors()).T

    @property
    def vol(self):
        return abs(float(np.linalg.det(self.lattice_vectors())))

    # --- shell accessors used by the rcut estimate (orbitals.py:262-278) ---
    def _shell_list(self):
        if self._shells is None:
            out = []
            for i in range(len(self._atom)):
                for shell in self._basis[self.atom_pure_symbol(i)]:
                    prim = np.asarray(shell[1:], dtype=float)
                    out.append((int(shell[0]), prim[:, 0], prim[:, 1]))
            self._shells = out
        return self._shells

    @property
    def nbas(self):
        return len(self._shell_list())

    def bas_angular(self, ib):
        return self._shell_list()[ib][0]

    def bas_exp(self, ib):
        return self._shell_list()[ib][1]

    def _libcint_ctr_coeff(self, ib):
        l, exps, coefs = self._shell_list()[ib]
        return _basis.normalized_coefficients(l, exps, coefs)[:, None]

    def get_lattice_Ls(self, rcut, dimension=3):
        """Candidate lattice images: every lattice vector with ``|L| < rcut + d_max`` (d_max = the
        largest inter-atomic distance in the cell).  Restates the published pyscf behaviour
        (``discard=True``); the consumers sort by norm and trim by their own cutoffs
        (pbcgto.py:603-614), so any superset of the images inside those cutoffs gives the same
        orbital values."""
        a = self.lattice_vectors()
        r = self.atom_coords()
        dmax = float(np.linalg.norm(r[:, None] - r[None], axis=2).max()) if len(r) > 1 else 0.0
        heights = 1.0 / np.linalg.norm(np.linalg.inv(a).T, axis=1)  # plane spacings
        bounds = np.ceil((rcut + dmax) / heights).astype(int) + 1
        grids = [np.arange(-b, b + 1) for b in bounds]
        Ts = np.stack(np.meshgrid(*grids, indexing="ij"), axis=-1).reshape(-1, 3)
        Ls = Ts @ a
        keep = np.linalg.norm(Ls, axis=1) < rcut + dmax
        return np.ascontiguousarray(Ls[keep])


def estimate_rcut(cell, precision):
    """Stand-in for ``pyscf.pbc.gto.cell.estimate_rcut`` (most diffuse primitive of each shell,
    two fixed-point iterations of  c^2 (2l+1) alpha r^(2l+2) exp(-alpha r^2 / 2) < precision)."""
    rmax = 0.01
    for shell in range(cell.nbas):
        ang = cell.bas_angular(shell)
        exps = cell.bas_exp(shell)
        i = int(np.argmin(exps))
        alpha = float(exps[i])
        c = float(abs(cell._libcint_ctr_coeff(shell)[i]).max())
        C = c * c * (2 * ang + 1) * alpha / precision
        r0 = 20.0
        for _pass in (0, 1):
            r0 = np.sqrt(2.0 * np.log(C * (r0 * r0 * alpha) ** (ang + 1) + 1.0) / alpha)
        rmax = max(rmax, float(r0))
    return rmax


def shell_rcut(cell, eval_gto_precision):
    """``_estimate_rcut`` of the reference (orbitals.py:258-278 == pbcgto.py:672-695): per shell, two fixed-point
    passes of the radius at which the shell has decayed to the requested precision (same arithmetic, term by term:
    these numbers select the lattice images of every periodic orbital evaluation)."""
    start = estimate_rcut(cell, eval_gto_precision)
    precision = eval_gto_precision / max(cell.vol, 1)
    radii = np.empty(cell.nbas)
    for shell in range(cell.nbas):
        ang, exps = cell.bas_angular(shell), cell.bas_exp(shell)
        largest = abs(cell._libcint_ctr_coeff(shell)).max(axis=1)
        fac = 2 * np.pi / cell.vol * largest * ((2 * ang + 1) / (4 * np.pi)) ** 0.5 / exps / precision
        r = start
        for _pass in (0, 1):
            r = (np.log(fac * r ** (ang + 1) + 1.0) / exps) ** 0.5
        radii[shell] = r.max()
    return radii


def _integer_points_in_cell(S):
    """Integer row vectors n with ``n S^-1`` in [0, 1)^3, i.e. the lattice points of the small lattice inside
    one cell of the lattice spanned by the rows of the integer matrix S -- in lexicographic order of n, as
    the reference's meshgrid enumeration (supercell.py:18-43).  Membership is decided in exact integer
    arithmetic (n adj(S) against det S), not by comparing floats with 0 and 1."""
    S = np.asarray(np.round(S), dtype=np.int64)
    det = int(round(np.linalg.det(S)))
    adj = np.asarray(np.round(np.linalg.inv(S) * det), dtype=np.int64)  # adjugate: S adj = det 1
    corners = np.array([[i, j, k] for i in (0, 1) for j in (0, 1) for k in (0, 1)]) @ S
    spans = [range(int(lo), int(hi)) for lo, hi in zip(corners.min(axis=0), corners.max(axis=0))]
    inside = []
    for n in ((i, j, k) for i in spans[0] for j in spans[1] for k in spans[2]):
        scaled = np.array(n, dtype=np.int64) @ adj  # = det * (n S^-1)
        if det < 0:
            scaled, bound = -scaled, -det
        else:
            bound = det
        if np.all((scaled >= 0) & (scaled < bound)):
            inside.append(n)
    return np.array(inside, dtype=np.int64).reshape(-1, 3), adj, det


def get_supercell_copies(latvec, S):
    """Translations (Cartesian) of the primitive cells that make up the supercell ``S . latvec``."""
    n, _, _ = _integer_points_in_cell(S)
    return n @ np.asarray(latvec, dtype=float)


def get_supercell_kpts(supercell):
    """Primitive-cell k-points that fold onto the supercell's Gamma point: ``k = m S^-T b`` for the integer
    vectors m inside one cell of ``S^T`` (b = primitive reciprocal vectors incl. 2 pi)."""
    m, adj, det = _integer_points_in_cell(np.asarray(supercell.S).T)
    reduced = (m @ adj) / det
    b = 2 * np.pi * np.linalg.inv(supercell.original_cell.lattice_vectors()).T
    return reduced @ b


def get_supercell(cell, S):
    """Simulation cell ``S . a`` holding every primitive copy of every atom (supercell.py:46-75 without
    pyscf); the copies of one primitive atom are consecutive.  Carries ``original_cell``, ``S``, ``scale``."""
    S = np.asarray(S)
    shifts = get_supercell_copies(cell.lattice_vectors(), S)
    ncopy = len(shifts)
    atoms = [(name, np.asarray(xyz) + R) for name, xyz in cell._atom for R in shifts]
    charges = np.repeat(cell.atom_charges(), ncopy)
    nelec = tuple(n * ncopy for n in cell.nelec)
    sc = Cell(atoms, cell._basis, cell._ecp, nelec, charges, S @ cell.lattice_vectors())
    sc.original_cell, sc.S, sc.scale = cell, S.tolist(), abs(int(np.round(np.linalg.det(S))))
    return sc


def create_supercell_twists(supercell, mf, tol=12):
    """Groups the mean field's k-points by the supercell twist they belong to: two k-points share a twist
    when they differ by a supercell reciprocal vector (twists.py:34-65).  Returns the distinct twists
    (Cartesian, reduced into the first supercell reciprocal cell, sorted), how many k-points each has and
    the indices of those k-points."""
    g = supercell.reciprocal_vectors()
    reduced = np.around(np.asarray(mf.kpts) @ np.linalg.inv(g), tol) % 1
    folded = np.round(reduced @ g, tol) + 0.0  # + 0.0 turns -0.0 into 0.0 so equal twists compare equal
    groups = {}
    for index, twist in enumerate(map(tuple, folded)):
        groups.setdefault(twist, []).append(index)
    order = sorted(groups)
    return {"twists": np.array(order).reshape(-1, 3), "counts": np.array([len(groups[t]) for t in order]),
            "primitive_ks": [np.array(groups[t]) for t in order]}


class KMF:
    """k-point mean-field look-alike in KUHF layout: ``kpts (nk,3)``, ``mo_coeff[s][k] (A, nmo)``,
    ``mo_occ[s][k] (nmo,)`` (read at pyqmc/pyscftools.py:140-186, 206-219)."""

    def __init__(self, kpts, mo_coeff, mo_occ):
        self.kpts = np.asarray(kpts, dtype=float)
        self.mo_coeff = np.asarray(mo_coeff)
        self.mo_occ = np.asarray(mo_occ)

    def to_uhf(self, *args):
        return self


# ---- orbital-evaluator tables (pbcgto.py:518-621) ---------------------------------------------
def max_distance_in_cell(lvecs):
    combos = np.array([[1.0, 1.0, 1.0], [-1.0, 1.0, 1.0], [1.0, -1.0, 1.0], [1.0, 1.0, -1.0]])
    vecs = combos @ lvecs
    d = np.sum(vecs**2, axis=-1)
    return vecs[np.argmax(d)] / 2


def image_tables(cell, kpts, eval_gto_precision=None):
    """Sorted image list ``Ls``, ``num_Ls`` per atom, r^2 cutoffs per atom / per shell and the
    phase table ``exp(i Ls . k)`` -- what ``PeriodicAtomicOrbitalEvaluator.__init__`` builds
    (pbcgto.py:594-621) with ``max_Ls`` (551-591)."""
    prec = 1e-2 if eval_gto_precision is None else eval_gto_precision
    t = _basis.shell_tables(cell)
    rcut = shell_rcut(cell, prec)
    Ls = cell.get_lattice_Ls(rcut=rcut.max(), dimension=3)
    Ls = Ls[np.argsort(np.linalg.norm(Ls, axis=1))]
    expcutoff = -3.5 * np.log(prec)
    natom = len(cell._atom)
    v = max_distance_in_cell(cell.lattice_vectors())
    r2 = np.sum((v - Ls) ** 2, axis=-1)
    nshell = len(t["shell_l"])
    l_cutoff = np.zeros(nshell)
    atom_cutoff = np.zeros(natom)
    Lmax_a = np.zeros(natom, dtype=np.int32)
    for sh in range(nshell):
        a, l = int(t["shell_atom"][sh]), int(t["shell_l"][sh])
        al = t["exps"][t["prim_off"][sh]:t["prim_off"][sh + 1]]
        cf = t["coefs"][t["prim_off"][sh]:t["prim_off"][sh + 1]]
        log_c = np.log(np.abs(cf))
        if l == 0:
            l_cutoff[sh] = np.amax((expcutoff + log_c) / al)
        else:
            r2sup = 0.5 * l / np.amin(al)
            lconst = 0.5 * np.log(r2sup) * l
            l_cutoff[sh] = np.amax((expcutoff + log_c + lconst) / al)
        atom_cutoff[a] = max(atom_cutoff[a], l_cutoff[sh])
        with np.errstate(divide="ignore", invalid="ignore"):
            min_exp = np.amin(al[None, :] * r2[:, None] - log_c[None, :] - 0.5 * np.log(r2)[:, None] * l, axis=1)
        where = np.where(min_exp < expcutoff)[0]
        lm = where.max() + 1 if len(where) > 0 else 1
        Lmax_a[a] = max(Lmax_a[a], lm)
    Lmax = int(Lmax_a.max())
    phases = np.real_if_close(np.exp(1j * Ls[:Lmax] @ np.asarray(kpts).T))
    return dict(Ls=np.ascontiguousarray(Ls[:Lmax]), num_Ls=Lmax_a, atom_cutoff=atom_cutoff, l_cutoff=l_cutoff,
                phases=phases)


# ---- minimal image (distance.py:83-121) --------------------------------------------------------
def minimal_image_tables(latvec):
    """(mode, shifts (27,3)): mode 1 diagonal, 2 orthogonal, 3 general (27-shift argmin)."""
    latvec = np.asarray(latvec, dtype=float)
    tol = 1e-10

    def is_diag(M):
        return np.all(np.abs(M - np.diag(np.diagonal(M))) < tol)

    if is_diag(latvec):
        mode = 1
    elif is_diag(np.dot(latvec, latvec.T)):
        mode = 2
    else:
        mode = 3
    mesh_grid = np.meshgrid(*[np.array(range(3)) for _ in range(3)])
    point_list = np.stack([m.ravel() for m in mesh_grid], axis=0).T - 1
    return mode, np.dot(point_list, latvec)


# ---- Ewald tables (ewald.py:93-200, 356-379) ----------------------------------------------------
def ewald_tables(cell, ewald_gmax=200, nlatvec=1):
    """alpha, real-space displacements, selected reciprocal points / weights, ion structure factor
    and the position-independent constants of ``Ewald.__init__``."""
    latvec = cell.lattice_vectors()
    charges = np.asarray(cell.atom_charges(), dtype=float)
    coords = cell.atom_coords()
    XYZ = np.meshgrid(*[np.arange(-nlatvec, nlatvec + 1)] * 3, indexing="ij")
    xyz = np.stack(XYZ, axis=-1).reshape((-1, 3))
    disp = np.dot(xyz, latvec)
    cellvolume = np.linalg.det(latvec)
    recvec = np.linalg.inv(latvec).T
    smallestheight = np.amin(1 / np.linalg.norm(recvec, axis=1))
    alpha = 5.0 / smallestheight
    # generate_positive_gpoints(gmax) enumerates (2 gmax + 1)^3 / 2 integer triples and select_big
    # keeps gweight > 1e-10; the same points in the same order follow from a box that encloses the
    # sphere |G| < Gcut where the weight first drops below that threshold.
    gcut2 = 4 * alpha**2 * 40.0  # exp(-40) / (V G^2) << 1e-10 for any physical cell
    while 4 * np.pi * np.exp(-gcut2 / (4 * alpha**2)) / (abs(cellvolume) * gcut2) > 1e-11:
        gcut2 *= 1.5
    nmax = np.minimum(np.ceil(np.sqrt(gcut2) * np.linalg.norm(latvec, axis=1) / (2 * np.pi)).astype(int) + 1, ewald_gmax)
    nx, ny, nz = (int(v) for v in nmax)
    gXpos = np.mgrid[1:nx + 1, -ny:ny + 1, -nz:nz + 1].reshape(3, -1)
    gX0Ypos = np.mgrid[0:1, 1:ny + 1, -nz:nz + 1].reshape(3, -1)
    gX0Y0Zpos = np.mgrid[0:1, 0:1, 1:nz + 1].reshape(3, -1)
    gpts = np.concatenate([gXpos, gX0Ypos, gX0Y0Zpos], axis=1)
    gpoints = np.einsum("ji,jk->ik", gpts, recvec * 2 * np.pi)
    gsquared = np.einsum("jk,jk->j", gpoints, gpoints)
    gweight = 4 * np.pi * np.exp(-gsquared / (4 * alpha**2))
    gweight /= cellvolume * gsquared
    big = gweight > 1e-10
    gpoints, gweight = gpoints[big], gweight[big]
    i_sum = np.sum(charges)
    ii_sum2 = np.sum(charges**2)
    ii_sum = (i_sum**2 - ii_sum2) / 2
    ijconst = -np.pi / (cellvolume * alpha**2)
    squareconst = -alpha / np.sqrt(np.pi) + ijconst / 2
    ii_const = ii_sum * ijconst + ii_sum2 * squareconst
    # ion-ion (ewald.py:202-240)
    from scipy.special import erfc

    if len(charges) == 1:
        ion_ion_real = 0.0
    else:
        mode, shifts = minimal_image_tables(latvec)
        ion_ion_real = 0.0
        for i in range(len(charges)):
            for j in range(i + 1, len(charges)):
                d = minimal_image(coords[i] - coords[j], latvec, mode, shifts)
                r = np.linalg.norm(d[None, :] + disp, axis=-1)
                ion_ion_real += charges[i] * charges[j] * np.sum(erfc(alpha * r) / r)
    GdotR = np.dot(gpoints, coords.T)
    ion_exp = np.dot(np.exp(1j * GdotR), charges)
    ion_ion_rec = np.dot(gweight, np.abs(ion_exp) ** 2)
    return dict(alpha=float(alpha), disp=disp, gpoints=np.ascontiguousarray(gpoints), gweight=np.ascontiguousarray(gweight),
                ion_exp=ion_exp, ijconst=float(ijconst), squareconst=float(squareconst), i_sum=float(i_sum),
                ii=float(ion_ion_real + ion_ion_rec + ii_const))


def minimal_image(d, latvec, mode, shifts):
    """distance.py:133-159 for one displacement vector (host-side set-up only)."""
    d = np.asarray(d, dtype=float)
    if mode == 1:
        out = d.copy()
        for i in range(3):
            L = latvec[i, i]
            out[i] = (out[i] + L / 2) % L - L / 2
        return out
    if mode == 2:
        frac = d @ np.linalg.inv(latvec)
        frac = (frac + 0.5) % 1 - 0.5
        return frac @ latvec
    allv = d[None, :] + shifts
    return allv[np.argmin(np.sum(allv**2, axis=-1))]
