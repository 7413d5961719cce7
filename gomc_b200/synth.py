"""Synthetic benchmark / parity systems and writers for the
GOMC input files (PDB / PSF / Mie parameter file / in.conf) that let the
reference itself (oracle/_ref probe) read exactly the same system.

Pure numpy host code: no CUDA, no oracle.  Coordinates are rounded to the
three decimals a PDB file carries so that the arrays here and what the
reference parses are bit-identical.

Force-field derivation follows Forcefield::Init (src/Forcefield.cpp:77-84) and
FFParticle::Blend (src/FFParticle.cpp:155-199).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field

import numpy as np

VDW_STD, VDW_SHIFT, VDW_SWITCH, VDW_EXP6 = 0, 1, 2, 3
_POT_NAME = {VDW_STD: "VDW", VDW_SHIFT: "SHIFT", VDW_SWITCH: "SWITCH", VDW_EXP6: "EXP6"}
KCAL_PER_MOL_TO_K = 503.21959899      # src/FFSetup.cpp:25
RIJ_OVER_2_TO_SIG = 1.7817974362807   # src/FFSetup.cpp:26


@dataclass
class MolKind:
    name: str                 # residue name (<= 4 chars)
    atom_names: list          # per atom
    atom_types: list          # per atom, FF type name
    charges: list             # per atom (e)
    masses: list
    bonds: list = field(default_factory=list)    # (i, j) local indices
    angles: list = field(default_factory=list)   # (i, j, k)
    dihedrals: list = field(default_factory=list)  # (i, j, k, l)
    local_xyz: np.ndarray | None = None           # rigid template, (natoms, 3)


@dataclass
class ForceField:
    type_names: list          # FF atom types, order == kind index
    epsilon: np.ndarray       # K, per type
    sigma: np.ndarray         # Angstrom, per type
    n: np.ndarray             # Mie exponent, per type
    vdw_kind: int = VDW_STD
    r_cut: float = 10.0
    r_cut_low: float = 1.0
    r_switch: float = 0.0
    r_cut_coulomb: float = 10.0
    tolerance: float = 1e-5
    ewald: bool = True
    electrostatic: bool = True
    lrc: bool = True
    is_martini: bool = False       # ParaTypeMARTINI + Potential SWITCH -> FF_SWITCH_MARTINI
    dielectric: float = 15.0       # Martini only (src/ConfigSetup.cpp:1684-1686 default)
    bond_params: list = field(default_factory=list)   # (t1, t2, b0)
    angle_params: list = field(default_factory=list)  # (t1, t2, t3, theta0)
    angle_k: float = 999999999999.0                   # rigid unless a flexible kind sets it
    dihedral_params: list = field(default_factory=list)  # (t1, t2, t3, t4, Kchi, n, delta)

    # ---- derived exactly as the reference derives them -------------------
    @property
    def alpha(self) -> float:           # src/Forcefield.cpp:80
        return math.sqrt(-math.log(self.tolerance)) / self.r_cut_coulomb

    @property
    def recip_rcut(self) -> float:      # src/Forcefield.cpp:82
        return -2.0 * math.log(self.tolerance) / self.r_cut_coulomb

    def tables(self):
        """sigmaSq, epsilon_cn, n as [K*K] tables, index k1 + k2*K
        (src/FFParticle.cpp:155-199, arithmetic-mean sigma, geometric epsilon,
        arithmetic n)."""
        K = len(self.type_names)
        sig = np.zeros(K * K)
        eps_cn = np.zeros(K * K)
        nn = np.zeros(K * K)
        for i in range(K):
            for j in range(K):
                idx = i + j * K
                if self.vdw_kind == VDW_EXP6:   # geometric mean, FFParticle.cpp:165-167
                    n_ij = math.sqrt(self.n[i] * self.n[j])
                else:
                    n_ij = (self.n[i] + self.n[j]) * 0.5
                cn = n_ij / (n_ij - 6.0) * math.pow(n_ij / 6.0, 6.0 / (n_ij - 6.0))
                s = (self.sigma[i] + self.sigma[j]) * 0.5
                e = math.sqrt(self.epsilon[i] * self.epsilon[j])
                sig[idx] = s * s
                eps_cn[idx] = cn * e
                nn[idx] = n_ij
        return sig, eps_cn, nn

    def exp6_tables(self):
        """rMin, expConst, rMaxSq of FF_EXP6::Init (src/FFExp6.h:99-147).  The
        reference finds the two roots with a float-precision Brent solver
        (lib/NumLib.h:232-300, tol 1e-7); here scipy's brentq on the same
        functions -- parity tests feed the SAME arrays to the oracle and the
        engine, and the golden fixtures carry the reference's own values."""
        from scipy.optimize import brentq
        sig, eps_cn, nn = self.tables()
        K = len(self.type_names)
        r_min, exp_c, r_max_sq = np.zeros(K * K), np.zeros(K * K), np.zeros(K * K)
        for i in range(K):
            for j in range(K):
                idx = i + j * K
                a, sigma = nn[idx], math.sqrt(sig[idx])
                eps = math.sqrt(self.epsilon[i] * self.epsilon[j])
                exp_c[idx] = eps * a / (a - 6.0)
                if sigma == 0.0:
                    continue
                f1 = lambda x: (6.0 / a) * math.exp(a * (1.0 - sigma / x)) - (x / sigma) ** 6
                r_min[idx] = brentq(f1, sigma, 3.0 * sigma, xtol=1e-9)
                rm = r_min[idx]
                f2 = lambda x: (-1.0 / rm) * math.exp(a * (1.0 - x / rm)) + (rm / x) ** 6 / x
                r_max_sq[idx] = brentq(f2, 1e-3 * sigma, sigma, xtol=1e-9) ** 2
        return r_min, exp_c, r_max_sq


@dataclass
class System:
    name: str
    ff: ForceField
    mol_kinds: list           # [MolKind]
    axis: np.ndarray          # (3,) orthogonal box
    x: np.ndarray
    y: np.ndarray
    z: np.ndarray
    kind: np.ndarray          # int32 per atom: FF type index
    mol: np.ndarray           # int32 per atom: molecule index
    charge: np.ndarray        # per atom
    mol_start: np.ndarray     # int32 [nMols+1]
    mol_kind: np.ndarray      # int32 per molecule
    # non-orthogonal cell (None for an orthogonal box): rows = cell vectors a, b, c;
    # cell_basis = the normalised rows, cell_basis_inv = its inverse
    # (BoxDimensionsNonOrth::Init, src/BoxDimensionsNonOrth.cpp:14-110); axis = edge lengths
    cell_vectors: np.ndarray | None = None
    cell_basis: np.ndarray | None = None
    cell_basis_inv: np.ndarray | None = None

    @property
    def n_atoms(self) -> int:
        return int(self.x.shape[0])

    @property
    def n_mols(self) -> int:
        return int(self.mol_kind.shape[0])

    def com(self):
        """Geometric centre of every molecule after unwrapping about its first
        atom, wrapped back into the box (what COM::CalcCOM does with
        uniform weights; used only as the torque reference point)."""
        if self.cell_basis is not None:
            r = np.stack([self.x, self.y, self.z], 1)
            u = r @ self.cell_basis_inv
            ref = u[self.mol_start[:-1]]
            lens = np.diff(self.mol_start)
            d = u - np.repeat(ref, lens, axis=0)
            d -= self.axis * np.round(d / self.axis)
            c = ref + np.add.reduceat(d, self.mol_start[:-1], axis=0) / lens[:, None]
            c = np.mod(c, self.axis) @ self.cell_basis
            return c[:, 0].copy(), c[:, 1].copy(), c[:, 2].copy()
        L = self.axis
        cx = np.zeros(self.n_mols)
        cy = np.zeros(self.n_mols)
        cz = np.zeros(self.n_mols)
        for arr, out, ax in ((self.x, cx, L[0]), (self.y, cy, L[1]), (self.z, cz, L[2])):
            ref = arr[self.mol_start[:-1]]
            lens = np.diff(self.mol_start)
            refa = np.repeat(ref, lens)
            d = arr - refa
            d -= ax * np.round(d / ax)
            s = np.add.reduceat(d, self.mol_start[:-1])
            c = ref + s / lens
            out[:] = np.mod(c, ax)
        return cx, cy, cz


# --------------------------------------------------------------------------
def _rand_rotations(rng, n):
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    R = np.empty((n, 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z)
    R[:, 0, 1] = 2 * (x * y - z * w)
    R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w)
    R[:, 1, 1] = 1 - 2 * (x * x + z * z)
    R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w)
    R[:, 2, 1] = 2 * (y * z + x * w)
    R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def _lattice(n, L, rng, jitter):
    m = int(math.ceil(n ** (1.0 / 3.0) - 1e-9))
    a = L / m
    g = np.arange(m)
    pts = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    sel = rng.permutation(pts.shape[0])[:n]
    sel.sort()
    pos = (pts[sel] + 0.5) * a
    pos += rng.uniform(-jitter, jitter, size=pos.shape)
    return pos


def _assemble(name, ff, mol_kinds, counts, L, seed, jitter=0.3, cell_vectors=None):
    rng = np.random.default_rng(seed)
    n_mols = int(sum(counts))
    centres = _lattice(n_mols, L, rng, jitter)
    if cell_vectors is not None:      # lattice sites in fractional coordinates of the cell
        cell_vectors = np.asarray(cell_vectors, dtype=np.float64)
        centres = (centres / L) @ cell_vectors
    order = rng.permutation(n_mols)          # mix kinds over lattice sites
    type_index = {t: i for i, t in enumerate(ff.type_names)}
    xs, kinds, mols, charges, starts, mkind = [], [], [], [], [0], []
    site = 0
    m = 0
    for k, (mk, cnt) in enumerate(zip(mol_kinds, counts)):
        na = len(mk.atom_names)
        tmpl = mk.local_xyz if mk.local_xyz is not None else np.zeros((na, 3))
        tmpl = tmpl - tmpl.mean(axis=0)
        R = _rand_rotations(rng, cnt) if na > 1 else np.tile(np.eye(3), (cnt, 1, 1))
        c = centres[order[site:site + cnt]]
        site += cnt
        pos = np.einsum("mij,aj->mai", R, tmpl) + c[:, None, :]
        xs.append(pos.reshape(-1, 3))
        kinds.append(np.tile([type_index[t] for t in mk.atom_types], cnt))
        charges.append(np.tile(np.asarray(mk.charges, dtype=np.float64), cnt))
        mols.append(np.repeat(np.arange(m, m + cnt), na))
        starts.extend(starts[-1] + na * (np.arange(cnt) + 1))
        mkind.extend([k] * cnt)
        m += cnt
    pos = np.concatenate(xs)
    extra = {}
    if cell_vectors is not None:
        lengths = np.linalg.norm(cell_vectors, axis=1)
        basis = cell_vectors / lengths[:, None]
        basis_inv = np.linalg.inv(basis)
        pos = np.round(pos, 3)
        u = np.mod(pos @ basis_inv, lengths)          # wrap in unslant space
        u = np.minimum(u, np.nextafter(lengths, 0))
        pos = u @ basis
        extra = dict(cell_vectors=cell_vectors, cell_basis=basis, cell_basis_inv=basis_inv)
        axis = lengths
    else:
        pos = np.mod(pos, L)
        pos = np.round(pos, 3)
        pos[pos >= L] -= L     # rounding may land exactly on L
        pos = np.round(pos, 3)
        axis = np.array([L, L, L], dtype=np.float64)
    return System(
        **extra,
        name=name, ff=ff, mol_kinds=list(mol_kinds),
        axis=axis,
        x=np.ascontiguousarray(pos[:, 0]), y=np.ascontiguousarray(pos[:, 1]),
        z=np.ascontiguousarray(pos[:, 2]),
        kind=np.concatenate(kinds).astype(np.int32),
        mol=np.concatenate(mols).astype(np.int32),
        charge=np.concatenate(charges),
        mol_start=np.asarray(starts, dtype=np.int32),
        mol_kind=np.asarray(mkind, dtype=np.int32))


def _spce_kind():
    r, ang = 1.0, math.radians(109.47)
    h = np.array([[0.0, 0.0, 0.0],
                  [r * math.sin(ang / 2), r * math.cos(ang / 2), 0.0],
                  [-r * math.sin(ang / 2), r * math.cos(ang / 2), 0.0]])
    return MolKind("SPCE", ["O1", "H1", "H2"], ["OW", "HW", "HW"],
                   [-0.8476, 0.4238, 0.4238], [15.9994, 1.008, 1.008],
                   bonds=[(0, 1), (0, 2)], angles=[(1, 0, 2)], local_xyz=h)


def make_argon(n_atoms=4000, density=0.0213, seed=123, r_cut=10.0, vdw_kind=VDW_STD,
               r_switch=0.0):
    """Config 1: LJ argon, sigma 3.4 A, eps/k 119.8 K, no electrostatics."""
    L = round((n_atoms / density) ** (1.0 / 3.0), 3)
    ff = ForceField(["AR"], np.array([119.8]), np.array([3.4]), np.array([12.0]),
                    vdw_kind=vdw_kind, r_cut=r_cut, r_cut_coulomb=r_cut,
                    r_switch=r_switch, ewald=False, electrostatic=False,
                    lrc=(vdw_kind == VDW_STD))
    mk = MolKind("AR", ["AR"], ["AR"], [0.0], [39.948])
    return _assemble(f"argon{n_atoms}", ff, [mk], [n_atoms], L, seed)


def triclinic_cell(L, angles_deg=(80.0, 75.0, 70.0)):
    """Cell vectors with edge length L and the given (alpha, beta, gamma)
    (same construction as BoxDimensions::Init, src/BoxDimensions.cpp:21-39)."""
    al, be, ga = (math.cos(math.radians(a)) for a in angles_deg)
    t = (al - be * ga) / math.sqrt(1.0 - ga * ga)
    return L * np.array([[1.0, 0.0, 0.0], [ga, math.sqrt(1.0 - ga * ga), 0.0],
                         [be, t, math.sqrt(1.0 - be * be - t * t)]])


def make_spce(n_mols=10000, density=0.0334, seed=123, r_cut=10.0, r_cut_coulomb=None,
              tolerance=1e-5, vdw_kind=VDW_STD, r_switch=0.0, ewald=True, cell_vectors=None):
    """Configs 2 and 4: rigid SPC/E water, Ewald on."""
    rcc = r_cut if r_cut_coulomb is None else r_cut_coulomb
    L = round((n_mols / density) ** (1.0 / 3.0), 3)
    ff = ForceField(["OW", "HW"], np.array([78.2, 0.0]), np.array([3.166, 0.0]),
                    np.array([12.0, 12.0]), vdw_kind=vdw_kind, r_cut=r_cut,
                    r_cut_coulomb=rcc, tolerance=tolerance, r_switch=r_switch,
                    ewald=ewald, electrostatic=True, lrc=(vdw_kind == VDW_STD),
                    bond_params=[("OW", "HW", 1.0)],
                    angle_params=[("HW", "OW", "HW", 109.47)])
    return _assemble(f"spce{n_mols}", ff, [_spce_kind()], [n_mols], L, seed,
                     jitter=0.15, cell_vectors=cell_vectors)


def make_electrolyte(n_water=330000, n_pairs=5000, seed=123, r_cut=10.0,
                     tolerance=1e-5):
    """Config 5: SPC/E + Na+ + Cl- (1:1), 1 000 002 atoms at the default size."""
    n_mols = n_water + 2 * n_pairs
    L = round((n_mols / 0.0334) ** (1.0 / 3.0), 3)
    ff = ForceField(["OW", "HW", "NA", "CL"], np.array([78.2, 0.0, 65.4, 50.3]),
                    np.array([3.166, 0.0, 2.35, 4.40]), np.array([12.0] * 4),
                    r_cut=r_cut, r_cut_coulomb=r_cut, tolerance=tolerance,
                    bond_params=[("OW", "HW", 1.0)],
                    angle_params=[("HW", "OW", "HW", 109.47)])
    na = MolKind("NA", ["NA"], ["NA"], [1.0], [22.99])
    cl = MolKind("CL", ["CL"], ["CL"], [-1.0], [35.45])
    return _assemble(f"electrolyte{n_water}_{n_pairs}", ff, [_spce_kind(), na, cl],
                     [n_water, n_pairs, n_pairs], L, seed, jitter=0.15)


def make_mixture(n_a=150, n_b=100, seed=5, L=26.0, r_cut=8.0, vdw_kind=VDW_STD,
                 r_switch=0.0, n_b_exp=14.0, martini=False, ewald=True, du_eps=0.0,
                 du_sigma=0.0):
    """Small two-kind Mie mixture with charged dimers: exercises the kind table
    (non-integer n/2 via n=13 cross terms), charged + neutral atoms in one
    molecule and multi-kind LRC."""
    ff = ForceField(["CA", "CB", "DU"], np.array([98.0, 46.0, du_eps]),
                    np.array([3.75, 3.0, du_sigma]), np.array([12.0, n_b_exp, 12.0]),
                    vdw_kind=vdw_kind, r_cut=r_cut, r_cut_coulomb=r_cut,
                    r_switch=r_switch, tolerance=1e-5, ewald=ewald,
                    lrc=(vdw_kind == VDW_STD), is_martini=martini,
                    bond_params=[("CB", "CB", 1.5), ("CB", "DU", 0.8)],
                    angle_params=[("CB", "CB", "DU", 120.0)])
    a = MolKind("AAA", ["C1"], ["CA"], [0.0], [16.0])
    tb = np.array([[0.0, 0.0, 0.0], [1.5, 0.0, 0.0], [1.9, 0.693, 0.0]])
    b = MolKind("BBB", ["B1", "B2", "D1"], ["CB", "CB", "DU"], [0.35, -0.6, 0.25],
                [14.0, 14.0, 1.0], bonds=[(0, 1), (1, 2)], angles=[(0, 1, 2)],
                local_xyz=tb)
    return _assemble(f"mixture{n_a}_{n_b}", ff, [a, b], [n_a, n_b], L, seed,
                     jitter=0.2)


def make_pentane(n_mols=150, L=34.0, seed=31, r_cut=10.0, charged=False, ewald=True):
    """Config 3's molecule: TraPPE-UA n-pentane (CH3 eps/k 98 K sigma 3.75 A, CH2 46 K
    3.95 A, bond 1.54 A, angle 114 deg; parameters as in the reference's
    test/input/Systems/PEN_HEX/Base force-field file), five united atoms, flexible
    angles and dihedrals -- the multi-site chain CBMC grows, with one intramolecular
    non-bonded pair (sites 0 and 4, four bonds apart) under Exclude 1-4.  `charged` puts
    partial charges on the sites (net zero) so that the Ewald deltas are not the degenerate
    zero of the all-neutral alkane."""
    q = [0.25, -0.15, -0.2, -0.15, 0.25] if charged else [0.0] * 5
    ff = ForceField(["CH3", "CH2"], np.array([98.0, 46.0]), np.array([3.75, 3.95]),
                    np.array([12.0, 12.0]), r_cut=r_cut, r_cut_coulomb=r_cut, tolerance=1e-5,
                    ewald=ewald, electrostatic=True,
                    bond_params=[("CH3", "CH2", 1.54), ("CH2", "CH2", 1.54)],
                    angle_params=[("CH3", "CH2", "CH2", 114.0), ("CH2", "CH2", "CH2", 114.0)],
                    angle_k=31250.0,
                    dihedral_params=[("CH3", "CH2", "CH2", "CH2", kc, n_, dl) for kc, n_, dl in
                                     ((2156.3, 0, 90.0), (-355.03, 1, 180.0), (68.19, 2, 0.0),
                                      (-791.32, 3, 180.0))])
    # all-trans zig-zag in the xy plane
    th = math.radians(114.0)
    dx, dy = 1.54 * math.sin(th / 2), 1.54 * math.cos(th / 2)
    tmpl = np.array([[i * dx, (i % 2) * dy, 0.0] for i in range(5)])
    mk = MolKind("PEN", ["C1", "C2", "C3", "C4", "C5"], ["CH3", "CH2", "CH2", "CH2", "CH3"], q,
                 [15.035, 14.027, 14.027, 14.027, 15.035],
                 bonds=[(0, 1), (1, 2), (2, 3), (3, 4)], angles=[(0, 1, 2), (1, 2, 3), (2, 3, 4)],
                 dihedrals=[(0, 1, 2, 3), (1, 2, 3, 4)], local_xyz=tmpl)
    return _assemble(f"pentane{n_mols}{'q' if charged else ''}", ff, [mk], [n_mols], L, seed,
                     jitter=0.2)


# --------------------------------------------------------------------------
# GOMC input writers (consumed by oracle/_ref/gomc_probe_*)

def write_gomc_inputs(sys: System, out_dir: str, multiparticle=True,
                      cached_fourier=False, run_steps=0, pressure_calc=False, npt=False,
                      second: System | None = None, gemc_freqs=None):
    """second: the system of box 1 (GEMC two-box input; same force field and molecule kinds);
    gemc_freqs: {"DisFreq": ..., "RotFreq": ..., "SwapFreq": ..., "RegrowthFreq": ...,
    "VolFreq": ...} for that case."""
    os.makedirs(out_dir, exist_ok=True)
    ff = sys.ff
    # ---- parameter file (Mie / "EXOTIC" style, epsilon in K) --------------
    with open(os.path.join(out_dir, "par.inp"), "w") as f:
        f.write("* synthetic Mie parameter file written by gomc_b200.synth\n*\n\n")
        f.write("BONDS\n")
        for t1, t2, b0 in ff.bond_params:
            f.write(f"{t1}\t{t2}\t999999999999\t{b0}\n")
        f.write("\nANGLES\n")
        for t1, t2, t3, th in ff.angle_params:
            f.write(f"{t1}\t{t2}\t{t3}\t{ff.angle_k!r}\t{th}\n")
        if ff.is_martini:   # CHARMM units: -eps in kcal/mol, Rmin/2 (src/FFSetup.cpp:265-279)
            f.write("\nDIHEDRALS\n\nNONBONDED\n")
            for t, e, s, n in zip(ff.type_names, ff.epsilon, ff.sigma, ff.n):
                f.write(f"{t}\t0.0\t{-float(e) / KCAL_PER_MOL_TO_K!r}\t"
                        f"{float(s) / RIJ_OVER_2_TO_SIG!r}\n")
        else:
            f.write("\nDIHEDRALS\n")
            for t1, t2, t3, t4, kc, nn_, dl in ff.dihedral_params:
                f.write(f"{t1}\t{t2}\t{t3}\t{t4}\t{kc!r}\t{int(nn_)}\t{dl!r}\n")
            f.write("\nNONBONDED_MIE\n")
            for t, e, s, n in zip(ff.type_names, ff.epsilon, ff.sigma, ff.n):
                f.write(f"{t}\t{float(e)!r}\t{float(s)!r}\t{float(n)!r}\n")
        f.write("\nEND\n")
    # ---- PDB + PSF (one pair per box) --------------------------------------
    for bi, bs in enumerate([sys] + ([second] if second is not None else [])):
        n = bs.n_atoms
        with open(os.path.join(out_dir, f"box{bi}.pdb"), "w") as f:
            if bs.cell_vectors is None:   # (triclinic cells come from in.conf only)
                f.write("CRYST1%9.3f%9.3f%9.3f  90.00  90.00  90.00 P 1           1\n"
                        % tuple(bs.axis))
            for a in range(n):
                m = int(bs.mol[a])
                mk = bs.mol_kinds[int(bs.mol_kind[m])]
                la = a - int(bs.mol_start[m])
                f.write("ATOM  %5d %-4s %-4s%1s%4d    %8.3f%8.3f%8.3f%6.2f%6.2f\n" % (
                    (a + 1) % 100000, mk.atom_names[la], mk.name, "A",
                    (m + 1) % 10000, bs.x[a], bs.y[a], bs.z[a], 1.0, 0.0))
            f.write("END\n")
        bonds, angles, dihedrals = [], [], []
        for m in range(bs.n_mols):
            mk = bs.mol_kinds[int(bs.mol_kind[m])]
            s = int(bs.mol_start[m]) + 1
            bonds.extend((s + i, s + j) for i, j in mk.bonds)
            angles.extend((s + i, s + j, s + k) for i, j, k in mk.angles)
            dihedrals.extend((s + i, s + j, s + k, s + l) for i, j, k, l in mk.dihedrals)
        with open(os.path.join(out_dir, f"box{bi}.psf"), "w") as f:
            f.write("PSF\n\n       1 !NTITLE\n REMARKS synthetic system written by gomc_b200.synth\n\n")
            f.write("%8d !NATOM\n" % n)
            for a in range(n):
                m = int(bs.mol[a])
                mk = bs.mol_kinds[int(bs.mol_kind[m])]
                la = a - int(bs.mol_start[m])
                f.write("%8d %-4s %-4d %-4s %-4s %-4s %10.6f %13.4f %11d\n" % (
                    a + 1, "S", m + 1, mk.name, mk.atom_names[la], mk.atom_types[la],
                    mk.charges[la], mk.masses[la], 0))
            f.write("\n%8d !NBOND: bonds\n" % len(bonds))
            for i in range(0, len(bonds), 4):
                f.write("".join("%8d%8d" % b for b in bonds[i:i + 4]) + "\n")
            f.write("\n%8d !NTHETA: angles\n" % len(angles))
            for i in range(0, len(angles), 3):
                f.write("".join("%8d%8d%8d" % t for t in angles[i:i + 3]) + "\n")
            f.write("\n%8d !NPHI: dihedrals\n" % len(dihedrals))
            for i in range(0, len(dihedrals), 2):
                f.write("".join("%8d%8d%8d%8d" % t for t in dihedrals[i:i + 2]) + "\n")
            f.write("\n\n%8d !NIMPHI: impropers\n\n\n" % 0)
            f.write("%8d !NDON: donors\n\n\n%8d !NACC: acceptors\n\n\n" % (0, 0))

    # ---- in.conf ---------------------------------------------------------
    L = sys.axis
    CV = sys.cell_vectors if sys.cell_vectors is not None else np.diag(L)
    multi_site = any(len(k.atom_names) > 1 for k in sys.mol_kinds)
    if second is not None:
        fr = gemc_freqs or {"DisFreq": 0.5, "RotFreq": 0.2, "RegrowthFreq": 0.1, "SwapFreq": 0.2}
        move_lines = "\n".join(f"{k} {v}" for k, v in fr.items())
        L1 = second.axis
        box1_cell = (f"CellBasisVector1 1 {float(L1[0])!r} 0.0 0.0\n"
                     f"CellBasisVector2 1 0.0 {float(L1[1])!r} 0.0\n"
                     f"CellBasisVector3 1 0.0 0.0 {float(L1[2])!r}")
    else:
        box1_cell = ""
        dis = (0.40 if multiparticle else 0.60) - (0.02 if npt else 0.0)
        rot = 0.40 if multiparticle and multi_site else (0.40 if not multiparticle else 0.0)
        move_lines = f"DisFreq {dis}\nRotFreq {rot}"
        if multiparticle:
            move_lines += "\nMultiParticleFreq " + ("0.20" if multi_site else "0.60")
    conf = f"""ExpertMode True
Restart false
PRNG INTSEED
Random_Seed 123
{'ParaTypeMARTINI on' if ff.is_martini else 'ParaTypeMie on'}
{('Dielectric ' + repr(float(ff.dielectric))) if ff.is_martini else ''}
Parameters par.inp
Coordinates 0 box0.pdb
Structure 0 box0.psf
{'Coordinates 1 box1.pdb' if second is not None else ''}
{'Structure 1 box1.psf' if second is not None else ''}
{'GEMC NVT' if second is not None else ''}
Temperature 298.0
Potential {_POT_NAME[ff.vdw_kind]}
{('Rswitch ' + repr(float(ff.r_switch))) if ff.vdw_kind == VDW_SWITCH else ''}
LRC {'true' if ff.lrc else 'false'}
Rcut {float(ff.r_cut)!r}
RcutLow {float(ff.r_cut_low)!r}
Exclude 1-4
Ewald {'true' if ff.ewald else 'false'}
ElectroStatic {'true' if ff.electrostatic else 'false'}
CachedFourier {'true' if cached_fourier else 'false'}
Tolerance {ff.tolerance!r}
1-4scaling false
RcutCoulomb 0 {float(ff.r_cut_coulomb)!r}
{('RcutCoulomb 1 ' + repr(float(ff.r_cut_coulomb))) if second is not None else ''}
PressureCalc {'true 1000' if pressure_calc else 'false'}
RunSteps {max(run_steps, 10)}
EqSteps 5
AdjSteps 5
{'Pressure 1.01325' if npt else ''}
{'VolFreq 0.02' if npt else ''}
{move_lines}
CellBasisVector1 0 {float(CV[0][0])!r} {float(CV[0][1])!r} {float(CV[0][2])!r}
CellBasisVector2 0 {float(CV[1][0])!r} {float(CV[1][1])!r} {float(CV[1][2])!r}
CellBasisVector3 0 {float(CV[2][0])!r} {float(CV[2][1])!r} {float(CV[2][2])!r}
{box1_cell}
CBMC_First 10
CBMC_Nth 8
CBMC_Ang 50
CBMC_Dih 50
OutputName out
RestartFreq false 1000
CheckpointFreq false 1000
CoordinatesFreq false 1000
DCDFreq false 1000
ConsoleFreq true 1000
BlockAverageFreq false 1000
OutEnergy true true
OutPressure false false
OutMolNum true true
OutDensity false false
OutSurfaceTension false false
"""
    with open(os.path.join(out_dir, "in.conf"), "w") as f:
        f.write(conf)
    return os.path.join(out_dir, "in.conf")
