"""TEST INFRASTRUCTURE ONLY: ctypes loader for oracle/libgomc_oracle.so (the
plain-C CPU restatement of the reference hot path) and a reader for the dumps
written by the reference probe (oracle/ref_probe.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; nothing under gomc_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_LIB = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class OrcParams(C.Structure):
    _fields_ = [("vdwKind", C.c_int), ("ewald", C.c_int), ("electrostatic", C.c_int),
                ("kindCount", C.c_int), ("rCut", C.c_double), ("rCutLow", C.c_double),
                ("rOn", C.c_double), ("rCutCoulomb", C.c_double), ("alpha", C.c_double),
                ("recip_rcut", C.c_double), ("axis", C.c_double * 3),
                ("sigmaSq", _dp), ("epsilon_cn", _dp), ("n", _dp),
                ("rMin", _dp), ("expConst", _dp), ("rMaxSq", _dp),
                ("isMartini", C.c_int), ("diElectric_1", C.c_double),
                ("nonOrth", C.c_int), ("cellBasis", C.c_double * 9),
                ("cellBasisInv", C.c_double * 9), ("volume", C.c_double)]


def build(force=False):
    so = os.path.join(_HERE, "libgomc_oracle.so")
    src = os.path.join(_HERE, "gomc_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-f", "oracle/Makefile"], cwd=_ROOT)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        L.orc_calc_en.restype = C.c_double
        L.orc_calc_vir.restype = C.c_double
        L.orc_calc_coulomb.restype = C.c_double
        L.orc_calc_coulomb_vir.restype = C.c_double
        L.orc_energy_lrc.restype = C.c_double
        L.orc_box_reciprocal.restype = C.c_double
        L.orc_mol_reciprocal.restype = C.c_double
        L.orc_mol_reciprocal_l.restype = C.c_double
        L.orc_swap_recip.restype = C.c_double
        L.orc_recip_weighted.restype = C.c_double
        L.orc_recip_weighted.argtypes = [C.c_int] + [_dp] * 4 + [C.c_int] + [_dp] * 6 + \
            [C.c_double, _dp, _dp]
        L.orc_change_recip.restype = None
        L.orc_box_correction.restype = C.c_double
        L.orc_box_self.restype = C.c_double
        L.orc_swap_correction.restype = C.c_double
        L.orc_swap_self.restype = C.c_double
        L.orc_calc_en.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int]
        L.orc_calc_vir.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int]
        L.orc_calc_coulomb.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.orc_calc_coulomb_vir.argtypes = [C.c_void_p, C.c_double, C.c_double]
    return _LIB


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_ip)


class Oracle:
    """Stateless reference calculator bound to one (force field, box)."""

    def __init__(self, *, vdw_kind, ewald, electrostatic, kind_count, r_cut, r_cut_low,
                 r_switch, r_cut_coulomb, alpha, recip_rcut, axis, sigma_sq, epsilon_cn, n,
                 r_min=None, exp_const=None, r_max_sq=None, is_martini=0, dielectric=1.0,
                 cell_basis=None, cell_basis_inv=None, volume=0.0):
        self.L = lib()
        self._keep = [_d(sigma_sq), _d(epsilon_cn), _d(n)]
        if r_min is not None:
            self._keep += [_d(r_min), _d(exp_const), _d(r_max_sq)]
        p = OrcParams()
        p.vdwKind, p.ewald, p.electrostatic = int(vdw_kind), int(ewald), int(electrostatic)
        p.kindCount = int(kind_count)
        p.rCut, p.rCutLow, p.rOn = float(r_cut), float(r_cut_low), float(r_switch)
        p.rCutCoulomb, p.alpha, p.recip_rcut = float(r_cut_coulomb), float(alpha), float(recip_rcut)
        p.axis[0], p.axis[1], p.axis[2] = (float(v) for v in axis)
        p.sigmaSq, p.epsilon_cn, p.n = (k[1] for k in self._keep[:3])
        if r_min is not None:
            p.rMin, p.expConst, p.rMaxSq = (k[1] for k in self._keep[3:6])
        p.isMartini = int(is_martini)
        p.diElectric_1 = 1.0 / float(dielectric)
        p.volume = float(volume)   # BoxDimensions::volume after SetVolume; 0: product of the axes
        p.nonOrth = 0
        if cell_basis is not None:
            p.nonOrth = 1
            for i, v in enumerate(np.asarray(cell_basis, dtype=np.float64).reshape(-1)):
                p.cellBasis[i] = float(v)
            for i, v in enumerate(np.asarray(cell_basis_inv, dtype=np.float64).reshape(-1)):
                p.cellBasisInv[i] = float(v)
        self.p = p
        self.pp = C.byref(p)
        self.lambda_mol = -1
        self.set_lambda(-1)

    def set_lambda(self, mol, lambda_vdw=1.0, lambda_coulomb=1.0, sc_alpha=0.0, sc_sigma_6=0.0,
                   sc_power=0, sc_coul=0, mol_kind=-1):
        """Fractional molecule + soft-core constants (global state of the C oracle; the
        last Oracle constructed / configured wins)."""
        self.lambda_mol, self.lambda_coulomb = int(mol), float(lambda_coulomb)
        self.L.orc_set_lambda(C.c_int(int(mol)), C.c_double(lambda_vdw),
                              C.c_double(lambda_coulomb), C.c_double(sc_alpha),
                              C.c_double(sc_sigma_6), C.c_int(int(sc_power)),
                              C.c_int(int(sc_coul)), C.c_int(int(mol_kind)))

    def lambda_coef(self, mol):
        """Ewald::GetLambdaCoef, src/Ewald.cpp:1598-1602."""
        return float(np.sqrt(self.lambda_coulomb)) if int(mol) == self.lambda_mol else 1.0

    @classmethod
    def from_system(cls, s):
        sig, eps, nn = s.ff.tables()
        ff = s.ff
        return cls(vdw_kind=ff.vdw_kind, ewald=ff.ewald, electrostatic=ff.electrostatic,
                   kind_count=len(ff.type_names), r_cut=ff.r_cut, r_cut_low=ff.r_cut_low,
                   r_switch=ff.r_switch, r_cut_coulomb=ff.r_cut_coulomb, alpha=ff.alpha,
                   recip_rcut=ff.recip_rcut, axis=s.axis, sigma_sq=sig, epsilon_cn=eps, n=nn,
                   is_martini=ff.is_martini, dielectric=ff.dielectric,
                   cell_basis=getattr(s, "cell_basis", None),
                   cell_basis_inv=getattr(s, "cell_basis_inv", None),
                   **(dict(zip(("r_min", "exp_const", "r_max_sq"), ff.exp6_tables()))
                      if ff.vdw_kind == 3 else {}))

    @classmethod
    def from_dump(cls, d, box=0):
        o = cls._from_dump(d, box)
        if "lambda.params" in d and box == 0:
            lp = d["lambda.params"]
            o.set_lambda(int(lp[0]), lp[1], lp[2], lp[3], lp[4], int(lp[5]), int(lp[6]),
                         int(lp[7]))
        return o

    @classmethod
    def _from_dump(cls, d, box=0):
        return cls(vdw_kind=sc(d, "ff.vdwKind"), ewald=sc(d, "ff.ewald"),
                   electrostatic=sc(d, "ff.electrostatic"), kind_count=sc(d, "ff.kindCount"),
                   r_cut=sc(d, "ff.rCut"), r_cut_low=sc(d, "ff.rCutLow"),
                   r_switch=sc(d, "ff.rswitch"),
                   r_cut_coulomb=d["ff.rCutCoulomb"][box], alpha=d["ff.alpha"][box],
                   recip_rcut=d["ff.recip_rcut"][box], axis=d[f"box{box}.axis"],
                   sigma_sq=d["ff.sigmaSq"], epsilon_cn=d["ff.epsilon_cn"], n=d["ff.n"],
                   is_martini=sc(d, "ff.isMartini"),
                   dielectric=(sc(d, "ff.dielectric") if "ff.dielectric" in d else 1.0),
                   cell_basis=d.get(f"box{box}.cellBasis") if not sc(d, f"box{box}.orthogonal") else None,
                   cell_basis_inv=d.get(f"box{box}.cellBasisInv"),
                   **({"r_min": d["ff.rMin"], "exp_const": d["ff.expConst"],
                       "r_max_sq": d["ff.rMaxSq"]} if "ff.rMin" in d else {}))

    # ---- pair functors ---------------------------------------------------
    def calc_en(self, r2, k1, k2):
        return self.L.orc_calc_en(self.pp, r2, k1, k2)

    def calc_vir(self, r2, k1, k2):
        return self.L.orc_calc_vir(self.pp, r2, k1, k2)

    def calc_coulomb(self, r2, qq):
        return self.L.orc_calc_coulomb(self.pp, r2, qq)

    def calc_coulomb_vir(self, r2, qq):
        return self.L.orc_calc_coulomb_vir(self.pp, r2, qq)

    # ---- cell list -------------------------------------------------------
    def cell_list(self, x, y, z, box_atoms):
        n = len(x)
        edge = (C.c_int * 3)()
        ncell = self.L.orc_cell_edges(self.pp, edge)
        (x, px), (y, py), (z, pz) = _d(x), _d(y), _d(z)
        ba, pba = _i(box_atoms)
        cv = np.zeros(len(ba), np.int32)
        cs = np.zeros(ncell + 1, np.int32)
        mp = np.zeros(n, np.int32)
        nb = np.zeros(ncell * 27, np.int32)
        self.L.orc_cell_list_build(self.pp, n, px, py, pz, pba, len(ba),
                                   cv.ctypes.data_as(_ip), cs.ctypes.data_as(_ip),
                                   mp.ctypes.data_as(_ip), nb.ctypes.data_as(_ip))
        return cv, cs, mp, nb.reshape(ncell, 27), tuple(edge)

    # ---- pair path -------------------------------------------------------
    def box_inter(self, x, y, z, kind, mol, charge, box_atoms):
        (x, px), (y, py), (z, pz), (q, pq) = _d(x), _d(y), _d(z), _d(charge)
        (k, pk), (m, pm), (ba, pba) = _i(kind), _i(mol), _i(box_atoms)
        lj, re = C.c_double(), C.c_double()
        self.L.orc_box_inter(self.pp, len(x), px, py, pz, pk, pm, pq, pba, len(ba),
                             C.byref(lj), C.byref(re))
        return lj.value, re.value

    def box_force(self, x, y, z, kind, mol, charge, box_atoms, n_mols):
        (x, px), (y, py), (z, pz), (q, pq) = _d(x), _d(y), _d(z), _d(charge)
        (k, pk), (m, pm), (ba, pba) = _i(kind), _i(mol), _i(box_atoms)
        lj, re = C.c_double(), C.c_double()
        aF = [np.zeros(len(x)) for _ in range(3)]
        mF = [np.zeros(n_mols) for _ in range(3)]
        self.L.orc_box_force(self.pp, len(x), n_mols, px, py, pz, pk, pm, pq, pba, len(ba),
                             C.byref(lj), C.byref(re),
                             *[a.ctypes.data_as(_dp) for a in aF],
                             *[a.ctypes.data_as(_dp) for a in mF])
        return lj.value, re.value, aF, mF

    def mp_transform(self, move_type, box_mols, mol_start, xyz, com, f, rf, vmax, lam, beta,
                     step, seed, key):
        """MultiParticle trial transform: returns (k, inForceRange, newXYZ, newCOM)."""
        (bm, pbm), (ms, pms) = _i(box_mols), _i(mol_start)
        n_mols = len(ms) - 1
        new = [np.array(a, dtype=np.float64, copy=True) for a in xyz]
        ncm = [np.array(a, dtype=np.float64, copy=True) for a in com]
        fa = [_d(a) for a in f]
        ra = [_d(a) for a in rf] if rf is not None else None
        k = [np.zeros(n_mols) for _ in range(3)]
        inr = np.zeros(n_mols, dtype=np.int32)
        u64 = C.c_uint64
        self.L.orc_mp_transform.argtypes = None
        self.L.orc_mp_transform(
            self.pp, C.c_int(move_type), C.c_int(len(bm)), pbm, pms, *[a[1] for a in fa],
            *([a[1] for a in ra] if ra else [None] * 3), C.c_double(vmax), C.c_double(lam),
            C.c_double(beta), u64(int(step)), u64(int(seed)), u64(int(key)),
            *[a.ctypes.data_as(_dp) for a in k], inr.ctypes.data_as(_ip),
            *[a.ctypes.data_as(_dp) for a in new], *[a.ctypes.data_as(_dp) for a in ncm])
        return k, inr, new, ncm

    def bm_transform(self, move_type, box_mols, mol_start, xyz, com, f, rf, vmax, beta, step,
                     seed, key):
        """MultiParticleBrownian trial transform: returns (k, newXYZ, newCOM)."""
        (bm, pbm), (ms, pms) = _i(box_mols), _i(mol_start)
        n_mols = len(ms) - 1
        new = [np.array(a, dtype=np.float64, copy=True) for a in xyz]
        ncm = [np.array(a, dtype=np.float64, copy=True) for a in com]
        fa = [_d(a) for a in f]
        ra = [_d(a) for a in rf] if rf is not None else None
        k = [np.zeros(n_mols) for _ in range(3)]
        u64 = C.c_uint64
        self.L.orc_bm_transform(
            self.pp, C.c_int(move_type), C.c_int(len(bm)), pbm, pms, *[a[1] for a in fa],
            *([a[1] for a in ra] if ra else [None] * 3), C.c_double(vmax), C.c_double(beta),
            u64(int(step)), u64(int(seed)), u64(int(key)), *[a.ctypes.data_as(_dp) for a in k],
            *[a.ctypes.data_as(_dp) for a in new], *[a.ctypes.data_as(_dp) for a in ncm])
        return k, new, ncm

    def bm_coeff(self, box_mols, old_f, old_rf, new_f, new_rf, k, vmax, beta):
        (bm, pbm) = _i(box_mols)
        def ptrs(v):
            if v is None:
                return [None] * 3, []
            a = [_d(t) for t in v]
            return [t[1] for t in a], a
        po, k1 = ptrs(old_f); pr, k2 = ptrs(old_rf); pn, k3 = ptrs(new_f); pq, k4 = ptrs(new_rf)
        pk, k5 = ptrs(k)
        self.L.orc_bm_coeff.restype = C.c_double
        return self.L.orc_bm_coeff(C.c_int(len(bm)), pbm, *po, *pr, *pn, *pq, *pk,
                                   C.c_double(vmax), C.c_double(beta))

    def mp_coeff(self, box_mols, in_force_range, old_f, old_rf, new_f, new_rf, k, vmax, lam,
                 beta):
        (bm, pbm), (ir, pir) = _i(box_mols), _i(in_force_range)
        def ptrs(v):
            if v is None:
                return [None] * 3, []
            a = [_d(t) for t in v]
            return [t[1] for t in a], a
        po, k1 = ptrs(old_f); pr, k2 = ptrs(old_rf); pn, k3 = ptrs(new_f); pq, k4 = ptrs(new_rf)
        pk, k5 = ptrs(k)
        self.L.orc_mp_coeff.restype = C.c_double
        return self.L.orc_mp_coeff(C.c_int(len(bm)), pbm, pir, *po, *pr, *pn, *pq, *pk,
                                   C.c_double(vmax), C.c_double(lam), C.c_double(beta))

    def virial_calc(self, x, y, z, kind, mol, charge, com, box_atoms):
        (x, px), (y, py), (z, pz), (q, pq) = _d(x), _d(y), _d(z), _d(charge)
        (k, pk), (m, pm), (ba, pba) = _i(kind), _i(mol), _i(box_atoms)
        cm = [_d(a) for a in com]
        vT, rT = np.zeros(3), np.zeros(3)
        self.L.orc_virial_calc(self.pp, len(x), px, py, pz, pk, pm, pq, *[a[1] for a in cm],
                               pba, len(ba), vT.ctypes.data_as(_dp), rT.ctypes.data_as(_dp))
        return vT, rT

    def virial_reciprocal(self, box_mols, mol_start, x, y, z, charge, com, kx, ky, kz, hsqr,
                          prefact, sR, sI):
        (bm, pbm), (ms, pms) = _i(box_mols), _i(mol_start)
        arrs = [_d(a) for a in (x, y, z, charge, *com)]
        ks = [_d(a) for a in (kx, ky, kz, hsqr, prefact, sR, sI)]
        wT = np.zeros(3)
        self.L.orc_virial_reciprocal(self.pp, len(bm), pbm, pms, *[a[1] for a in arrs],
                                     len(ks[0][0]), *[a[1] for a in ks],
                                     wT.ctypes.data_as(_dp))
        return wT

    def molecule_inter(self, x, y, z, kind, mol, charge, box_atoms, mol_index, mol_start,
                       mol_len, nx, ny, nz):
        (x, px), (y, py), (z, pz), (q, pq) = _d(x), _d(y), _d(z), _d(charge)
        (k, pk), (m, pm), (ba, pba) = _i(kind), _i(mol), _i(box_atoms)
        (nx, pnx), (ny, pny), (nz, pnz) = _d(nx), _d(ny), _d(nz)
        lj, re = C.c_double(), C.c_double()
        ov = self.L.orc_molecule_inter(self.pp, len(x), px, py, pz, pk, pm, pq, pba, len(ba),
                                       int(mol_index), int(mol_start), int(mol_len),
                                       pnx, pny, pnz, C.byref(lj), C.byref(re))
        return lj.value, re.value, bool(ov)

    def particle_inter(self, x, y, z, kind, mol, charge, box_atoms, mol_index, kind_i, q_i,
                       tx, ty, tz):
        (x, px), (y, py), (z, pz), (q, pq) = _d(x), _d(y), _d(z), _d(charge)
        (k, pk), (m, pm), (ba, pba) = _i(kind), _i(mol), _i(box_atoms)
        (tx, ptx), (ty, pty), (tz, ptz) = _d(tx), _d(ty), _d(tz)
        t = len(tx)
        en, re, ov = np.zeros(t), np.zeros(t), np.zeros(t, np.int32)
        self.L.orc_particle_inter(self.pp, len(x), px, py, pz, pk, pm, pq, pba, len(ba),
                                  int(mol_index), int(kind_i), C.c_double(q_i), t, ptx, pty,
                                  ptz, en.ctypes.data_as(_dp), re.ctypes.data_as(_dp),
                                  ov.ctypes.data_as(_ip))
        return en, re, ov.astype(bool)

    def particle_nonbonded(self, kind_i, q_i, partner_kind, partner_charge, px, py, pz,
                           tx, ty, tz):
        (pk, ppk), (pq, ppq) = _i(partner_kind), _d(partner_charge)
        (px, ppx), (py, ppy), (pz, ppz) = _d(px), _d(py), _d(pz)
        (tx, ptx), (ty, pty), (tz, ptz) = _d(tx), _d(ty), _d(tz)
        inter = np.zeros(len(tx))
        self.L.orc_particle_nonbonded(self.pp, int(kind_i), C.c_double(q_i), len(pk), ppk, ppq,
                                      ppx, ppy, ppz, len(tx), ptx, pty, ptz,
                                      inter.ctypes.data_as(_dp))
        return inter

    def calculate_torque(self, box_mols, mol_start, x, y, z, com, aF, rF, n_mols):
        (bm, pbm), (ms, pms) = _i(box_mols), _i(mol_start)
        arrs = [_d(a) for a in (x, y, z, *com, *aF, *rF)]
        t = [np.zeros(n_mols) for _ in range(3)]
        self.L.orc_calculate_torque(self.pp, len(bm), pbm, pms, *[a[1] for a in arrs],
                                    *[a.ctypes.data_as(_dp) for a in t])
        return t

    def energy_lrc(self, mol_kind_start, mol_kind_atom_kinds, num_kind_in_box):
        (a, pa), (b, pb), (c, pc) = _i(mol_kind_start), _i(mol_kind_atom_kinds), _i(num_kind_in_box)
        return self.L.orc_energy_lrc(self.pp, len(a) - 1, pa, pb, pc)

    # ---- reciprocal path -------------------------------------------------
    def recip_init_orth(self):
        """RecipInit: the orthogonal or non-orthogonal enumeration, as Ewald::RecipInit
        dispatches (src/Ewald.cpp:644-652)."""
        kmax = C.c_int()
        fn = self.L.orc_recip_init_nonorth if self.p.nonOrth else self.L.orc_recip_init_orth
        nk = fn(self.pp, None, None, None, None, None, C.byref(kmax))
        arr = [np.zeros(nk) for _ in range(5)]
        fn(self.pp, *[a.ctypes.data_as(_dp) for a in arr], C.byref(kmax))
        return (*arr, kmax.value)   # kx, ky, kz, hsqr, prefact, kmax

    def box_recip_sums(self, box_mols, mol_start, x, y, z, charge, kx, ky, kz, k0=0, k1=None):
        (bm, pbm), (ms, pms) = _i(box_mols), _i(mol_start)
        (x, px), (y, py), (z, pz), (q, pq) = _d(x), _d(y), _d(z), _d(charge)
        (kx, pkx), (ky, pky), (kz, pkz) = _d(kx), _d(ky), _d(kz)
        nk = len(kx)
        k1 = nk if k1 is None else k1
        sR, sI = np.zeros(nk), np.zeros(nk)
        self.L.orc_box_recip_sums_slab(len(bm), pbm, pms, px, py, pz, pq, int(k0), int(k1),
                                       pkx, pky, pkz, sR.ctypes.data_as(_dp),
                                       sI.ctypes.data_as(_dp))
        return sR, sI

    def box_reciprocal(self, sR, sI, prefact):
        (sR, a), (sI, b), (pf, c) = _d(sR), _d(sI), _d(prefact)
        return self.L.orc_box_reciprocal(len(sR), a, b, c)

    def mol_reciprocal(self, q, old, new, kx, ky, kz, prefact, sRref, sIref, lambda_coef=1.0):
        (q, pq) = _d(q)
        o = [_d(a) for a in old]
        nw = [_d(a) for a in new]
        ks = [_d(a) for a in (kx, ky, kz, prefact, sRref, sIref)]
        nk = len(ks[0][0])
        sRn, sIn = np.zeros(nk), np.zeros(nk)
        e = self.L.orc_mol_reciprocal_l(len(q), pq, *[a[1] for a in o], *[a[1] for a in nw], nk,
                                        *[a[1] for a in ks], sRn.ctypes.data_as(_dp),
                                        sIn.ctypes.data_as(_dp), C.c_double(lambda_coef))
        return e, sRn, sIn

    def swap_recip(self, insert, q, mxyz, kx, ky, kz, prefact, sRref, sIref):
        (q, pq) = _d(q)
        m = [_d(a) for a in mxyz]
        ks = [_d(a) for a in (kx, ky, kz, prefact, sRref, sIref)]
        nk = len(ks[0][0])
        sRn, sIn = np.zeros(nk), np.zeros(nk)
        e = self.L.orc_swap_recip(int(insert), len(q), pq, *[a[1] for a in m], nk,
                                  *[a[1] for a in ks], sRn.ctypes.data_as(_dp),
                                  sIn.ctypes.data_as(_dp))
        return e, sRn, sIn

    def recip_weighted(self, w, xyz, kx, ky, kz, prefact, baseR, baseI, scale=1.0):
        (w, pw) = _d(w)
        m = [_d(a) for a in xyz]
        ks = [_d(a) for a in (kx, ky, kz, prefact, baseR, baseI)]
        nk = len(ks[0][0])
        sRn, sIn = np.zeros(nk), np.zeros(nk)
        e = self.L.orc_recip_weighted(len(w), pw, *[a[1] for a in m], nk, *[a[1] for a in ks],
                                      float(scale), sRn.ctypes.data_as(_dp),
                                      sIn.ctypes.data_as(_dp))
        return e, sRn, sIn

    def change_recip(self, q, mxyz, kx, ky, kz, prefact, sRref, sIref, lambda_coul, i_state):
        (q, pq) = _d(q)
        m = [_d(a) for a in mxyz]
        ks = [_d(a) for a in (kx, ky, kz, prefact, sRref, sIref)]
        (lam, pl) = _d(lambda_coul)
        out = np.zeros(len(lam))
        self.L.orc_change_recip(len(q), pq, *[a[1] for a in m], len(ks[0][0]),
                                *[a[1] for a in ks], len(lam), pl, int(i_state),
                                out.ctypes.data_as(_dp))
        return out

    def box_force_reciprocal(self, box_mols, mol_start, x, y, z, charge, kx, ky, kz, prefact,
                             sR, sI, n_mols):
        (bm, pbm), (ms, pms) = _i(box_mols), _i(mol_start)
        arrs = [_d(a) for a in (x, y, z, charge)]
        ks = [_d(a) for a in (kx, ky, kz, prefact, sR, sI)]
        rF = [np.zeros(len(arrs[0][0])) for _ in range(3)]
        mF = [np.zeros(n_mols) for _ in range(3)]
        self.L.orc_box_force_reciprocal(self.pp, len(bm), pbm, pms, *[a[1] for a in arrs],
                                        len(ks[0][0]), *[a[1] for a in ks],
                                        *[a.ctypes.data_as(_dp) for a in rF],
                                        *[a.ctypes.data_as(_dp) for a in mF])
        return rF, mF

    def box_correction(self, box_mols, mol_start, x, y, z, charge):
        (bm, pbm), (ms, pms) = _i(box_mols), _i(mol_start)
        arrs = [_d(a) for a in (x, y, z, charge)]
        return self.L.orc_box_correction(self.pp, len(bm), pbm, pms, *[a[1] for a in arrs])

    def box_self(self, box_mols, mol_start, charge):
        (bm, pbm), (ms, pms), (q, pq) = _i(box_mols), _i(mol_start), _d(charge)
        return self.L.orc_box_self(self.pp, len(bm), pbm, pms, pq)

    def swap_correction(self, q, mxyz):
        (q, pq) = _d(q)
        m = [_d(a) for a in mxyz]
        return self.L.orc_swap_correction(self.pp, len(q), pq, *[a[1] for a in m])

    def change_self_correction(self, q, mxyz):
        (q, pq) = _d(q)
        m = [_d(a) for a in mxyz]
        es, ec = C.c_double(), C.c_double()
        self.L.orc_change_self_correction(self.pp, len(q), pq, *[a[1] for a in m],
                                          C.byref(es), C.byref(ec))
        return es.value, ec.value

    def swap_self(self, q):
        (q, pq) = _d(q)
        return self.L.orc_swap_self(self.pp, len(q), pq)


def max_threads():
    return lib().orc_max_threads()


def set_threads(n):
    lib().orc_set_threads(int(n))


# --------------------------------------------------------------------------
def read_dump(path):
    """Read a GOMCDUMP file written by oracle/ref_probe.cpp -> dict of arrays
    (scalars are unwrapped to python numbers)."""
    out = {}
    with open(path, "rb") as f:
        assert f.read(8) == b"GOMCDUMP", "bad magic"
        while True:
            h = f.read(4)
            if not h:
                break
            (nl,) = struct.unpack("<I", h)
            name = f.read(nl).decode()
            (dt,) = struct.unpack("<B", f.read(1))
            (cnt,) = struct.unpack("<Q", f.read(8))
            dtype = np.float64 if dt == 0 else np.int32
            a = np.frombuffer(f.read(cnt * np.dtype(dtype).itemsize), dtype=dtype).copy()
            out[name] = a
    return out


def sc(d, key):
    """Scalar entry of a dump."""
    return d[key].reshape(-1)[0].item()
