"""ctypes binding of the C ABI in include/gomc_b200.h (libgomc_b200.so).

This is the Python face of the product path: it only marshals numpy arrays to
the C entry points.  There is no CPU fallback -- if the CUDA library has not
been built, or no B200 is present, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgomc_b200.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p

ATOM_FORCE, MOL_FORCE, ATOM_FORCE_REC, MOL_FORCE_REC, MOL_TORQUE = range(5)
K_NEW, K_REF = 0, 1
K_DEVICE = 2          # OR-ed into K_NEW / K_REF: read back the device-resident copy
SUM_NEW, SUM_REF = 0, 1

# name -> (restype, argtypes); mirrors include/gomc_b200.h one to one
_SIGS = {
    "gomcb200_last_error": (C.c_char_p, []),
    "gomcb200_version": (C.c_int, []),
    "gomcb200_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int]),
    "gomcb200_destroy": (C.c_int, [_vp]),
    "gomcb200_launch_count": (C.c_longlong, [_vp]),
    "gomcb200_init_forcefield": (C.c_int, [_vp, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int,
                                           C.c_double, _dp, C.c_double, C.c_double, _dp,
                                           C.c_int, C.c_int, C.c_double]),
    "gomcb200_init_exp6": (C.c_int, [_vp, _dp, _dp, _dp, C.c_int]),
    "gomcb200_init_topology": (C.c_int, [_vp, C.c_int, C.c_int, _ip, _ip, _dp, _ip]),
    "gomcb200_set_box_molecules": (C.c_int, [_vp, C.c_int, _ip, C.c_int]),
    "gomcb200_set_box_axes": (C.c_int, [_vp, C.c_int, _dp]),
    "gomcb200_set_box_cell_basis": (C.c_int, [_vp, C.c_int, _dp, _dp, _dp]),
    "gomcb200_set_coords": (C.c_int, [_vp, _dp, _dp, _dp, C.c_int, C.c_int]),
    "gomcb200_get_coords": (C.c_int, [_vp, _dp, _dp, _dp, C.c_int, C.c_int]),
    "gomcb200_set_com": (C.c_int, [_vp, _dp, _dp, _dp, C.c_int, C.c_int]),
    "gomcb200_get_com": (C.c_int, [_vp, _dp, _dp, _dp, C.c_int, C.c_int]),
    "gomcb200_set_molecule_coords": (C.c_int, [_vp, C.c_int, _dp, _dp, _dp, _dp]),
    "gomcb200_box_inter": (C.c_int, [_vp, C.c_int, _dp, _dp]),
    "gomcb200_box_force": (C.c_int, [_vp, C.c_int, _dp, _dp]),
    "gomcb200_molecule_inter": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _ip]),
    "gomcb200_molecule_trial": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _ip, _dp]),
    "gomcb200_init_softcore": (C.c_int, [_vp, C.c_double, C.c_double, C.c_int, C.c_int]),
    "gomcb200_update_lambda": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                         C.c_int]),
    "gomcb200_mp_transform": (C.c_int, [_vp, C.c_int, C.c_int, C.c_double, C.c_double,
                                        C.c_ulonglong, C.c_uint, C.c_ulonglong, C.c_void_p]),
    "gomcb200_bm_transform": (C.c_int, [_vp, C.c_int, C.c_int, C.c_double, C.c_double,
                                        C.c_ulonglong, C.c_uint, C.c_ulonglong, C.c_void_p]),
    "gomcb200_bm_coeff": (C.c_int, [_vp, C.c_int, C.c_int, C.c_double, C.c_double, _dp]),
    "gomcb200_mp_get_trial": (C.c_int, [_vp, _dp, _dp, _dp, _ip]),
    "gomcb200_mp_select": (C.c_int, [_vp, C.c_int]),
    "gomcb200_mp_coeff": (C.c_int, [_vp, C.c_int, C.c_int, C.c_double, C.c_double, _dp]),
    "gomcb200_mp_accept": (C.c_int, [_vp, C.c_int]),
    "gomcb200_box_inter_virial": (C.c_int, [_vp, C.c_int, _dp, _dp]),
    "gomcb200_virial_reciprocal": (C.c_int, [_vp, C.c_int, _dp]),
    "gomcb200_mol_exchange_reciprocal": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, _dp, _dp,
                                                   C.c_int, C.c_double, _dp]),
    "gomcb200_change_lambda_mol_reciprocal": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, _dp,
                                                        C.c_double, _dp]),
    "gomcb200_change_recip": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _dp, C.c_int, _dp]),
    "gomcb200_change_self_correction": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp]),
    "gomcb200_swap_correction": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp]),
    "gomcb200_swap_trial": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, _dp, _dp, _dp]),
    "gomcb200_particle_inter": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp,
                                          _dp, _dp, _dp, _ip]),
    "gomcb200_calculate_torque": (C.c_int, [_vp, C.c_int]),
    "gomcb200_get_forces": (C.c_int, [_vp, C.c_int, _dp, _dp, _dp, C.c_int, C.c_int]),
    "gomcb200_init_ewald": (C.c_int, [_vp, C.c_int, _dp]),
    "gomcb200_recip_init": (C.c_int, [_vp, C.c_int, _dp, _ip, _ip]),
    "gomcb200_recip_init_volume": (C.c_int, [_vp, C.c_int, _dp, C.c_double, _ip, _ip]),
    "gomcb200_recip_count": (C.c_int, [_vp, C.c_int, _dp, C.c_double, _ip]),
    "gomcb200_get_kvectors": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, C.c_int]),
    "gomcb200_box_reciprocal_setup": (C.c_int, [_vp, C.c_int, _dp]),
    "gomcb200_box_reciprocal_sums": (C.c_int, [_vp, C.c_int, _dp]),
    "gomcb200_box_reciprocal": (C.c_int, [_vp, C.c_int, C.c_int, _dp]),
    "gomcb200_mol_reciprocal": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, _dp, _dp]),
    "gomcb200_swap_reciprocal": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, _dp]),
    "gomcb200_box_force_reciprocal": (C.c_int, [_vp, C.c_int]),
    "gomcb200_get_recip_sums": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, C.c_int]),
    "gomcb200_set_recip_ref": (C.c_int, [_vp, C.c_int]),
    "gomcb200_copy_recip": (C.c_int, [_vp, C.c_int]),
    "gomcb200_update_recip": (C.c_int, [_vp, C.c_int]),
    "gomcb200_update_recip_vec": (C.c_int, [_vp, C.c_int]),
    "gomcb200_box_self_correction": (C.c_int, [_vp, C.c_int, _dp, _dp]),
    "gomcb200_call_box_inter": (C.c_int, [_vp, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp]),
    "gomcb200_call_box_reciprocal_sums": (C.c_int, [_vp, C.c_int, _dp, _dp, _dp, _dp]),
    "gomcb200_call_box_force": (C.c_int, [_vp, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp,
                                         _dp, _dp, _dp, _dp, _dp, _dp]),
    "gomcb200_call_full_box_energy": (C.c_int, [_vp, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp]),
    "gomcb200_set_kvectors": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp]),
    "gomcb200_call_box_reciprocal_points": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _dp, _dp,
                                                      _dp, _dp, _dp, _dp, _dp]),
    "gomcb200_call_mol_reciprocal": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp,
                                               _dp, _dp, _dp, _dp, _dp]),
    "gomcb200_call_swap_reciprocal": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, _dp, _dp,
                                                C.c_int, _dp, _dp, _dp]),
    "gomcb200_set_recip_sums": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, C.c_int]),
    "gomcb200_set_forces": (C.c_int, [_vp, C.c_int, _dp, _dp, _dp, C.c_int, C.c_int]),
    "gomcb200_set_shard": (C.c_int, [_vp, C.c_int, C.c_int]),
    "gomcb200_mark_coords_changed": (C.c_int, [_vp]),
    "gomcb200_set_recip_algo": (C.c_int, [_vp, C.c_int]),
    "gomcb200_set_pair_algo": (C.c_int, [_vp, C.c_int]),
    "gomcb200_particle_nonbonded": (C.c_int, [_vp, C.c_int, C.c_int, C.c_double, C.c_int, _ip,
                                              _dp, _dp, _dp, _dp, C.c_int, _dp, _dp, _dp, _dp]),
    "gomcb200_comm_unique_id": (C.c_int, [_vp]),
    "gomcb200_set_comm": (C.c_int, [_vp, _vp, C.c_int, C.c_int]),
    "gomcb200_set_recip_auto_work": (C.c_int, [_vp, C.c_double]),
    "gomcb200_last_timing": (C.c_int, [_vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "gomcb200_enable_timing": (C.c_int, [_vp, C.c_int]),
}
EXPORTED_SYMBOLS = tuple(_SIGS)

_lib = None


def comm_unique_id():
    """128-byte NCCL unique id for Engine.set_comm (create on one rank, share with all)."""
    L = load_library()
    buf = C.create_string_buffer(128)
    if L.gomcb200_comm_unique_id(buf):
        raise RuntimeError(L.gomcb200_last_error().decode())
    return buf.raw


def load_library():
    """dlopen libgomc_b200.so and type every symbol of include/gomc_b200.h."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ "
                "as g; g.build()'` (nvcc, sm_100a).  gomc_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)     # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class EngineError(RuntimeError):
    pass


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_ip)


class Engine:
    """Thin object wrapper over the opaque gomcb200_engine handle."""

    def __init__(self, n_boxes=1, device=-1):
        self.L = load_library()
        h = _vp()
        rc = self.L.gomcb200_create(C.byref(h), device, n_boxes)
        if rc != 0:
            raise EngineError(f"gomcb200_create failed ({rc}): "
                              f"{self.L.gomcb200_last_error().decode()}")
        self.h = h
        self.n_boxes = n_boxes
        self.n_atoms = 0
        self.n_mols = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.gomcb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise EngineError(f"gomc_b200 error {rc}: {self.L.gomcb200_last_error().decode()}")

    # ---- setup -----------------------------------------------------------
    def init_forcefield(self, sigma_sq, epsilon_cn, n, vdw_kind, count, r_cut, r_cut_coulomb,
                        r_cut_low, r_switch, alpha, ewald, electrostatic, is_martini=0,
                        dielectric=1.0):
        (a, pa), (b, pb), (c, pc) = _d(sigma_sq), _d(epsilon_cn), _d(n)
        (rc_, prc), (al, pal) = _d(np.atleast_1d(r_cut_coulomb)), _d(np.atleast_1d(alpha))
        self._ck(self.L.gomcb200_init_forcefield(self.h, pa, pb, pc, int(vdw_kind),
                                                 int(is_martini), int(count), float(r_cut), prc,
                                                 float(r_cut_low), float(r_switch), pal,
                                                 int(ewald), int(electrostatic),
                                                 1.0 / float(dielectric)))

    def init_exp6(self, r_min, exp_const, r_max_sq):
        (a, pa), (b, pb), (c, pc) = _d(r_min), _d(exp_const), _d(r_max_sq)
        self._ck(self.L.gomcb200_init_exp6(self.h, pa, pb, pc, len(a)))

    def init_topology(self, kind, mol, charge, mol_start):
        (k, pk), (m, pm), (q, pq), (s, ps) = _i(kind), _i(mol), _d(charge), _i(mol_start)
        self.n_atoms, self.n_mols = len(k), len(s) - 1
        self.mol_start = s.copy()
        self._ck(self.L.gomcb200_init_topology(self.h, len(k), len(s) - 1, pk, pm, pq, ps))

    def set_box_molecules(self, box, mols):
        m, pm = _i(mols)
        self._ck(self.L.gomcb200_set_box_molecules(self.h, box, pm, len(m)))

    def set_box_axes(self, box, axis):
        a, pa = _d(axis)
        self._ck(self.L.gomcb200_set_box_axes(self.h, box, pa))

    def set_box_cell_basis(self, box, cell_basis, cell_basis_inv, axis):
        (b, pb), (bi, pbi), (a, pa) = _d(np.reshape(cell_basis, -1)), \
            _d(np.reshape(cell_basis_inv, -1)), _d(axis)
        self._ck(self.L.gomcb200_set_box_cell_basis(self.h, box, pb, pbi, pa))

    def set_coords(self, x, y, z, first=0):
        (x, px), (y, py), (z, pz) = _d(x), _d(y), _d(z)
        self._ck(self.L.gomcb200_set_coords(self.h, px, py, pz, first, len(x)))

    def get_coords(self, first=0, count=None):
        count = self.n_atoms - first if count is None else count
        out = [np.zeros(count) for _ in range(3)]
        self._ck(self.L.gomcb200_get_coords(self.h, *[o.ctypes.data_as(_dp) for o in out],
                                            first, count))
        return out

    def get_com(self, first=0, count=None):
        count = self.n_mols - first if count is None else count
        out = [np.zeros(count) for _ in range(3)]
        self._ck(self.L.gomcb200_get_com(self.h, *[o.ctypes.data_as(_dp) for o in out],
                                         first, count))
        return out

    def set_com(self, x, y, z, first=0):
        (x, px), (y, py), (z, pz) = _d(x), _d(y), _d(z)
        self._ck(self.L.gomcb200_set_com(self.h, px, py, pz, first, len(x)))

    def set_molecule_coords(self, mol_index, x, y, z, com=None):
        (x, px), (y, py), (z, pz) = _d(x), _d(y), _d(z)
        pc = None
        if com is not None:
            com, pc = _d(com)
        self._ck(self.L.gomcb200_set_molecule_coords(self.h, mol_index, px, py, pz, pc))

    # ---- pair path -------------------------------------------------------
    def box_inter(self, box=0):
        lj, re = C.c_double(), C.c_double()
        self._ck(self.L.gomcb200_box_inter(self.h, box, C.byref(lj), C.byref(re)))
        return lj.value, re.value

    def box_force(self, box=0):
        lj, re = C.c_double(), C.c_double()
        self._ck(self.L.gomcb200_box_force(self.h, box, C.byref(lj), C.byref(re)))
        return lj.value, re.value

    def molecule_inter(self, box, mol_index, nx, ny, nz):
        (nx, px), (ny, py), (nz, pz) = _d(nx), _d(ny), _d(nz)
        lj, re, ov = C.c_double(), C.c_double(), C.c_int()
        self._ck(self.L.gomcb200_molecule_inter(self.h, box, mol_index, px, py, pz,
                                                C.byref(lj), C.byref(re), C.byref(ov)))
        return lj.value, re.value, bool(ov.value)

    def molecule_trial(self, box, mol_index, nx, ny, nz):
        (nx, px), (ny, py), (nz, pz) = _d(nx), _d(ny), _d(nz)
        lj, re, ov, er = C.c_double(), C.c_double(), C.c_int(), C.c_double()
        self._ck(self.L.gomcb200_molecule_trial(self.h, box, mol_index, px, py, pz, C.byref(lj),
                                                C.byref(re), C.byref(ov), C.byref(er)))
        return lj.value, re.value, bool(ov.value), er.value

    def init_softcore(self, sc_alpha, sc_sigma_6, sc_power, sc_coul):
        self._ck(self.L.gomcb200_init_softcore(self.h, float(sc_alpha), float(sc_sigma_6),
                                               int(sc_power), int(sc_coul)))

    def update_lambda(self, box, mol_index, mol_kind, lambda_vdw, lambda_coulomb,
                      is_fraction=True):
        self._ck(self.L.gomcb200_update_lambda(self.h, box, int(mol_index), int(mol_kind),
                                               float(lambda_vdw), float(lambda_coulomb),
                                               int(is_fraction)))

    def mp_transform(self, box, move_type, vmax, lambda_beta, step, key, seed, involved=None):
        ptr = None
        if involved is not None:
            involved = np.ascontiguousarray(involved, dtype=np.int8)
            ptr = involved.ctypes.data_as(C.c_void_p)
        self._ck(self.L.gomcb200_mp_transform(self.h, box, int(move_type), float(vmax),
                                              float(lambda_beta), int(step), int(key),
                                              int(seed), ptr))

    def bm_transform(self, box, move_type, vmax, beta, step, key, seed):
        self._ck(self.L.gomcb200_bm_transform(self.h, box, int(move_type), float(vmax),
                                              float(beta), int(step), int(key), int(seed), None))

    def bm_coeff(self, box, move_type, vmax, beta):
        w = C.c_double()
        self._ck(self.L.gomcb200_bm_coeff(self.h, box, int(move_type), float(vmax), float(beta),
                                          C.byref(w)))
        return w.value

    def mp_get_trial(self, n_mols):
        k = [np.zeros(n_mols) for _ in range(3)]
        inr = np.zeros(n_mols, dtype=np.int32)
        self._ck(self.L.gomcb200_mp_get_trial(self.h, *[a.ctypes.data_as(_dp) for a in k],
                                              inr.ctypes.data_as(_ip)))
        return k, inr

    def mp_select(self, trial):
        self._ck(self.L.gomcb200_mp_select(self.h, int(trial)))

    def mp_coeff(self, box, move_type, vmax, lambda_beta):
        w = C.c_double()
        self._ck(self.L.gomcb200_mp_coeff(self.h, box, int(move_type), float(vmax),
                                          float(lambda_beta), C.byref(w)))
        return w.value

    def mp_accept(self, box=0):
        self._ck(self.L.gomcb200_mp_accept(self.h, box))

    def box_inter_virial(self, box=0):
        vT, rT = np.zeros(3), np.zeros(3)
        self._ck(self.L.gomcb200_box_inter_virial(self.h, box, vT.ctypes.data_as(_dp),
                                                  rT.ctypes.data_as(_dp)))
        return vT, rT

    def virial_reciprocal(self, box=0):
        wT = np.zeros(3)
        self._ck(self.L.gomcb200_virial_reciprocal(self.h, box, wT.ctypes.data_as(_dp)))
        return wT

    def mol_exchange_reciprocal(self, box, w, x, y, z, first_call=True, scale=1.0):
        (w, pw), (x, px), (y, py), (z, pz) = _d(w), _d(x), _d(y), _d(z)
        en = C.c_double()
        self._ck(self.L.gomcb200_mol_exchange_reciprocal(self.h, box, len(w), pw, px, py, pz,
                                                         int(first_call), float(scale),
                                                         C.byref(en)))
        return en.value

    def change_lambda_mol_reciprocal(self, box, mol_index, x, y, z, lambda_coef):
        (x, px), (y, py), (z, pz) = _d(x), _d(y), _d(z)
        en = C.c_double()
        self._ck(self.L.gomcb200_change_lambda_mol_reciprocal(self.h, box, mol_index, px, py, pz,
                                                              float(lambda_coef), C.byref(en)))
        return en.value

    def change_recip(self, box, mol_index, lambda_coul, i_state):
        (lam, pl) = _d(lambda_coul)
        out = np.zeros(len(lam))
        self._ck(self.L.gomcb200_change_recip(self.h, box, mol_index, len(lam), pl, int(i_state),
                                              out.ctypes.data_as(_dp)))
        return out

    def change_self_correction(self, box, mol_index):
        es, ec = C.c_double(), C.c_double()
        self._ck(self.L.gomcb200_change_self_correction(self.h, box, mol_index, C.byref(es),
                                                        C.byref(ec)))
        return es.value, ec.value

    def swap_correction(self, box, mol_index, x, y, z):
        (x, px), (y, py), (z, pz) = _d(x), _d(y), _d(z)
        co, se = C.c_double(), C.c_double()
        self._ck(self.L.gomcb200_swap_correction(self.h, box, mol_index, px, py, pz, C.byref(co),
                                                 C.byref(se)))
        return co.value, se.value

    def swap_trial(self, box, mol_index, x, y, z, insert):
        (x, px), (y, py), (z, pz) = _d(x), _d(y), _d(z)
        en, co, se = C.c_double(), C.c_double(), C.c_double()
        self._ck(self.L.gomcb200_swap_trial(self.h, box, mol_index, px, py, pz, int(insert),
                                            C.byref(en), C.byref(co), C.byref(se)))
        return en.value, co.value, se.value

    def particle_inter(self, box, mol_index, part_index, tx, ty, tz):
        (tx, px), (ty, py), (tz, pz) = _d(tx), _d(ty), _d(tz)
        t = len(tx)
        en, re, ov = np.zeros(t), np.zeros(t), np.zeros(t, np.int32)
        self._ck(self.L.gomcb200_particle_inter(self.h, box, mol_index, part_index, t, px, py,
                                                pz, en.ctypes.data_as(_dp),
                                                re.ctypes.data_as(_dp), ov.ctypes.data_as(_ip)))
        return en, re, ov.astype(bool)

    def calculate_torque(self, box=0):
        self._ck(self.L.gomcb200_calculate_torque(self.h, box))

    def particle_nonbonded(self, box, kind_i, q_i, partner_kind, partner_charge, px, py, pz,
                           tx, ty, tz):
        pk = np.ascontiguousarray(partner_kind, dtype=np.int32)
        (pq, ppq), (px, ppx), (py, ppy), (pz, ppz) = _d(partner_charge), _d(px), _d(py), _d(pz)
        (tx, ptx), (ty, pty), (tz, ptz) = _d(tx), _d(ty), _d(tz)
        inter = np.zeros(len(tx))
        self._ck(self.L.gomcb200_particle_nonbonded(
            self.h, box, int(kind_i), float(q_i), len(pk), pk.ctypes.data_as(_ip), ppq, ppx, ppy,
            ppz, len(tx), ptx, pty, ptz, inter.ctypes.data_as(_dp)))
        return inter

    def get_forces(self, which, first=0, count=None):
        limit = self.n_atoms if which in (ATOM_FORCE, ATOM_FORCE_REC) else self.n_mols
        count = limit - first if count is None else count
        out = [np.zeros(count) for _ in range(3)]
        self._ck(self.L.gomcb200_get_forces(self.h, which, *[o.ctypes.data_as(_dp) for o in out],
                                            first, count))
        return out

    # ---- reciprocal path -------------------------------------------------
    def init_ewald(self, image_total, recip_rcut):
        r, pr = _d(np.atleast_1d(recip_rcut))
        self._ck(self.L.gomcb200_init_ewald(self.h, int(image_total), pr))

    def recip_init(self, box, axis, volume=0.0):
        a, pa = _d(axis)
        n, kmax = C.c_int(), C.c_int()
        self._ck(self.L.gomcb200_recip_init_volume(self.h, box, pa, float(volume), C.byref(n),
                                                   C.byref(kmax)))
        return n.value, kmax.value

    def recip_count(self, box, axis, excess=1.0):
        a, pa = _d(axis)
        n = C.c_int()
        self._ck(self.L.gomcb200_recip_count(self.h, box, pa, float(excess), C.byref(n)))
        return n.value

    def get_kvectors(self, box, which, n):
        out = [np.zeros(n) for _ in range(5)]
        self._ck(self.L.gomcb200_get_kvectors(self.h, box, which,
                                              *[o.ctypes.data_as(_dp) for o in out], n))
        return out  # kx, ky, kz, hsqr, prefact

    def box_reciprocal_setup(self, box=0):
        e = C.c_double()
        self._ck(self.L.gomcb200_box_reciprocal_setup(self.h, box, C.byref(e)))
        return e.value

    def box_reciprocal_sums(self, box=0):
        e = C.c_double()
        self._ck(self.L.gomcb200_box_reciprocal_sums(self.h, box, C.byref(e)))
        return e.value

    def box_reciprocal(self, box=0, is_new_volume=False):
        e = C.c_double()
        self._ck(self.L.gomcb200_box_reciprocal(self.h, box, int(is_new_volume), C.byref(e)))
        return e.value

    def mol_reciprocal(self, box, mol_index, nx, ny, nz):
        (nx, px), (ny, py), (nz, pz) = _d(nx), _d(ny), _d(nz)
        e = C.c_double()
        self._ck(self.L.gomcb200_mol_reciprocal(self.h, box, mol_index, px, py, pz, C.byref(e)))
        return e.value

    def swap_reciprocal(self, box, mol_index, x, y, z, insert):
        (x, px), (y, py), (z, pz) = _d(x), _d(y), _d(z)
        e = C.c_double()
        self._ck(self.L.gomcb200_swap_reciprocal(self.h, box, mol_index, px, py, pz,
                                                 int(insert), C.byref(e)))
        return e.value

    def box_force_reciprocal(self, box=0):
        self._ck(self.L.gomcb200_box_force_reciprocal(self.h, box))

    # ---- literal drop-ins of the reciprocal seam (host arrays in, host sums out) ----
    def set_kvectors(self, box, kx, ky, kz, hsqr, prefact):
        arrs = [_d(a) for a in (kx, ky, kz, hsqr, prefact)]
        self._ck(self.L.gomcb200_set_kvectors(self.h, box, len(arrs[0][0]),
                                              *[p for _, p in arrs]))

    def call_box_reciprocal_points(self, box, new_set, x, y, z, q, nk):
        (x, px), (y, py), (z, pz), (q, pq) = _d(x), _d(y), _d(z), _d(q)
        r, i, e = np.zeros(nk), np.zeros(nk), C.c_double()
        self._ck(self.L.gomcb200_call_box_reciprocal_points(
            self.h, box, int(new_set), len(x), px, py, pz, pq, r.ctypes.data_as(_dp),
            i.ctypes.data_as(_dp), C.byref(e)))
        return e.value, r, i

    def call_mol_reciprocal(self, box, q, old, new, nk):
        arrs = [_d(a) for a in (q, *old, *new)]
        r, i, e = np.zeros(nk), np.zeros(nk), C.c_double()
        self._ck(self.L.gomcb200_call_mol_reciprocal(
            self.h, box, len(arrs[0][0]), *[p for _, p in arrs], r.ctypes.data_as(_dp),
            i.ctypes.data_as(_dp), C.byref(e)))
        return e.value, r, i

    def call_swap_reciprocal(self, box, q, xyz, insert, nk):
        arrs = [_d(a) for a in (q, *xyz)]
        r, i, e = np.zeros(nk), np.zeros(nk), C.c_double()
        self._ck(self.L.gomcb200_call_swap_reciprocal(
            self.h, box, len(arrs[0][0]), *[p for _, p in arrs], int(insert),
            r.ctypes.data_as(_dp), i.ctypes.data_as(_dp), C.byref(e)))
        return e.value, r, i

    def set_recip_sums(self, box, which, sum_r, sum_i):
        (r, pr), (i, pi) = _d(sum_r), _d(sum_i)
        self._ck(self.L.gomcb200_set_recip_sums(self.h, box, which, pr, pi, len(r)))

    def set_forces(self, which, x, y, z, first=0):
        (x, px), (y, py), (z, pz) = _d(x), _d(y), _d(z)
        self._ck(self.L.gomcb200_set_forces(self.h, which, px, py, pz, first, len(x)))

    def get_recip_sums(self, box, which, n):
        r, i = np.zeros(n), np.zeros(n)
        self._ck(self.L.gomcb200_get_recip_sums(self.h, box, which, r.ctypes.data_as(_dp),
                                                i.ctypes.data_as(_dp), n))
        return r, i

    def set_recip_ref(self, box=0):
        self._ck(self.L.gomcb200_set_recip_ref(self.h, box))

    def copy_recip(self, box=0):
        self._ck(self.L.gomcb200_copy_recip(self.h, box))

    def update_recip(self, box=0):
        self._ck(self.L.gomcb200_update_recip(self.h, box))

    def update_recip_vec(self, box=0):
        self._ck(self.L.gomcb200_update_recip_vec(self.h, box))

    def box_self_correction(self, box=0):
        s, c = C.c_double(), C.c_double()
        self._ck(self.L.gomcb200_box_self_correction(self.h, box, C.byref(s), C.byref(c)))
        return s.value, c.value

    # ---- host-buffer drop-ins -------------------------------------------
    def call_box_inter(self, box, x, y, z, axis):
        (x, px), (y, py), (z, pz), (a, pa) = _d(x), _d(y), _d(z), _d(axis)
        lj, re = C.c_double(), C.c_double()
        self._ck(self.L.gomcb200_call_box_inter(self.h, box, px, py, pz, pa, C.byref(re),
                                                C.byref(lj)))
        return lj.value, re.value

    def call_full_box_energy(self, box=0, x=None, y=None, z=None):
        """E1: BoxInter + BoxReciprocalSums + BoxReciprocal.  x/y/z host arrays
        (uploaded inside the call) or None to use the resident coordinates."""
        if x is None:
            px = py = pz = None
        else:
            (x, px), (y, py), (z, pz) = _d(x), _d(y), _d(z)
        lj, re, rc = C.c_double(), C.c_double(), C.c_double()
        self._ck(self.L.gomcb200_call_full_box_energy(self.h, box, px, py, pz, C.byref(lj),
                                                      C.byref(re), C.byref(rc)))
        return lj.value, re.value, rc.value

    def call_full_box_energy_ptr(self, box, px, py, pz, out):
        """Same with pre-extracted pointers (bench inner loop)."""
        return self.L.gomcb200_call_full_box_energy(self.h, box, px, py, pz, out[0], out[1],
                                                    out[2])

    # ---- misc --------------------------------------------------------------
    def set_shard(self, rank, world):
        self._ck(self.L.gomcb200_set_shard(self.h, int(rank), int(world)))

    def set_comm(self, unique_id, rank, world):
        """unique_id: the 128 bytes of comm_unique_id() of rank 0 (None for world == 1)."""
        buf = C.create_string_buffer(bytes(unique_id), 128) if unique_id is not None else None
        self._ck(self.L.gomcb200_set_comm(self.h, buf, int(rank), int(world)))

    def mark_coords_changed(self):
        self._ck(self.L.gomcb200_mark_coords_changed(self.h))

    def set_recip_algo(self, algo):
        self._ck(self.L.gomcb200_set_recip_algo(self.h, int(algo)))

    def set_pair_algo(self, algo):
        self._ck(self.L.gomcb200_set_pair_algo(self.h, int(algo)))

    def set_recip_auto_work(self, work):
        self._ck(self.L.gomcb200_set_recip_auto_work(self.h, float(work)))

    def enable_timing(self, on=True):
        self._ck(self.L.gomcb200_enable_timing(self.h, int(on)))

    def last_timing(self):
        a, b = C.c_float(), C.c_float()
        self._ck(self.L.gomcb200_last_timing(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def launch_count(self):
        return int(self.L.gomcb200_launch_count(self.h))

    # ---- convenience: bring up an engine for a synth.System ----------------
    @classmethod
    def from_system(cls, s, device=-1, recip=True):
        ff = s.ff
        sig, eps, nn = ff.tables()
        e = cls(1, device)
        e.init_forcefield(sig, eps, nn, ff.vdw_kind, len(ff.type_names), ff.r_cut,
                          [ff.r_cut_coulomb], ff.r_cut_low, ff.r_switch, [ff.alpha],
                          ff.ewald, ff.electrostatic, is_martini=ff.is_martini,
                          dielectric=ff.dielectric)
        if ff.vdw_kind == 3:
            e.init_exp6(*ff.exp6_tables())
        e.init_topology(s.kind, s.mol, s.charge, s.mol_start)
        e.set_box_molecules(0, np.arange(s.n_mols, dtype=np.int32))
        if getattr(s, "cell_basis", None) is not None:
            e.set_box_cell_basis(0, s.cell_basis, s.cell_basis_inv, s.axis)
        else:
            e.set_box_axes(0, s.axis)
        e.set_coords(s.x, s.y, s.z)
        e.set_com(*s.com())
        if recip and ff.ewald and ff.electrostatic:
            n = e.setup_ewald(s.axis, [ff.recip_rcut])
            e.nk = n
        else:
            e.nk = 0
        return e

    def setup_ewald(self, axis, recip_rcut, excess=1.0):
        """Ewald::Init sequence (src/Ewald.cpp:100-139): AllocMem, RecipInit,
        BoxReciprocalSetup, SetRecipRef."""
        self.init_ewald(0, recip_rcut)
        total = self.recip_count(0, axis, excess)
        self.init_ewald(total, recip_rcut)
        n, _ = self.recip_init(0, axis)
        self.box_reciprocal_setup(0)
        self.set_recip_ref(0)
        return n
