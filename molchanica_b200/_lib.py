"""ctypes binding of libmolchanica_md.so (include/molchanica_md.h).

The shared library is the product; this module only loads it.  There is no fallback of any
kind: a missing library, or a library that exports fewer symbols than the header declares,
raises immediately.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# MOLCHANICA_MD_LIB points the harness at another build of the same library (A/B kernel experiments)
LIB_PATH = os.environ.get("MOLCHANICA_MD_LIB") or os.path.join(_HERE, "libmolchanica_md.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "molchanica_md.h")

MC_OK = 0
MC_E_INVALID, MC_E_CUDA, MC_E_NODEVICE, MC_E_CAPACITY, MC_E_COMM = -1, -2, -3, -4, -5
COULOMB_NONE, COULOMB_PLAIN, COULOMB_ERFC = 0, 1, 2
FLAG_STATIC = 1


class McEnergy(C.Structure):
    _fields_ = [("energy_potential", C.c_double), ("energy_potential_nonbonded", C.c_double),
                ("energy_potential_bonded", C.c_double), ("energy_kinetic", C.c_double),
                ("temperature", C.c_double),
                ("energy_bond", C.c_double), ("energy_angle", C.c_double), ("energy_dihedral", C.c_double),
                ("volume", C.c_double), ("density", C.c_double), ("energy_pme", C.c_double)]


class McStats(C.Structure):
    _fields_ = [("n_atoms", C.c_int64), ("n_ghosts", C.c_int64), ("n_pairs_listed", C.c_int64),
                ("n_rebuilds", C.c_int64), ("n_steps", C.c_int64), ("n_kernel_launches", C.c_int64),
                ("n_cells", C.c_int64 * 3),
                ("pair_ms_sum", C.c_double), ("pair_launches_timed", C.c_int64),
                ("build_ms_sum", C.c_double), ("builds_timed", C.c_int64),
                ("integrate_ms_sum", C.c_double), ("integrate_launches_timed", C.c_int64),
                ("halo_ms_sum", C.c_double), ("halos_timed", C.c_int64), ("n_list_violations", C.c_int64),
                ("list_bytes", C.c_int64), ("ext_upload_bytes", C.c_int64)]


def declared_symbols():
    """Every function name include/molchanica_md.h declares."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mc_[a-z0-9_]+)\s*\(", text)))


def build(verbose=False):
    """Compile the library in-tree with nvcc for sm_100a (no GPU needed)."""
    r = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j8"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libmolchanica_md.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout[-2000:])
    return LIB_PATH


_LIB = None


def lib():
    """Load the library (once).  Raises if it is missing or incomplete -- never falls back."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU or PyTorch fallback for this engine)")
    L = C.CDLL(LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    if missing:
        raise RuntimeError(f"{LIB_PATH} does not export {missing}; rebuild it")
    vp, i64, i32, f32 = C.c_void_p, C.c_int64, C.c_int, C.c_float
    se, ss = C.c_int(0), C.c_int(0)
    L.mc_struct_sizes(C.byref(se), C.byref(ss))
    if (se.value, ss.value) != (C.sizeof(McEnergy), C.sizeof(McStats)):
        raise RuntimeError(f"{LIB_PATH}: mc_energy / mc_stats are {se.value} / {ss.value} bytes, the ctypes mirrors "
                           f"{C.sizeof(McEnergy)} / {C.sizeof(McStats)}: molchanica_b200/_lib.py is out of step with the header")
    L.mc_last_error.restype = C.c_char_p
    L.mc_last_error.argtypes = [vp]
    L.mc_create.argtypes = [i32, C.POINTER(vp)]
    L.mc_destroy.argtypes = [vp]
    L.mc_set_box.argtypes = [vp, vp, vp, i32]
    L.mc_set_atoms.argtypes = [vp, i64, vp, vp, vp, vp]
    L.mc_set_lj_table.argtypes = [vp, i32, vp]
    L.mc_set_exclusions.argtypes = [vp, vp, vp]
    L.mc_set_pairs14.argtypes = [vp, i64, vp, f32, f32]
    L.mc_set_bonds.argtypes = [vp, i64, vp, vp]
    L.mc_set_angles.argtypes = [vp, i64, vp, vp]
    L.mc_set_dihedrals.argtypes = [vp, i64, vp, vp]
    L.mc_set_molecule_ids.argtypes = [vp, vp]
    L.mc_get_energy_between_mols.argtypes = [vp, C.POINTER(C.c_double)]
    L.mc_get_pressure.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.mc_set_barostat.argtypes = [vp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_uint64]
    L.mc_get_box.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.mc_set_hbond_constraints.argtypes = [vp, i64, vp, vp]
    L.mc_set_virtual_sites.argtypes = [vp, i64, vp, f32, f32]
    L.mc_set_thermostat.argtypes = [vp, i32, f32, f32, C.c_uint64]
    L.mc_set_pme.argtypes = [vp, i32, i32, i32]
    L.mc_pme_suggest.argtypes = [f32, f32, vp, C.POINTER(f32), vp]
    L.mc_set_rigid_waters.argtypes = [vp, i64, vp, f32, f32, f32, f32]
    L.mc_set_cutoffs.argtypes = [vp, f32, f32, f32, i32, f32]
    L.mc_set_overrides.argtypes = [vp, i32, i32]
    L.mc_set_option.argtypes = [vp, C.c_char_p, C.c_double]
    L.mc_set_positions.argtypes = [vp, vp]
    L.mc_set_velocities.argtypes = [vp, vp]
    L.mc_build_neighbors.argtypes = [vp]
    L.mc_compute_forces.argtypes = [vp]
    L.mc_step.argtypes = [vp, f32, i32, vp]
    L.mc_minimize_energy.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    for f in ("mc_get_positions", "mc_get_velocities", "mc_get_forces", "mc_get_positions_global", "mc_get_forces_global"):
        getattr(L, f).argtypes = [vp, vp]
    L.mc_get_energy.argtypes = [vp, C.POINTER(McEnergy)]
    L.mc_get_stats.argtypes = [vp, C.POINTER(McStats)]
    L.mc_reset_timers.argtypes = [vp]
    L.mc_get_neighbors.argtypes = [vp, vp, vp, i64, C.POINTER(i64)]
    L.mc_time_kernels.argtypes = [vp, i32, i32]
    L.mc_last_pair_kernel_ms.restype = C.c_double
    L.mc_last_pair_kernel_ms.argtypes = [vp]
    L.mc_last_step_ms.restype = C.c_double
    L.mc_last_step_ms.argtypes = [vp]
    L.mc_last_dock_kernel_ms.restype = C.c_double
    L.mc_last_dock_kernel_ms.argtypes = [vp]
    L.mc_dock_score.argtypes = [vp, i64, vp, vp, vp, i64, vp, vp, vp, vp, i32, i32, vp, i64, vp, vp]
    L.mc_dock_score_flex.argtypes = [vp, i64, vp, vp, vp, i64, vp, vp, vp, vp, i32, i32, vp, i32, vp, vp, i64, vp, vp]
    L.mc_dock_make_poses_flex.argtypes = [vp, C.c_double, i32, i32, i32, i32, vp, i64, C.POINTER(i64)]
    L.mc_dock_flex_masks.argtypes = [i64, i64, vp, i32, vp, vp, vp]
    L.mc_dock_make_poses.argtypes = [vp, C.c_double, i32, i32, vp, i64, C.POINTER(i64)]
    L.mc_dock_orientation_count.argtypes = [i32]
    L.mc_dock_near_site.argtypes = [i64, vp, vp, vp, C.c_double, vp, C.POINTER(i64)]
    L.mc_dock_filter_poses.argtypes = [i64, vp, vp, i64, vp, vp, vp, f32, i64, vp, vp, C.POINTER(i64)]
    L.mc_dock_filter_poses_gpu.argtypes = [vp, i64, vp, vp, i64, vp, vp, vp, f32, i64, vp, vp, C.POINTER(i64)]
    L.mc_comm_unique_id.argtypes = [vp]
    L.mc_comm_init.argtypes = [vp, vp, i32, i32]
    L.mc_comm_counts.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    L.mc_dd_plan.argtypes = [vp, f32, i32, i32, vp]
    L.mc_snapshot_begin.argtypes = [vp, vp, vp, C.POINTER(i64)]
    L.mc_snapshot_begin_pv.argtypes = [vp, vp, vp, vp, C.POINTER(C.c_int64)]
    L.mc_snapshot_begin_xyz.argtypes = [vp, vp, vp, C.POINTER(i64), C.POINTER(i64)]
    L.mc_snapshot_wait.argtypes = [vp]
    L.mc_comm_schedule.argtypes = [vp, C.POINTER(i32), C.POINTER(C.c_double)]
    L.mc_comm_halo_mode.argtypes = [vp, C.POINTER(i32), C.c_char_p, i32]
    _LIB = L
    return L
