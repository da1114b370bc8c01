"""Thin Python handle over the C ABI (include/molchanica_md.h) -- the harness that tests/ and
bench.py drive.  It adds nothing to the data path: every method is one C call on host numpy
buffers, exactly what a Rust `extern "C"` caller would do.

Method names follow the reference-side surface they stand in for:
    MdState::new                 -> MdEngine.from_workload / set_* (src/md/mod.rs:641-693)
    MdState::step(dev, dt, ext)  -> MdEngine.step(dt, n_steps, ext_forces)  (src/md/mod.rs:716,748)
    rebuild_spatial_caches(dev)  -> MdEngine.build_neighbors()  (properties/sol_shrinking_box.rs:632)
    compute_energy_snapshot      -> MdEngine.compute_forces() + energy()  (src/md/mod.rs:1036)
    calc_binding_energy          -> MdEngine.dock_score(...)  (src/docking/legacy/mod.rs:210-383)
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class McError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"molchanica_md error {code}: {msg}")
        self.code = code


def _f4(a, n=None):
    a = np.ascontiguousarray(a, np.float32)
    assert a.ndim == 2 and a.shape[1] == 4 and (n is None or len(a) == n), a.shape
    return a


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class MdEngine:
    """One engine handle == one CUDA device + stream (ComputationDevice::Gpu in the reference,
    src/util.rs:1072-1119).  Raises McError(MC_E_NODEVICE) when no B200 is visible."""

    def __init__(self, device=0):
        self._L = _lib.lib()
        h = C.c_void_p()
        rc = self._L.mc_create(int(device), C.byref(h))
        if rc != 0:
            raise McError(rc, self._L.mc_last_error(None).decode())
        self._h = h
        self.n = 0
        self.warnings = []

    # -- plumbing
    def _chk(self, rc):
        if rc < 0:
            raise McError(rc, self._L.mc_last_error(self._h).decode())
        if rc > 0:  # MC_W_*: the call completed; keep the warning for whoever wants to look
            self.warnings.append((rc, self._L.mc_last_error(self._h).decode()))

    def close(self):
        if getattr(self, "_h", None):
            self._L.mc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- system definition
    def set_box(self, lo, hi, periodic):
        lo = np.ascontiguousarray(lo, np.float32)
        hi = np.ascontiguousarray(hi, np.float32)
        self._chk(self._L.mc_set_box(self._h, _ptr(lo), _ptr(hi), int(bool(periodic))))

    def set_atoms(self, xyzq, type=None, vel=None, flags=None):
        xyzq = _f4(xyzq)
        n = len(xyzq)
        t = None if type is None else np.ascontiguousarray(type, np.uint16)
        v = None if vel is None else _f4(vel, n)
        f = None if flags is None else np.ascontiguousarray(flags, np.uint8)
        self._chk(self._L.mc_set_atoms(self._h, n, _ptr(xyzq), _ptr(t), _ptr(v), _ptr(f)))
        self.n = n

    def set_lj_table(self, sigma_eps):
        t = np.ascontiguousarray(sigma_eps, np.float32)
        assert t.ndim == 3 and t.shape[0] == t.shape[1] and t.shape[2] == 2
        self._chk(self._L.mc_set_lj_table(self._h, t.shape[0], _ptr(t)))

    def set_exclusions(self, start, idx):
        if start is None or idx is None or len(idx) == 0:
            self._chk(self._L.mc_set_exclusions(self._h, None, None))
            return
        s = np.ascontiguousarray(start, np.int32)
        i = np.ascontiguousarray(idx, np.int32)
        self._chk(self._L.mc_set_exclusions(self._h, _ptr(s), _ptr(i)))

    def set_pairs14(self, pairs, scale_lj=0.5, scale_q=1.0 / 1.2):
        p = None if pairs is None or len(pairs) == 0 else np.ascontiguousarray(pairs, np.int32)
        self._chk(self._L.mc_set_pairs14(self._h, 0 if p is None else len(p), _ptr(p), scale_lj, scale_q))

    def set_bonded(self, bonds=None, bond_kr0=None, angles=None, angle_kt0=None, dihedrals=None, dihedral_prm=None):
        """mc_set_bonds / mc_set_angles / mc_set_dihedrals (None or empty clears a kind)."""
        def pair(ids, prm, width):
            if ids is None or len(ids) == 0:
                return 0, None, None
            i = np.ascontiguousarray(ids, np.int32).reshape(-1, width)
            return len(i), i, np.ascontiguousarray(prm, np.float32)
        m, i, p = pair(bonds, bond_kr0, 2)
        self._chk(self._L.mc_set_bonds(self._h, m, _ptr(i), _ptr(p)))
        m, i, p = pair(angles, angle_kt0, 3)
        self._chk(self._L.mc_set_angles(self._h, m, _ptr(i), _ptr(p)))
        m, i, p = pair(dihedrals, dihedral_prm, 4)
        self._chk(self._L.mc_set_dihedrals(self._h, m, _ptr(i), _ptr(p)))

    def energy_between_mols(self, mol_id):
        """mc_set_molecule_ids + mc_get_energy_between_mols: nonbonded energy between atoms of different molecules."""
        m = np.ascontiguousarray(mol_id, np.uint16)
        assert len(m) == self.n
        self._chk(self._L.mc_set_molecule_ids(self._h, _ptr(m)))
        out = C.c_double(0.0)
        self._chk(self._L.mc_get_energy_between_mols(self._h, C.byref(out)))
        return float(out.value)

    def pressure(self):
        """mc_get_pressure: (pressure in bar, virial in kcal/mol)."""
        p, w = C.c_double(0.0), C.c_double(0.0)
        self._chk(self._L.mc_get_pressure(self._h, C.byref(p), C.byref(w)))
        return p.value, w.value

    def set_barostat(self, kind, pressure_bar=1.0, tau_ps=5.0, compressibility_per_bar=4.5e-5, every=10, seed=0):
        """mc_set_barostat: kind 0 none, 1 Berendsen, 2 stochastic cell rescaling."""
        self._chk(self._L.mc_set_barostat(self._h, int(kind), float(pressure_bar), float(tau_ps), float(compressibility_per_bar), int(every),
                                          int(seed)))

    def box(self):
        lo, hi = (C.c_float * 3)(), (C.c_float * 3)()
        self._chk(self._L.mc_get_box(self._h, lo, hi))
        return np.array(lo[:], np.float32), np.array(hi[:], np.float32)

    def set_hbond_constraints(self, clusters, lengths):
        """clusters (m, 4): heavy atom + up to three hydrogens (-1 = unused); lengths (m, 3)."""
        if clusters is None or len(clusters) == 0:
            self._chk(self._L.mc_set_hbond_constraints(self._h, 0, None, None))
            return
        q = np.ascontiguousarray(clusters, np.int32).reshape(-1, 4)
        d = np.ascontiguousarray(lengths, np.float32).reshape(-1, 3)
        self._chk(self._L.mc_set_hbond_constraints(self._h, len(q), _ptr(q), _ptr(d)))

    def set_virtual_sites(self, quads, a, b):
        q = None if quads is None or len(quads) == 0 else np.ascontiguousarray(quads, np.int32).reshape(-1, 4)
        self._chk(self._L.mc_set_virtual_sites(self._h, 0 if q is None else len(q), _ptr(q), a, b))

    def set_thermostat(self, kind, temperature_k=300.0, gamma_per_ps=1.0, seed=0):
        """kind: 0 = none, 1 = Langevin (mc_set_thermostat)."""
        self._chk(self._L.mc_set_thermostat(self._h, int(kind), temperature_k, gamma_per_ps, int(seed)))

    def set_pme(self, k1, k2, k3):
        self._chk(self._L.mc_set_pme(self._h, int(k1), int(k2), int(k3)))

    def set_rigid_waters(self, triples, d_oh, d_hh, m_o=15.999, m_h=1.008):
        t = None if triples is None or len(triples) == 0 else np.ascontiguousarray(triples, np.int32).reshape(-1, 3)
        self._chk(self._L.mc_set_rigid_waters(self._h, 0 if t is None else len(t), _ptr(t), d_oh, d_hh, m_o, m_h))

    def set_cutoffs(self, rc_lj, rc_q, skin, coulomb_mode, alpha=0.35):
        self._chk(self._L.mc_set_cutoffs(self._h, rc_lj, rc_q, skin, int(coulomb_mode), alpha))

    def set_overrides(self, lj_disabled=False, coulomb_disabled=False):
        self._chk(self._L.mc_set_overrides(self._h, int(lj_disabled), int(coulomb_disabled)))

    def set_option(self, name, value):
        self._chk(self._L.mc_set_option(self._h, name.encode(), float(value)))

    def set_positions(self, xyzq):
        a = _f4(xyzq, self.n)
        self._chk(self._L.mc_set_positions(self._h, _ptr(a)))

    def set_velocities(self, vel):
        a = _f4(vel, self.n)
        self._chk(self._L.mc_set_velocities(self._h, _ptr(a)))

    @classmethod
    def from_workload(cls, w, device=0, bonded=False):
        """Everything MdState::new hands to the engine, from a workloads.py dict (bonded = True: also the
        bonds / angles / dihedrals the workload carries)."""
        e = cls(device)
        lo = np.asarray(w["box_lo"], np.float32)
        e.set_box(lo, lo + np.asarray(w["box_ext"], np.float32), w["periodic"])
        e.set_cutoffs(w["rc_lj"], w["rc_q"], w["skin"], w["coul_mode"], w.get("alpha", 0.35))
        e.set_lj_table(w["ljtab"])
        e.set_atoms(w["xyzq"], w["type"], w["vel"], w.get("flags"))
        e.set_exclusions(w.get("excl_start"), w.get("excl_idx"))
        e.set_pairs14(w.get("pairs14"), w.get("scale14_lj", 0.5), w.get("scale14_q", 1 / 1.2))
        if bonded:
            e.set_bonded(w.get("bonds"), w.get("bond_kr0"), w.get("angles"), w.get("angle_kt0"),
                         w.get("dihedrals"), w.get("dihedral_prm"))
        return e

    # -- hot path
    def build_neighbors(self):
        self._chk(self._L.mc_build_neighbors(self._h))

    def compute_forces(self):
        self._chk(self._L.mc_compute_forces(self._h))

    def step(self, dt, n_steps=1, ext_forces=None):
        ef = None if ext_forces is None else np.ascontiguousarray(ext_forces, np.float32)
        self._chk(self._L.mc_step(self._h, dt, int(n_steps), _ptr(ef)))

    def minimize_energy(self, max_iters):
        """mc_minimize_energy: (accepted moves, potential energy before, after)."""
        k, e0, e1 = C.c_int32(0), C.c_double(0.0), C.c_double(0.0)
        self._chk(self._L.mc_minimize_energy(self._h, int(max_iters), C.byref(k), C.byref(e0), C.byref(e1)))
        return int(k.value), float(e0.value), float(e1.value)

    def last_step_ms(self):
        return self._L.mc_last_step_ms(self._h)

    def step_raw(self, dt, n_steps, ext_ptr):
        """mc_step with a caller-held host pointer (pinned memory), no numpy conversion."""
        self._chk(self._L.mc_step(self._h, dt, int(n_steps), ext_ptr))

    def get_positions_into(self, ptr):
        self._chk(self._L.mc_get_positions(self._h, ptr))

    # -- read-back
    def _get4(self, fn, n=None):
        out = np.empty((self.n if n is None else n, 4), np.float32)
        self._chk(fn(self._h, _ptr(out)))
        return out

    def positions(self):
        return self._get4(self._L.mc_get_positions)

    def velocities(self):
        return self._get4(self._L.mc_get_velocities)

    def snapshot_begin(self, out_positions, out_ids=None):
        """mc_snapshot_begin into caller-held (ideally pinned) host arrays; returns the entry count."""
        n_out = C.c_int64(0)
        self._chk(self._L.mc_snapshot_begin(self._h, _ptr(out_positions), _ptr(out_ids), C.byref(n_out)))
        return int(n_out.value)

    def snapshot_begin_pv(self, out_positions, out_velocities, out_ids=None):
        """mc_snapshot_begin_pv: positions and velocities."""
        n_out = C.c_int64(0)
        self._chk(self._L.mc_snapshot_begin_pv(self._h, _ptr(out_positions), _ptr(out_velocities), _ptr(out_ids), C.byref(n_out)))
        return int(n_out.value)

    def snapshot_begin_xyz(self, out_xyz, out_ids=None, have_epoch=-1):
        """mc_snapshot_begin_xyz: packed float3 positions; returns (entries, layout epoch).  have_epoch: the epoch whose ids the
        caller already holds (out_ids is only filled when the layout is another one)."""
        n_out, ep = C.c_int64(0), C.c_int64(have_epoch)
        self._chk(self._L.mc_snapshot_begin_xyz(self._h, _ptr(out_xyz), _ptr(out_ids), C.byref(n_out), C.byref(ep)))
        return int(n_out.value), int(ep.value)

    def snapshot_wait(self):
        self._chk(self._L.mc_snapshot_wait(self._h))

    def forces(self):
        return self._get4(self._L.mc_get_forces)

    def energy(self):
        e = _lib.McEnergy()
        self._chk(self._L.mc_get_energy(self._h, C.byref(e)))
        return {k: getattr(e, k) for k, _ in e._fields_}

    # -- docking pose set
    def dock_make_poses(self, site_center, site_radius, num_posits=8, num_orientations=60):
        c = np.ascontiguousarray(site_center, np.float64)
        n = C.c_int64(0)
        self._chk(self._L.mc_dock_make_poses(_ptr(c), float(site_radius), int(num_posits), int(num_orientations), None, 0, C.byref(n)))
        out = np.zeros((n.value, 7), np.float32)
        self._chk(self._L.mc_dock_make_poses(_ptr(c), float(site_radius), int(num_posits), int(num_orientations), _ptr(out), n.value, C.byref(n)))
        return out

    def dock_make_poses_flex(self, site_center, site_radius, n_flex_bonds, angles_per_bond, num_posits=8, num_orientations=60):
        """mc_dock_make_poses_flex: n x (7 + n_flex_bonds) floats."""
        c = np.ascontiguousarray(site_center, np.float64)
        n = C.c_int64(0)
        args = (_ptr(c), float(site_radius), int(num_posits), int(num_orientations), int(n_flex_bonds), int(angles_per_bond))
        self._chk(self._L.mc_dock_make_poses_flex(*args, None, 0, C.byref(n)))
        out = np.zeros((n.value, 7 + n_flex_bonds), np.float32)
        self._chk(self._L.mc_dock_make_poses_flex(*args, _ptr(out), n.value, C.byref(n)))
        return out

    def dock_flex_masks(self, n_lig, bonds, flex_bond_idx):
        """mc_dock_flex_masks: (axis (F, 2) int32, mask (F, n_lig) uint8)."""
        b = np.ascontiguousarray(bonds, np.int32).reshape(-1, 2)
        fi = np.ascontiguousarray(flex_bond_idx, np.int32)
        axis = np.zeros((len(fi), 2), np.int32)
        mask = np.zeros((len(fi), n_lig), np.uint8)
        rc = self._L.mc_dock_flex_masks(int(n_lig), len(b), _ptr(b), len(fi), _ptr(fi), _ptr(axis), _ptr(mask))
        if rc != 0:
            raise McError(rc, "mc_dock_flex_masks: bad bond list or a flexible bond inside a ring")
        return axis, mask

    def dock_filter_poses(self, rec_near, rec_is_carbon, lig, lig_is_carbon, lig_anchor, poses, vdw_radius=1.7, gpu=True):
        """keep mask of the clash pre-filter: on the device (mc_dock_filter_poses_gpu) or with the host twin."""
        rec = _f4(rec_near)
        lg = _f4(lig)
        rc_ = np.ascontiguousarray(rec_is_carbon, np.uint8)
        lc_ = np.ascontiguousarray(lig_is_carbon, np.uint8)
        an = np.ascontiguousarray(lig_anchor, np.float32)
        ps = np.ascontiguousarray(poses, np.float32)
        keep = np.zeros(len(ps), np.uint8)
        kept = C.c_int64(0)
        if gpu:
            self._chk(self._L.mc_dock_filter_poses_gpu(self._h, len(rec), _ptr(rec), _ptr(rc_), len(lg), _ptr(lg), _ptr(lc_), _ptr(an),
                                                       vdw_radius, len(ps), _ptr(ps), _ptr(keep), C.byref(kept)))
        else:
            rc = self._L.mc_dock_filter_poses(len(rec), _ptr(rec), _ptr(rc_), len(lg), _ptr(lg), _ptr(lc_), _ptr(an), vdw_radius, len(ps),
                                              _ptr(ps), _ptr(keep), C.byref(kept))
            assert rc == 0
        return keep

    def halo_mode(self):
        """(fused, why): 1 when the step kernels exchange ghosts over mapped peer memory, else 0 + the reason."""
        fused = C.c_int32(0)
        buf = C.create_string_buffer(256)
        self._chk(self._L.mc_comm_halo_mode(self._h, C.byref(fused), buf, 256))
        return int(fused.value), buf.value.decode()

    def schedule(self):
        """(interval, last_disp_frac) of a decomposed run, see mc_comm_schedule."""
        k, f = C.c_int32(0), C.c_double(0.0)
        self._chk(self._L.mc_comm_schedule(self._h, C.byref(k), C.byref(f)))
        return int(k.value), float(f.value)

    def stats(self):
        s = _lib.McStats()
        self._chk(self._L.mc_get_stats(self._h, C.byref(s)))
        d = {k: getattr(s, k) for k, _ in s._fields_}
        d["n_cells"] = list(s.n_cells)
        return d

    def ext_upload_bytes(self):
        """host-to-device bytes the last mc_step moved for its external forces on this rank"""
        return self.stats()["ext_upload_bytes"]

    def reset_timers(self):
        self._chk(self._L.mc_reset_timers(self._h))

    def neighbors(self):
        """Verlet list as CSR (start int64 n+1, idx int32) in original ids, rows ascending."""
        start = np.zeros(self.n + 1, np.int64)
        tot = C.c_int64(0)
        self._chk(self._L.mc_get_neighbors(self._h, _ptr(start), None, 0, C.byref(tot)))
        idx = np.zeros(max(tot.value, 1), np.int32)
        self._chk(self._L.mc_get_neighbors(self._h, _ptr(start), _ptr(idx), len(idx), C.byref(tot)))
        return start, idx[:tot.value]

    def time_pair_kernel(self, reps=20, flush_l2=True):
        self._chk(self._L.mc_time_kernels(self._h, int(reps), int(flush_l2)))
        return self._L.mc_last_pair_kernel_ms(self._h)

    # -- docking
    def dock_score(self, d, poses=None):
        """(P,5) f32 {score, vdw, hydrophobic, electrostatic, coulomb_e} for a docking_c5 dict."""
        poses = np.ascontiguousarray(d["poses"] if poses is None else poses, np.float32)
        rec, lig = _f4(d["rec"]), _f4(d["lig"])
        rt = np.ascontiguousarray(d["rec_type"], np.uint16)
        lt = np.ascontiguousarray(d["lig_type"], np.uint16)
        rh = np.ascontiguousarray(d["rec_hphob"], np.uint8)
        lh = np.ascontiguousarray(d["lig_hphob"], np.uint8)
        anchor = np.ascontiguousarray(d["lig_anchor"], np.float32)
        tab = np.ascontiguousarray(d["ljtab"], np.float32)
        out = np.empty((len(poses), 5), np.float32)
        self._chk(self._L.mc_dock_score(self._h, len(rec), _ptr(rec), _ptr(rt), _ptr(rh), len(lig), _ptr(lig), _ptr(lt),
                                        _ptr(lh), _ptr(anchor), tab.shape[0], tab.shape[1], _ptr(tab), len(poses),
                                        _ptr(poses), _ptr(out)))
        return out

    def dock_score_flex(self, d, poses, flex_axis, flex_mask):
        """mc_dock_score_flex: poses (P, 7 + F) with F torsion angles; flex_axis (F, 2), flex_mask (F, n_lig)."""
        poses = np.ascontiguousarray(poses, np.float32)
        rec, lig = _f4(d["rec"]), _f4(d["lig"])
        rt = np.ascontiguousarray(d["rec_type"], np.uint16)
        lt = np.ascontiguousarray(d["lig_type"], np.uint16)
        rh = np.ascontiguousarray(d["rec_hphob"], np.uint8)
        lh = np.ascontiguousarray(d["lig_hphob"], np.uint8)
        anchor = np.ascontiguousarray(d["lig_anchor"], np.float32)
        tab = np.ascontiguousarray(d["ljtab"], np.float32)
        ax = np.ascontiguousarray(flex_axis, np.int32).reshape(-1, 2)
        mk = np.ascontiguousarray(flex_mask, np.uint8).reshape(len(ax), len(lig))
        assert poses.shape[1] == 7 + len(ax)
        out = np.empty((len(poses), 5), np.float32)
        self._chk(self._L.mc_dock_score_flex(self._h, len(rec), _ptr(rec), _ptr(rt), _ptr(rh), len(lig), _ptr(lig), _ptr(lt),
                                             _ptr(lh), _ptr(anchor), tab.shape[0], tab.shape[1], _ptr(tab), len(ax), _ptr(ax),
                                             _ptr(mk), len(poses), _ptr(poses), _ptr(out)))
        return out

    def last_dock_kernel_ms(self):
        return self._L.mc_last_dock_kernel_ms(self._h)
