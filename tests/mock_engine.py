"""An engine stand-in built on the CPU oracle, with the MdEngine methods the GPU worker scripts use.  TEST INFRASTRUCTURE:
tests/test_gpu_workers_dry_run.py runs the worker scripts of the not-yet-run device components against it, so that
their own Python (keys, shapes, thresholds) is known to be sound before they first meet a GPU.  It is never used as,
or in place of, the product engine."""
import ctypes as C

import numpy as np

from oracle import oracle_py as O
from oracle import pme_oracle as P

class MockEngine:
    def __init__(self, w=0):
        if not isinstance(w, dict):  # MdEngine(device): a bare handle, as the docking helpers use it
            w = dict(xyzq=np.zeros((1, 4), np.float32), vel=np.zeros((1, 4), np.float32))
        self._init(w)

    def dock_make_poses(self, site, radius, n_pos=8, n_or=60):
        from oracle import dock_poses as DP
        return DP.make_poses(site, radius, n_pos, n_or)

    def dock_filter_poses(self, rec, rec_c, lig, lig_c, anchor, poses, vdw_radius=1.7, gpu=True):
        from oracle import dock_poses as DP
        return DP.filter_poses(np.asarray(rec), np.asarray(rec_c), np.asarray(lig), np.asarray(lig_c), np.asarray(anchor, np.float32), poses,
                               vdw_radius)

    def _init(self, w):
        self.w = dict(w); self.x = np.array(w["xyzq"], np.float32); self.v = np.array(w["vel"], np.float32)
        self.bonded = None; self.rigid = None; self.vs = None; self.lgv = None; self.csvr = None; self.pme = None
    @classmethod
    def from_workload(cls, w, device=0, bonded=False):
        e = cls(w)
        if bonded: e.bonded = True
        return e
    def set_bonded(self, bonds=None, kr0=None, *a):
        if bonds is not None and np.max(bonds) >= len(self.x): raise RuntimeError("range")
        self.with_bonds = True
    def set_rigid_waters(self, t, doh, dhh, mo=15.999, mh=1.008): self.rigid = (t, doh, dhh)
    def set_virtual_sites(self, q, a, b): self.vs = (q, a, b)
    def set_hbond_constraints(self, c, l): self.hc = (c, l)
    def set_thermostat(self, kind, T, g, seed=0):
        self.lgv = (T, g, seed) if kind == 1 else None; self.csvr = (T, g, seed) if kind == 2 else None
    def set_pme(self, *K): self.pme = K
    def _w(self): return dict(self.w, xyzq=self.x, vel=self.v)
    def step(self, dt, n):
        r = O.md_run(dict(self._w(), dt=dt), n, precision=32, with_bonds=getattr(self, "with_bonds", False), rigid_waters=self.rigid, hbond_constraints=getattr(self, "hc", None),
                     virtual_sites=self.vs, langevin=self.lgv, csvr=self.csvr)
        self.x, self.v = r["xyzq"], r["vel"]
    def compute_forces(self): pass
    def positions(self): return self.x.copy()
    def velocities(self): return self.v.copy()
    def forces(self):
        w = self._w(); nb = O.neighbors(w); f, _, _ = O.forces(w, nb, precision=64); f = f.astype(np.float32)
        if self.bonded: f[:, :3] += O.bonded(w)[0]
        if self.pme:
            f[:, :3] += P.spme(w["xyzq"], w["box_lo"], w["box_ext"], 0.35, self.pme)[1] + P.excl_correction(w["xyzq"], w["box_ext"], True, w["excl_start"], w["excl_idx"], 0.35)[1]
        if self.vs:
            q, a, b = self.vs
            for m, o, h1, h2 in q:
                f[o, :3] += (1 - a - b) * f[m, :3]; f[h1, :3] += a * f[m, :3]; f[h2, :3] += b * f[m, :3]; f[m, :3] = 0
        return f
    def energy(self):
        w = self._w(); nb = O.neighbors(w); _, _, en = O.forces(w, nb, precision=64)
        d = dict(energy_potential_nonbonded=float(en.sum()), energy_potential_bonded=0.0, energy_bond=0.0, energy_angle=0.0, energy_dihedral=0.0,
                 energy_pme=0.0, volume=float(np.prod(w["box_ext"])), density=1.0,
                 temperature=self._temp())
        if self.bonded or getattr(self, "with_bonds", False):
            e3 = O.bonded(w)[1]; d.update(energy_bond=float(e3[0]), energy_angle=float(e3[1]), energy_dihedral=float(e3[2]), energy_potential_bonded=float(e3.sum()))
        if self.pme:
            d["energy_pme"] = P.spme(w["xyzq"], w["box_lo"], w["box_ext"], 0.35, self.pme)[0] + P.excl_correction(w["xyzq"], w["box_ext"], True, w["excl_start"], w["excl_idx"], 0.35)[0] + P.self_energy(w["xyzq"], 0.35)
            d["energy_potential_nonbonded"] += d["energy_pme"]
        d["energy_potential"] = d["energy_potential_nonbonded"] + d["energy_potential_bonded"]
        return d
    def _temp(self):
        m = np.where(self.v[:, 3] > 0, 1.0 / np.maximum(self.v[:, 3], 1e-30), 0.0)
        ke = 0.5 * (m[:, None] * self.v[:, :3].astype(np.float64) ** 2).sum() / 418.4
        nm = int((self.v[:, 3] > 0).sum()); nw = 0 if self.rigid is None else len(self.rigid[0])
        return float(2 * ke / ((3 * nm - 3 * nw) * 0.0019872041))
    def minimize_energy(self, max_iters):
        r = O.minimize(dict(self._w(), pairs14=None), max_iters)
        self.x = r["xyzq"]
        return r["accepted"], r["e_initial"], r["e_final"]
    def energy_between_mols(self, mol):
        from util import between_mols_reference
        w = self._w(); s_, i_ = O.neighbors(w)
        return between_mols_reference(w, np.asarray(mol), s_, i_)
    def close(self): pass
