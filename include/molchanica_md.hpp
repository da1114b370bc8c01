// molchanica_md.hpp -- C++ host-side mirror of the `dynamics` API surface Molchanica drives
// (reference is compiled Rust; no Rust toolchain in this image, so the host layer above the C ABI is
// C++ and keeps the reference's names, argument meaning and error behaviour).  Header only; links
// against libmolchanica_md.so (include/molchanica_md.h).
//
//   reference (Rust, crate `dynamics`)                           here
//   ComputationDevice::{Cpu, Gpu(stream)}   src/util.rs:1072     ComputationDevice{cuda_ordinal}  (no Cpu variant: no fallback)
//   MdConfig{coulomb_cutoff, lj_cutoff, ..} ui/panels/md.rs:260  MdConfig
//   MdOverrides{lj_disabled, ..}            src/md/mod.rs:671    MdOverrides
//   AtomDynamics{posit, force, static_, ..} src/md/mod.rs:843    AtomDynamics
//   SimBox{bounds_low, bounds_high}         sol_shrinking_box.rs:600   SimBox
//   MdState::new(dev,&cfg,&mols,params)     src/md/mod.rs:689    MdState::create(dev, cfg, system)
//   md.step(dev, dt, Option<forces>)        src/md/mod.rs:716    MdState::step(dev, dt, ext)
//   md.rebuild_spatial_caches(dev)          sol_shrinking_box.rs:632   MdState::rebuild_spatial_caches(dev)
//   compute_energy_snapshot(..)             src/md/mod.rs:1036   MdState::energy_snapshot(dev)
//   Result<_, ParamError{descrip}>          src/md/mod.rs:651    throws ParamError
//
// `MdState::new` in the reference also does Amber typing, solvation and relaxation; that plumbing
// stays where it is (BASELINE.json: "Amber-parameter plumbing untouched") -- its OUTPUT (per-atom
// charge / LJ type / mass, the LJ table, exclusions, 1-4 pairs) is what `MdSystem` carries.
#pragma once
#include <cmath>
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "molchanica_md.h"

namespace molchanica {

struct ParamError : std::runtime_error {
    int code;
    ParamError(int c, const std::string &descrip) : std::runtime_error(descrip), code(c) {}
};

struct Vec3F32 { float x = 0, y = 0, z = 0; };

struct ComputationDevice {  // ComputationDevice::Gpu; there is deliberately no Cpu variant
    int cuda_ordinal = 0;
};

struct SimBox {
    Vec3F32 bounds_low, bounds_high;
    bool periodic = true;  // false == Solvent::None (vacuum), src/md/mod.rs:784
};

struct MdOverrides {
    bool lj_disabled = false;
    bool coulomb_disabled = false;
};

enum class CoulombMode : int { None = MC_COULOMB_NONE, PlainCutoff = MC_COULOMB_PLAIN, EwaldRealSpace = MC_COULOMB_ERFC };

// dynamics::BarostatCfg{pressure_target, tau} (ui/panels/md.rs:517-556, properties/crystal.rs:312)
struct BarostatCfg {
    float pressure_target = 1.0f;        // bar
    float tau = 5.0f;                    // ps ("5ps is a good default", ui/panels/md.rs:544)
    float compressibility = 4.5e-5f;     // 1/bar (water)
    int every_n_steps = 10;
    bool stochastic = true;              // stochastic cell rescaling (needs a thermostat); false = Berendsen
};

struct MdConfig {
    // Integrator::VerletVelocity{thermostat: Some(tau)} + temp_target (properties/crystal.rs:306-311): CSVR with time
    // constant tau; 0 = no thermostat (NVE)
    float thermostat_tau = 0.0f;   // ps
    float temp_target = 300.0f;    // K
    uint64_t seed = 0;
    bool zero_com_drift = false;   // properties/crystal.rs:310
    bool has_barostat = false;     // barostat_cfg: Option<BarostatCfg>
    BarostatCfg barostat_cfg;
    float coulomb_cutoff = 12.0f;  // A, ui/panels/md.rs:260
    float lj_cutoff = 12.0f;       // A, ui/panels/md.rs:261
    float skin = 2.0f;             // A, Verlet skin
    CoulombMode coulomb_mode = CoulombMode::PlainCutoff;
    float ewald_alpha = 0.35f;
    SimBox sim_box;
    MdOverrides overrides;
};

struct AtomDynamics {
    Vec3F32 posit, vel, force;
    float mass = 1.0f;            // amu
    float partial_charge = 0.0f;  // e (scaled by sqrt(332.0522) on upload)
    uint16_t lj_type = 0;
    bool static_ = false;
    float energy_row = 0.0f;      // sum of the pair energies of this atom's list row
};

// What the reference's MdState::new derives from MolDynamics + FfParamSet and hands to the engine.
struct MdSystem {
    std::vector<AtomDynamics> atoms;
    int n_lj_types = 1;
    std::vector<float> lj_sigma_eps;       // n_lj_types^2 x (sigma, eps)
    std::vector<int32_t> excl_start, excl_idx;  // CSR of excluded partners (1-2, 1-3, 1-4)
    std::vector<int32_t> pairs14;          // 2 x m
    float scale14_lj = 0.5f, scale14_coulomb = 1.0f / 1.2f;
    // bonded terms, Amber forms (mc_set_bonds / mc_set_angles / mc_set_dihedrals); empty = none
    std::vector<int32_t> bonds, angles, dihedrals;           // 2 / 3 / 4 atom ids per term
    std::vector<float> bond_k_r0, angle_k_theta0, dihedral_pk_n_phase;  // 2 / 2 / 3 parameters per term
    // rigid three-site waters (mc_set_rigid_waters): (O, H1, H2) ids; the reference's md.water (sol_shrinking_box.rs:605-613)
    std::vector<int32_t> rigid_waters;
    float water_d_oh = 0.9572f, water_d_hh = 1.5139f, water_m_o = 15.999f, water_m_h = 1.008f;
    // SPME grid (mc_set_pme), 0 = off; used with CoulombMode::EwaldRealSpace
    int pme_grid[3] = {0, 0, 0};
};

struct SnapshotEnergyData {  // src/md/mod.rs:1242-1245, ui/panels/md_viewer.rs:202-256
    double energy_potential = 0, energy_potential_nonbonded = 0, energy_potential_bonded = 0;
    double energy_kinetic = 0, temperature = 0;
    double volume = 0, density = 0;  // A^3, g/cm^3 (periodic boxes)
    double pressure = 0;             // bar (periodic boxes; 0 when it cannot be evaluated yet)
};

class MdState {
  public:
    std::vector<AtomDynamics> atoms;  // host-visible state, refreshed by sync_atoms() (sol_shrinking_box.rs:776-789)
    SimBox cell;

    static MdState create(const ComputationDevice &dev, const MdConfig &cfg, const MdSystem &sys) {
        MdState md;
        md.atoms = sys.atoms;
        md.cell = cfg.sim_box;
        md.chk(mc_create(dev.cuda_ordinal, &md.ctx_), nullptr);
        const float lo[3] = {cfg.sim_box.bounds_low.x, cfg.sim_box.bounds_low.y, cfg.sim_box.bounds_low.z};
        const float hi[3] = {cfg.sim_box.bounds_high.x, cfg.sim_box.bounds_high.y, cfg.sim_box.bounds_high.z};
        md.chk(mc_set_box(md.ctx_, lo, hi, cfg.sim_box.periodic ? 1 : 0));
        md.chk(mc_set_cutoffs(md.ctx_, cfg.lj_cutoff, cfg.coulomb_cutoff, cfg.skin, (int)cfg.coulomb_mode, cfg.ewald_alpha));
        md.chk(mc_set_overrides(md.ctx_, cfg.overrides.lj_disabled, cfg.overrides.coulomb_disabled));
        md.chk(mc_set_lj_table(md.ctx_, sys.n_lj_types, sys.lj_sigma_eps.data()));
        const size_t n = sys.atoms.size();
        std::vector<mc_float4> x(n), v(n);
        std::vector<uint16_t> t(n);
        std::vector<uint8_t> f(n);
        const float qs = std::sqrt(332.0522f);
        for (size_t i = 0; i < n; ++i) {
            const AtomDynamics &a = sys.atoms[i];
            x[i] = {a.posit.x, a.posit.y, a.posit.z, a.partial_charge * qs};
            v[i] = {a.vel.x, a.vel.y, a.vel.z, a.static_ ? 0.0f : 1.0f / a.mass};
            t[i] = a.lj_type;
            f[i] = a.static_ ? MC_FLAG_STATIC : 0;
        }
        md.chk(mc_set_atoms(md.ctx_, (int64_t)n, x.data(), t.data(), v.data(), f.data()));
        if (!sys.excl_idx.empty()) md.chk(mc_set_exclusions(md.ctx_, sys.excl_start.data(), sys.excl_idx.data()));
        if (!sys.pairs14.empty())
            md.chk(mc_set_pairs14(md.ctx_, (int64_t)sys.pairs14.size() / 2, sys.pairs14.data(), sys.scale14_lj, sys.scale14_coulomb));
        if (!sys.bonds.empty()) md.chk(mc_set_bonds(md.ctx_, (int64_t)sys.bonds.size() / 2, sys.bonds.data(), sys.bond_k_r0.data()));
        if (!sys.angles.empty()) md.chk(mc_set_angles(md.ctx_, (int64_t)sys.angles.size() / 3, sys.angles.data(), sys.angle_k_theta0.data()));
        if (!sys.dihedrals.empty())
            md.chk(mc_set_dihedrals(md.ctx_, (int64_t)sys.dihedrals.size() / 4, sys.dihedrals.data(), sys.dihedral_pk_n_phase.data()));
        if (!sys.rigid_waters.empty())
            md.chk(mc_set_rigid_waters(md.ctx_, (int64_t)sys.rigid_waters.size() / 3, sys.rigid_waters.data(), sys.water_d_oh,
                                       sys.water_d_hh, sys.water_m_o, sys.water_m_h));
        if (sys.pme_grid[0] > 0) md.chk(mc_set_pme(md.ctx_, sys.pme_grid[0], sys.pme_grid[1], sys.pme_grid[2]));
        if (cfg.thermostat_tau > 0.0f)
            md.chk(mc_set_thermostat(md.ctx_, MC_THERMOSTAT_CSVR, cfg.temp_target, 1.0f / cfg.thermostat_tau, cfg.seed));
        if (cfg.zero_com_drift) md.chk(mc_set_option(md.ctx_, "zero_com_drift", 100.0));
        if (cfg.has_barostat) {
            const BarostatCfg &b = cfg.barostat_cfg;
            md.chk(mc_set_barostat(md.ctx_, b.stochastic && cfg.thermostat_tau > 0.0f ? MC_BAROSTAT_CRESCALE : MC_BAROSTAT_BERENDSEN,
                                   b.pressure_target, b.tau, b.compressibility, b.every_n_steps, cfg.seed));
        }
        return md;
    }

    MdState(MdState &&o) noexcept : atoms(std::move(o.atoms)), cell(o.cell), ctx_(o.ctx_) { o.ctx_ = nullptr; }
    MdState &operator=(MdState &&o) noexcept {
        if (this != &o) { release(); atoms = std::move(o.atoms); cell = o.cell; ctx_ = o.ctx_; o.ctx_ = nullptr; }
        return *this;
    }
    MdState(const MdState &) = delete;
    MdState &operator=(const MdState &) = delete;
    ~MdState() { release(); }

    // md.step(dev, dt, Some(forces) | None): one velocity-Verlet step (src/md/mod.rs:716,748); `n_steps`
    // batches the 10 steps per frame the GUI takes (MD_STEPS_PER_APPLICATION_FRAME, src/md/mod.rs:45).
    void step(const ComputationDevice &, float dt, const std::optional<std::vector<Vec3F32>> &external_force = std::nullopt,
              int n_steps = 1) {
        const float *ext = nullptr;
        std::vector<float> flat;
        if (external_force) {
            if (external_force->size() != atoms.size()) throw ParamError(MC_E_INVALID, "external force count != atom count");
            flat.reserve(3 * atoms.size());
            for (const Vec3F32 &e : *external_force) { flat.push_back(e.x); flat.push_back(e.y); flat.push_back(e.z); }
            ext = flat.data();
        }
        chk(mc_step(ctx_, dt, n_steps, ext));
    }

    void rebuild_spatial_caches(const ComputationDevice &) { chk(mc_build_neighbors(ctx_)); }

    // md.minimize_energy(dev, max_iters, None) (ui/mol_editor.rs:375): returns the number of accepted descent moves
    int minimize_energy(const ComputationDevice &, int max_iters) {
        int accepted = 0;
        double e0 = 0, e1 = 0;
        chk(mc_minimize_energy(ctx_, max_iters, &accepted, &e0, &e1));
        return accepted;
    }

    // compute_energy_snapshot: one force / energy evaluation on the current positions (src/md/mod.rs:1036)
    SnapshotEnergyData energy_snapshot(const ComputationDevice &) {
        chk(mc_compute_forces(ctx_));
        mc_energy e;
        chk(mc_get_energy(ctx_, &e));
        SnapshotEnergyData s;
        s.energy_potential = e.energy_potential;
        s.energy_potential_nonbonded = e.energy_potential_nonbonded;
        s.energy_potential_bonded = e.energy_potential_bonded;
        s.energy_kinetic = e.energy_kinetic;
        s.temperature = e.temperature;
        s.volume = e.volume;
        s.density = e.density;
        double p = 0.0;
        if (cell.periodic && mc_get_pressure(ctx_, &p, nullptr) == MC_OK) s.pressure = p;
        return s;
    }

    // Snapshot hand-off without stalling the integrator (the queue of src/md/mod.rs:118-152): begin() stages the
    // current positions on the device and starts the copy into `out` (original atom order; page-locked memory
    // overlaps with the next steps), wait() blocks until the oldest outstanding snapshot has landed.
    void snapshot_begin(std::vector<mc_float4> &out) {
        out.resize(atoms.size());
        chk(mc_snapshot_begin(ctx_, out.data(), nullptr, nullptr));
    }
    void snapshot_begin(std::vector<mc_float4> &out, std::vector<mc_float4> &out_vel) {  // + Snapshot.atom_velocities
        out.resize(atoms.size());
        out_vel.resize(atoms.size());
        chk(mc_snapshot_begin_pv(ctx_, out.data(), out_vel.data(), nullptr, nullptr));
    }
    void snapshot_wait() { chk(mc_snapshot_wait(ctx_)); }

    // the box after barostat scaling (md.cell, properties/crystal.rs:633)
    SimBox current_box() {
        float lo[3], hi[3];
        chk(mc_get_box(ctx_, lo, hi));
        SimBox b = cell;
        b.bounds_low = {lo[0], lo[1], lo[2]};
        b.bounds_high = {hi[0], hi[1], hi[2]};
        return b;
    }

    // Make positions / velocities / forces host-visible on demand (the shrinking-box workflow reads
    // atom.force every chunk, properties/sol_shrinking_box.rs:776-789) -- never per step.
    void sync_atoms(bool with_forces = false) {
        const size_t n = atoms.size();
        std::vector<mc_float4> buf(n);
        chk(mc_get_positions(ctx_, buf.data()));
        for (size_t i = 0; i < n; ++i) atoms[i].posit = {buf[i].x, buf[i].y, buf[i].z};
        chk(mc_get_velocities(ctx_, buf.data()));
        for (size_t i = 0; i < n; ++i) atoms[i].vel = {buf[i].x, buf[i].y, buf[i].z};
        if (with_forces) {
            chk(mc_compute_forces(ctx_));
            chk(mc_get_forces(ctx_, buf.data()));
            for (size_t i = 0; i < n; ++i) { atoms[i].force = {buf[i].x, buf[i].y, buf[i].z}; atoms[i].energy_row = buf[i].w; }
        }
    }

    mc_ctx *raw() { return ctx_; }

  private:
    MdState() = default;
    mc_ctx *ctx_ = nullptr;
    void release() {
        if (ctx_) mc_destroy(ctx_);
        ctx_ = nullptr;
    }
    void chk(int rc) { chk(rc, ctx_); }
    void chk(int rc, mc_ctx *c) {
        if (rc < MC_OK) throw ParamError(rc, mc_last_error(c));  // positive codes (MC_W_*) are warnings: the call completed
    }
};

}  // namespace molchanica
