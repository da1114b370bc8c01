// host_mirror_smoke.cpp -- drives include/molchanica_md.hpp the way the reference drives `dynamics`:
// MdState::create -> energy_snapshot (two-body known answers) -> step -> sync_atoms.  Built by
// __graft_entry__.build(), run on the GPU box by tests/test_gpu_cpp_host.py.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "molchanica_md.hpp"

using namespace molchanica;

static int fail(const char *what, double got, double want) {
    std::printf("FAIL %s: got %.9g want %.9g\n", what, got, want);
    return 1;
}

// --npt adds the part written after the last hardware run of round 1 (thermostat + barostat + drift removal, pressure,
// snapshots with velocities): tests/test_library_on_host.py runs it against the host build, tests/newpaths_md.py on a GPU.
int main(int argc, char **argv) {
    const bool with_npt = argc > 1 && std::strcmp(argv[1], "--npt") == 0;
    ComputationDevice dev{0};
    try {
        // two argon atoms at the LJ minimum: E = -eps, F = 0; then at r = sigma: E = 0, |F| = 24 eps / sigma
        const float sigma = 3.405f, eps = 0.2381f;
        MdConfig cfg;
        cfg.lj_cutoff = cfg.coulomb_cutoff = 12.0f;
        cfg.skin = 1.0f;
        cfg.coulomb_mode = CoulombMode::None;
        cfg.sim_box.periodic = false;
        MdSystem sys;
        sys.n_lj_types = 1;
        sys.lj_sigma_eps = {sigma, eps};
        sys.atoms.resize(2);
        sys.atoms[0].mass = sys.atoms[1].mass = 39.948f;
        sys.atoms[1].posit.x = std::pow(2.0f, 1.0f / 6.0f) * sigma;
        {
            MdState md = MdState::create(dev, cfg, sys);
            SnapshotEnergyData e = md.energy_snapshot(dev);
            if (std::fabs(e.energy_potential_nonbonded + eps) > 1e-6) return fail("E(r_min)", e.energy_potential_nonbonded, -eps);
            md.sync_atoms(true);
            if (std::fabs(md.atoms[0].force.x) > 1e-5) return fail("F(r_min)", md.atoms[0].force.x, 0.0);
        }
        sys.atoms[1].posit.x = sigma;
        {
            MdState md = MdState::create(dev, cfg, sys);
            SnapshotEnergyData e = md.energy_snapshot(dev);
            if (std::fabs(e.energy_potential_nonbonded) > 1e-6) return fail("E(sigma)", e.energy_potential_nonbonded, 0.0);
            md.sync_atoms(true);
            const double want = 24.0 * eps / sigma;
            if (std::fabs(std::fabs(md.atoms[0].force.x) - want) > 1e-5 * want) return fail("|F|(sigma)", md.atoms[0].force.x, want);
            if (!(md.atoms[0].force.x < 0 && md.atoms[1].force.x > 0)) return fail("repulsion sign", md.atoms[0].force.x, -want);
            // step(dev, dt, None) x 10 and step(dev, dt, Some(forces)): the atoms must fly apart, momentum conserved
            md.step(dev, 0.002f, std::nullopt, 10);
            md.step(dev, 0.002f, std::vector<Vec3F32>(2));
            md.sync_atoms();
            if (!(md.atoms[1].posit.x - md.atoms[0].posit.x > sigma)) return fail("separation", md.atoms[1].posit.x - md.atoms[0].posit.x, sigma);
            if (std::fabs(md.atoms[0].vel.x + md.atoms[1].vel.x) > 1e-5) return fail("momentum", md.atoms[0].vel.x + md.atoms[1].vel.x, 0.0);
        }
        // NPT the way properties/crystal.rs:306-316 configures it: thermostat + zero_com_drift + barostat, on a small
        // periodic argon lattice; the box must shrink towards a target far above the starting pressure
        if (with_npt) {
            MdConfig npt = cfg;
            const int m = 6;
            const float a = 3.6f, L = a * m;
            npt.lj_cutoff = npt.coulomb_cutoff = 8.5f;
            npt.skin = 0.5f;
            npt.sim_box.periodic = true;
            npt.sim_box.bounds_low = {0, 0, 0};
            npt.sim_box.bounds_high = {L, L, L};
            npt.thermostat_tau = 0.1f;
            npt.temp_target = 90.0f;
            npt.zero_com_drift = true;
            npt.has_barostat = true;
            npt.barostat_cfg.pressure_target = 20000.0f;
            npt.barostat_cfg.tau = 0.5f;
            npt.barostat_cfg.compressibility = 1e-4f;
            MdSystem lat;
            lat.n_lj_types = 1;
            lat.lj_sigma_eps = {sigma, eps};
            for (int i = 0; i < m; ++i)
                for (int j = 0; j < m; ++j)
                    for (int k = 0; k < m; ++k) {
                        AtomDynamics at;
                        at.mass = 39.948f;
                        at.posit = {(i + 0.5f) * a + 0.01f * ((i * 7 + j * 3 + k) % 5), (j + 0.5f) * a, (k + 0.5f) * a};
                        at.vel = {0.5f * ((i + j) % 3 - 1), 0.5f * ((j + k) % 3 - 1), 0.5f * ((k + i) % 3 - 1)};
                        lat.atoms.push_back(at);
                    }
            MdState md = MdState::create(dev, npt, lat);
            md.step(dev, 0.002f, std::nullopt, 100);
            const SimBox b = md.current_box();
            if (!(b.bounds_high.x < L - 1e-3f)) return fail("barostat shrinks the box", b.bounds_high.x, L);
            SnapshotEnergyData e = md.energy_snapshot(dev);
            if (!(e.pressure != 0.0 && std::isfinite(e.pressure))) return fail("pressure", e.pressure, 1.0);
            std::vector<mc_float4> px, pv;
            md.snapshot_begin(px, pv);
            md.snapshot_wait();
            if (px.size() != lat.atoms.size() || !(pv[0].w > 0.f)) return fail("snapshot with velocities", (double)px.size(), (double)lat.atoms.size());
        }
        // error behaviour: a bad configuration surfaces as ParamError, never as a crash or a fallback
        bool threw = false;
        try {
            MdConfig bad = cfg;
            bad.lj_cutoff = -1.0f;
            MdState md = MdState::create(dev, bad, sys);
        } catch (const ParamError &) { threw = true; }
        if (!threw) return fail("ParamError", 0, 1);
    } catch (const ParamError &e) {
        std::printf("FAIL ParamError %d: %s\n", e.code, e.what());
        return 2;
    }
    std::printf("host mirror ok\n");
    return 0;
}
