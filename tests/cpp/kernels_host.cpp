// Test infrastructure: the kernel sources of bonded.cu, settle.cu, thermostat.cu and pme.cu compiled UNCHANGED for
// the host through tests/cpp/shim/cuda_runtime.h and run one thread at a time, so that their addressing (slot maps,
// minimum images, grid indices, atomics, energy reductions) is exercised on a machine without a GPU.
// tests/test_kernels_on_host.py compares the results with the fp64 oracles.  Not part of the product library.
#define MC_HOST_SHIM 1
#include "shim/cuda_runtime.h"

#include "../../molchanica_b200/csrc/bonded.cu"
#include "../../molchanica_b200/csrc/settle.cu"
#include "../../molchanica_b200/csrc/thermostat.cu"
#include "../../molchanica_b200/csrc/pme.cu"
#include "../../molchanica_b200/csrc/group_energy.cu"
#include "../../molchanica_b200/csrc/dock_filter.cu"

#define FOR_THREADS(n) gridDim.x = (unsigned)(n); for (blockIdx.x = 0; blockIdx.x < (unsigned)(n); ++blockIdx.x)

static NbParams nb(const float *ext, int periodic, float alpha) {
    NbParams p;
    memset(&p, 0, sizeof(p));
    for (int a = 0; a < 3; ++a) { p.ext[a] = periodic ? ext[a] : 1.f; p.inv_ext[a] = periodic ? 1.f / ext[a] : 1.f; }
    p.periodic = periodic;
    p.alpha = alpha;
    return p;
}

extern "C" {

void host_bonded(int n_bonds, const int2 *bonds, const float2 *kr0, int n_angles, const int4 *angles, const float2 *kt0, int n_dih,
                 const int4 *dih, const float4 *prm, const int *slot_of_orig, const float4 *xyzq, const float *ext, int periodic,
                 float4 *force, double *energy3) {
    BondedTerms t;
    t.n_bonds = n_bonds; t.n_angles = n_angles; t.n_dihedrals = n_dih;
    t.bonds = bonds; t.bond_kr0 = kr0; t.angles = angles; t.angle_kt0 = kt0; t.dihedrals = dih; t.dihedral_prm = prm;
    const int n = n_bonds + n_angles + n_dih;
    // round the thread count up like the launch does (threads past the last term must do nothing)
    FOR_THREADS(((n + 127) / 128) * 128) bonded_kernel(t, slot_of_orig, xyzq, nb(ext, periodic, 0.f), force, energy3, 1);
}

void host_settle(int n_w, const int4 *waters, const int *slot_of_orig, float4 *xyzq, float4 *vel, float m_o, float m_h, float d_oh,
                 float d_hh, const float *ext, int periodic, float dt) {
    const SettleParams sp = mc_settle_params(m_o, m_h, d_oh, d_hh);
    FOR_THREADS(n_w + 3) settle_kernel(n_w, waters, slot_of_orig, xyzq, vel, sp, nb(ext, periodic, 0.f), dt, nullptr);
}

int host_shake_h(int n_c, const int4 *clusters, const float *dist, const int *slot_of_orig, float4 *xyzq, float4 *vel, const float *ext,
                 int periodic, float dt, float tol) {
    int bad = 0;
    FOR_THREADS(n_c + 3) shake_h_kernel(n_c, clusters, dist, slot_of_orig, xyzq, vel, nb(ext, periodic, 0.f), dt, tol, &bad, nullptr);
    return bad;
}

void host_vsite_construct(int n_v, const int4 *sites, const int *slot_of_orig, float4 *xyzq, float a, float b, const float *ext,
                          int periodic) {
    FOR_THREADS(n_v + 3) vsite_construct_kernel(n_v, sites, slot_of_orig, xyzq, a, b, nb(ext, periodic, 0.f));
}

void host_vsite_spread(int n_v, const int4 *sites, const int *slot_of_orig, float4 *force, float a, float b) {
    FOR_THREADS(n_v + 3) vsite_spread_kernel(n_v, sites, slot_of_orig, force, a, b);
}

void host_langevin(int n, float4 *vel, const int *orig, const uint8_t *flags, float c1, float c2, float kT, uint64_t seed, uint64_t step) {
    FOR_THREADS(n + 5) langevin_ou_kernel(n, vel, orig, flags, c1, c2, kT, seed, step);
}

void host_dock_filter(int n_rs, const float4 *rec_sample, int n_ls, const float4 *lig_sample, const float *anchor, float limit, int n_poses,
                      const float *poses, uint8_t *keep) {
    FOR_THREADS(((n_poses + 127) / 128) * 128)
    dock_filter_kernel(n_rs, rec_sample, n_ls, lig_sample, make_float3(anchor[0], anchor[1], anchor[2]), limit, n_poses, poses, keep);
}

void host_zero_velocities(int n, float4 *vel) { FOR_THREADS(n + 11) zero_velocities_kernel(n, vel); }

void host_csvr(int n, float4 *vel, const double *red3, double kT, double c, double dof_removed, uint64_t seed, uint64_t step,
               float *lambda) {
    FOR_THREADS(3) csvr_lambda_kernel(red3, kT, c, dof_removed, seed, step, lambda);
    FOR_THREADS(n + 9) csvr_scale_kernel(n, vel, lambda);
}

double host_between_mols(int n, const float4 *xyzq, const uint16_t *type, const int *orig, const uint16_t *mol_of_orig,
                         const uint32_t *nbr_start, const uint32_t *nbr_count, const uint32_t *nbr_list, const float2 *ljtab, int n_types,
                         const float *ext, int periodic, float rc_lj, float rc_q, int lj_on, int coul_mode, float alpha) {
    NbParams p = nb(ext, periodic, alpha);
    p.rc2_lj = rc_lj * rc_lj; p.rc2_q = rc_q * rc_q; p.n_types = n_types;
    double e = 0.0;
    FOR_THREADS(((n + 127) / 128) * 128) between_mols_kernel(n, 0, xyzq, type, orig, mol_of_orig, nbr_start, nbr_count, nbr_list, ljtab, p,
                                                             lj_on, coul_mode, &e);
    return e;
}

static PmeGeom geom(const int *K, const float *lo, const float *ext) {
    PmeGeom g;
    for (int a = 0; a < 3; ++a) { g.K[a] = K[a]; g.lo[a] = lo[a]; g.inv_ext[a] = 1.0f / ext[a]; g.scale[a] = (float)K[a] / ext[a]; }
    return g;
}

void host_pme_spread(int n, const float4 *xyzq, const int *K, const float *lo, const float *ext, float *grid) {
    FOR_THREADS(n + 7) pme_spread_kernel(n, xyzq, geom(K, lo, ext), grid);
}

// cgrid: K1 x K2 x (K3/2+1) complex, multiplied in place; energy[0] accumulated
void host_pme_convolve(const int *K, float2 *cgrid, const float *ext, float alpha, double *energy) {
    float *bm[3];
    for (int a = 0; a < 3; ++a) {
        bm[a] = new float[K[a]];
        for (int m = 0; m < K[a]; ++m) bm[a][m] = (float)mc_pme_bmod4(m, K[a]);
    }
    const double vol = (double)ext[0] * ext[1] * ext[2];
    const float pi = 3.14159265358979f;
    // a grid-stride loop: a few "threads" are enough to exercise the striding
    FOR_THREADS(37) pme_convolve_kernel(K[0], K[1], K[2], cgrid, bm[0], bm[1], bm[2], 1.0f / ext[0], 1.0f / ext[1], 1.0f / ext[2],
                                        (float)(1.0 / (3.14159265358979323846 * vol)), pi * pi / (alpha * alpha), energy, 1);
    for (int a = 0; a < 3; ++a) delete[] bm[a];
}

void host_pme_gather(int n, const float4 *xyzq, const int *K, const float *lo, const float *ext, const float *grid, float4 *force) {
    FOR_THREADS(n + 7) pme_gather_kernel(n, xyzq, geom(K, lo, ext), grid, force);
}

void host_pme_excl(int n, const float4 *xyzq, const int *orig, const int *slot_of_orig, const int32_t *excl_start, const int32_t *excl_idx,
                   const float *ext, int periodic, float alpha, float4 *force, double *energy) {
    FOR_THREADS(n + 7) pme_excl_kernel(n, xyzq, orig, slot_of_orig, excl_start, excl_idx, nb(ext, periodic, alpha), force, energy, 1);
}
}
