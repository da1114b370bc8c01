/*
 * molchanica_md.h -- C ABI of libmolchanica_md.so, the B200 (sm_100a) force-evaluation engine
 * for Molchanica's src/md hot path and the src/docking pose-energy scan.
 *
 * This library replaces the reference's device boundary -- the PTX module built from
 * src/cuda/{cuda,util}.cu by build.rs:10-16 and loaded in src/util.rs:1072-1119 -- and the
 * force / neighbour-list / integrator loops that the `dynamics` crate runs behind
 * `MdState::step(dev, dt, ext_forces)` (call sites: reference src/md/mod.rs:716,748,
 * src/mol_editor/mod.rs:388, src/mol_alignment.rs:346).  A Rust `extern "C"` binding for it is
 * shown in INTEGRATION.md; include/molchanica_md.hpp mirrors the `dynamics` API surface on top
 * of it for C++ hosts.
 *
 * Conventions
 *   - plain pointers and sizes only; the caller owns every host buffer; the library copies in
 *     `mc_set_*`, owns all device memory until `mc_destroy`, and fills caller-allocated buffers
 *     in `mc_get_*` (the ownership model of clone_htod / clone_dtoh, src/reflection.rs:146-214)
 *   - every function returns 0 on success, a negative MC_E_* code on failure (a positive MC_W_* code is a warning); the message is
 *     available from mc_last_error(); nothing throws or aborts across the ABI
 *   - there is NO CPU fallback: without a usable CUDA device mc_create fails (the reference
 *     degrades to ComputationDevice::Cpu, src/util.rs:1065-1070; BASELINE.json forbids that here)
 *   - a handle is not re-entrant; distinct handles may be used from distinct threads; all work
 *     of a handle runs on its own non-blocking stream
 *   - units: Angstrom, ps, amu, kcal/mol; charges pre-multiplied by sqrt(332.0522); atom ids are
 *     the caller's ("original") ids everywhere on this boundary, whatever order the engine keeps
 *     internally
 */
#ifndef MOLCHANICA_MD_H
#define MOLCHANICA_MD_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MC_ABI_VERSION 2

#define MC_OK 0
#define MC_E_INVALID (-1)   /* bad argument / call order                      */
#define MC_E_CUDA (-2)      /* CUDA runtime error (see mc_last_error)          */
#define MC_E_NODEVICE (-3)  /* no usable sm_100 device -- there is no fallback */
#define MC_E_CAPACITY (-4)  /* caller buffer too small                         */
#define MC_E_COMM (-5)      /* NCCL / peer-exchange error                      */
#define MC_W_STALE_LIST 1   /* warning (the call completed): an atom moved more than skin/2 between two list builds of a
                             * fixed / decomposed rebuild schedule, so pairs inside the cutoff may have been missing from the
                             * forces of the last steps; the list is rebuilt before the next evaluation (all ranks agree) */

#define MC_COULOMB_NONE 0   /* q ignored                                                  */
#define MC_COULOMB_PLAIN 1  /* q_i q_j /(r^2 + 1e-6), truncated at rc_q (src/cuda/util.cu:54-63) */
#define MC_COULOMB_ERFC 2   /* Ewald real-space erfc(alpha r)/r (util.cu:15-18 INV_SQRT_PI)       */

#define MC_THERMOSTAT_NONE 0
#define MC_THERMOSTAT_LANGEVIN 1  /* Langevin dynamics (one of the reference's thermostats, README.md:238)      */
#define MC_THERMOSTAT_CSVR 2      /* canonical-sampling velocity rescaling, the reference's default for NVT      */

#define MC_FLAG_STATIC 1u   /* AtomDynamics.static_ : exerts forces, never moves (src/md/mod.rs:843-852) */

typedef struct mc_ctx mc_ctx;

/* One float4 per atom everywhere: {x, y, z, q} positions+charge, {vx, vy, vz, 1/m} velocities,
 * {fx, fy, fz, e_i} forces + per-atom pair-energy row sum. */
typedef struct { float x, y, z, w; } mc_float4;

/* SnapshotEnergyData subset (reference src/md/mod.rs:1242-1245, ui/panels/md_viewer.rs:202-256) */
typedef struct {
    double energy_potential;            /* nonbonded + bonded                                    */
    double energy_potential_nonbonded;  /* LJ + Coulomb + scaled 1-4, kcal/mol                   */
    double energy_potential_bonded;     /* bonds + angles + dihedrals set with mc_set_bonds ...  */
    double energy_kinetic;              /* sum 1/2 m v^2 / 418.4, kcal/mol                       */
    double temperature;                 /* 2 KE / (3 N_mobile k_B), K                            */
    double energy_bond, energy_angle, energy_dihedral;  /* parts of energy_potential_bonded      */
    double volume;                      /* A^3 (periodic boxes, else 0)                          */
    double density;                     /* g/cm^3 from the atoms' masses (periodic boxes)        */
    double energy_pme;                  /* SPME reciprocal + self + excluded-pair correction (part of nonbonded) */
} mc_energy;

/* Counters for the caller / benchmarks.  The *_ms_sum fields accumulate CUDA-event durations
 * measured on the handle's stream while option "profiling" is on. */
typedef struct {
    int64_t n_atoms;             /* atoms owned by this handle                          */
    int64_t n_ghosts;            /* ghost atoms held (domain decomposition), else 0     */
    int64_t n_pairs_listed;      /* full-list entries of the current Verlet list        */
    int64_t n_rebuilds;          /* list builds since mc_create                         */
    int64_t n_steps;             /* integrator steps since mc_create                    */
    int64_t n_kernel_launches;   /* kernels launched by this handle since mc_create     */
    int64_t n_cells[3];          /* current cell grid (periodic boxes)                  */
    double  pair_ms_sum;      int64_t pair_launches_timed;
    double  build_ms_sum;     int64_t builds_timed;
    double  integrate_ms_sum; int64_t integrate_launches_timed;
    double  halo_ms_sum;      int64_t halos_timed;
    int64_t n_list_violations;   /* fixed-schedule rebuilds only: times an atom outran skin/2 between builds */
    int64_t list_bytes;          /* bytes of the neighbour-list index stream one force evaluation reads (rows incl. padding) */
    int64_t ext_upload_bytes;    /* host-to-device bytes the last mc_step moved for its external forces (this rank) */
} mc_stats;

/* ---- lifetime -------------------------------------------------------------------------- */

/* Replaces get_computation_device (src/util.rs:1072-1119): binds CUDA device `device`, creates
 * the stream.  Fails with MC_E_NODEVICE when no device is usable. */
int mc_create(int device, mc_ctx **out);
int mc_destroy(mc_ctx *ctx);
/* Message of the last failure on this handle (or of the last failed mc_create when ctx==NULL). */
const char *mc_last_error(const mc_ctx *ctx);
int mc_abi_version(void);
/* sizeof(mc_energy) / sizeof(mc_stats) as this library was compiled: a binding whose struct mirrors differ must not
 * call mc_get_energy / mc_get_stats (they would write past the caller's buffer). */
int mc_struct_sizes(int *energy_bytes, int *stats_bytes);

/* ---- system definition (what MdState::new hands over, src/md/mod.rs:641-693) ------------ */

/* SimBox {bounds_low, bounds_high} (properties/sol_shrinking_box.rs:600-603).  periodic = 0
 * means vacuum (Solvent::None, src/md/mod.rs:784): the box is then ignored and the cell grid
 * follows the atoms' bounding box. */
int mc_set_box(mc_ctx *ctx, const float lo[3], const float hi[3], int periodic);

/* n atoms; type may be NULL (all 0), vel may be NULL (zero velocities, unit mass), flags may be
 * NULL.  Resets the neighbour list. */
int mc_set_atoms(mc_ctx *ctx, int64_t n, const mc_float4 *xyzq, const uint16_t *type,
                 const mc_float4 *vel_invmass, const uint8_t *flags);

/* T x T table of (sigma_ij [A], eps_ij [kcal/mol]) pairs, row-major -- replaces the dense
 * [N_tgt x N_src] sigma/eps arrays of lj_force_kernel (src/cuda/cuda.cu:73-102). */
int mc_set_lj_table(mc_ctx *ctx, int n_types, const float *sigma_eps);

/* Excluded partners per atom (1-2, 1-3 and 1-4), CSR: start[n+1], idx[start[n]]. NULL clears. */
int mc_set_exclusions(mc_ctx *ctx, const int32_t *start, const int32_t *idx);

/* Amber 1-4 pairs (pairs[2*m]) evaluated without cutoff, LJ x scale_lj, Coulomb x scale_q. */
int mc_set_pairs14(mc_ctx *ctx, int64_t m, const int32_t *pairs, float scale_lj, float scale_q);

/* Bonded terms, Amber functional forms, evaluated on the device in the same force evaluation
 * as the nonbonded terms.  Atom ids are the caller's; call after mc_set_atoms (which clears them); m = 0 clears one kind.
 * Decomposed handles: every rank is given the WHOLE term list and evaluates the terms that touch an atom it owns (the
 * partners are ghosts: a term must not reach further than cutoff + skin, MC_E_INVALID otherwise); mc_energy's bonded
 * energies are summed over the ranks.  Exclusions / 1-4 pairs that go with the bonds are set separately above.
 *   bonds:     pairs[2m],   k_r0[2m]       E = k (r - r0)^2                    kcal/mol/A^2, A
 *   angles:    triples[3m] (vertex second), k_theta0[2m]  E = k (theta - theta0)^2   kcal/mol/rad^2, rad
 *   dihedrals: quads[4m],   pk_n_phase[3m] E = pk (1 + cos(n phi - phase))     kcal/mol, -, rad (IUPAC phi) */
int mc_set_bonds(mc_ctx *ctx, int64_t m, const int32_t *pairs, const float *k_r0);
int mc_set_angles(mc_ctx *ctx, int64_t m, const int32_t *triples, const float *k_theta0);
int mc_set_dihedrals(mc_ctx *ctx, int64_t m, const int32_t *quads, const float *pk_n_phase);

/* Thermostat.  MC_THERMOSTAT_LANGEVIN: after the drift (and constraints) of every step the
 * velocities get the Ornstein-Uhlenbeck update v <- c1 v + sqrt(1 - c1^2) sqrt(kT/m) xi, c1 = exp(-gamma dt)
 * (splitting B A O B).  The noise is Philox4x32-10 keyed by (seed, atom id, step count since this call): it does
 * not depend on the engine's internal atom order.  Static atoms are left alone.
 * MC_THERMOSTAT_CSVR (Bussi-Donadio-Parrinello): at the same point of the step all velocities are scaled by one
 * stochastic factor computed on the device from the kinetic energy; gamma_per_ps is then 1 / tau.  The factor is a
 * function of (seed, step, kinetic energy) only.  Both work on decomposed handles: Langevin draws the same noise on any
 * decomposition, CSVR all-reduces the kinetic energy (24 bytes per step, on the stream). */
int mc_set_thermostat(mc_ctx *ctx, int kind, float temperature_k, float gamma_per_ps, uint64_t seed);

/* SPME reciprocal space (the reference's electrostatics, README.md:240): with coulomb_mode =
 * MC_COULOMB_ERFC the pair kernel evaluates erfc(alpha r)/r; a k1 x k2 x k3 grid (order-4 B-splines, cuFFT) adds
 * the reciprocal sum, the self term and the erf(alpha r)/r correction of the excluded pairs to every force
 * evaluation.  Periodic boxes, single-GPU handles; 0, 0, 0 switches it off.  About one grid point per Angstrom with
 * alpha = 0.35 gives forces within ~1e-3 of the exact Ewald sum. */
int mc_set_pme(mc_ctx *ctx, int k1, int k2, int k3);
/* Host-only helper (no GPU needed): the Ewald alpha with erfc(alpha rc) / rc <= tol and a grid K_a >= 2 alpha L_a /
 * (3 tol^(1/5)) rounded up to a product of 2, 3, 5, 7 -- what a caller would pass to mc_set_cutoffs / mc_set_pme. */
int mc_pme_suggest(float rc, float tol, const float box_ext[3], float *alpha, int32_t grid[3]);

/* Rigid three-site waters (the reference keeps its water rigid with SETTLE, README.md:239):
 * triples[3m] = (O, H1, H2) atom ids; after every drift of mc_step the molecules are put back onto the
 * triangle (d_oh, d_oh, d_hh) with the analytic SETTLE and their velocities corrected by the position change
 * / dt.  Intramolecular pairs must be excluded (mc_set_exclusions); masses are those of mc_set_atoms.
 * mc_energy.temperature then counts 3 degrees of freedom less per molecule.  Single-GPU handles; m = 0 clears. */
int mc_set_rigid_waters(mc_ctx *ctx, int64_t m, const int32_t *triples, float d_oh, float d_hh, float m_o, float m_h);

/* Constraints on bonds to hydrogen (SHAKE; the reference's "hydrogen constraints" at 2 fs, ui/panels/md.rs:362-371):
 * clusters[4m] = (heavy atom, h1, h2, h3) with -1 for unused hydrogen slots, lengths[3m] the constrained distances;
 * no atom may appear in two clusters (one thread owns a cluster).
 * Applied after every drift like mc_set_rigid_waters (which handles water); a cluster that does not converge in 64
 * sweeps makes mc_step return MC_E_INVALID.  mc_energy.temperature counts one degree of freedom less per constraint.
 * Single-GPU handles; m = 0 clears. */
int mc_set_hbond_constraints(mc_ctx *ctx, int64_t m, const int32_t *clusters, const float *lengths);

/* Virtual sites of four-site water (OPC / TIP4P; the reference's md.water {o, h0, h1, m},
 * properties/sol_shrinking_box.rs:605-613): quads[4m] = (M, O, H1, H2) atom ids, M = O + a (H1 - O) + b (H2 - O).
 * M is an atom of the system with inverse mass 0 and MC_FLAG_STATIC (it carries the charge); after every drift
 * (and SETTLE) of mc_step it is placed again, and after every force evaluation its force is handed to the three
 * parents (a, b and 1 - a - b) and zeroed.  The caller places M consistently in mc_set_atoms and excludes the
 * intramolecular pairs.  Single-GPU handles; m = 0 clears. */
int mc_set_virtual_sites(mc_ctx *ctx, int64_t m, const int32_t *quads, float a, float b);

/* cfg.lj_cutoff / cfg.coulomb_cutoff (ui/panels/md.rs:260-261), Verlet skin, Coulomb form. */
int mc_set_cutoffs(mc_ctx *ctx, float rc_lj, float rc_q, float skin, int coulomb_mode, float alpha);

/* MdOverrides.lj_disabled / coulomb_disabled (src/md/mod.rs:671-686). */
int mc_set_overrides(mc_ctx *ctx, int lj_disabled, int coulomb_disabled);

/* Tuning / instrumentation knobs: "pair_lanes" (4, 8, 16, 32 lanes per list row),
 * "profiling" (0/1: bracket the hot kernels with CUDA events), "rebuild_every" (0 = rebuild on
 * the displacement criterion, k > 0 = every k steps like GROMACS nstlist), "halo_fused" (decomposed
 * handles; 1 = peer-memory ghost exchange inside the step kernels (default), 0 = NCCL send / recv),
 * "dd_migrate" (decomposed handles; 1 = rebuilds exchange boundary layers with the two neighbour ranks only
 * (default), 0 = all-gather of the whole system), "subcell_sort" (Morton sub-cell code in the sort key),
 * "profile_every" (k: inside mc_step only every k-th step's kernels are bracketed with events),
 * "zero_com_drift" (k: remove the velocity of the centre of mass of the mobile atoms every k steps, MdConfig.zero_com_drift
 * of the reference; 0 = never (default)),
 * "defer_tail" (1: a single-GPU mc_step with ext_forces returns after its last drift and finishes that
 * step -- force evaluation, second half kick -- under the upload of the next call's array, or as soon as anything
 * but positions is asked for; results are the same, only the time at which the work is done moves (confirmed on
 * hardware in round 2; works on decomposed handles as well); 0 (default) = finish every step inside its own call),
 * "fused_steps" (1 (default): small plain-NVE systems run all steps of a call in one cooperative launch, md_fused.cu),
 * A/B knobs kept with their measurements under profiles/: "pair_tile" (TMA-staged pair kernel), "pair_tile_stages",
 * "rows_interleave" (quad-interleaved index rows), "build_variant" (1 = tile_build_kernel, 2 = rows_build_kernel),
 * "rows_dense" (0 = dense systems build their list as at the start of round 2), "build_split" (slices per cell),
 * "rows_min_blocks", "row_stage_limit", "pair_uniform", "sync_rebuild", "lazy_sync", "early_tail", "tile_sweep". */
int mc_set_option(mc_ctx *ctx, const char *name, double value);

/* Replace positions (and optionally velocities) of the existing atoms, original order. */
int mc_set_positions(mc_ctx *ctx, const mc_float4 *xyzq);
int mc_set_velocities(mc_ctx *ctx, const mc_float4 *vel_invmass);

/* ---- the hot path ---------------------------------------------------------------------- */

/* md.rebuild_spatial_caches(dev) (properties/sol_shrinking_box.rs:632): wrap into the box, cell
 * ids, on-device radix sort + prefix scan, spatial reorder, Verlet list of radius
 * max(rc_lj, rc_q) + skin. */
int mc_build_neighbors(mc_ctx *ctx);

/* One nonbonded force evaluation on the current positions (rebuilds the list first when it is
 * stale).  The single-point path of compute_energy_snapshot (src/md/mod.rs:1036). */
int mc_compute_forces(mc_ctx *ctx);

/* MdState::step(dev, dt, ext) x n_steps (src/md/mod.rs:716,748): velocity Verlet, list rebuilt
 * on the device-side displacement criterion (> skin/2).  ext_forces: n x {fx,fy,fz} in original
 * order added at every evaluation, or NULL (mol_alignment.rs:346 passes Some(forces)). */
int mc_step(mc_ctx *ctx, float dt, int n_steps, const float *ext_forces);
/* Device time of the last mc_step call: CUDA events on the handle's stream around all of its
 * steps (kernels, rebuilds and any idle gaps between them). */
double mc_last_step_ms(mc_ctx *ctx);

/* md.minimize_energy(dev, max_iters, None) (ui/mol_editor.rs:375, mol_alignment.rs:356, properties/sol_shrinking_box.rs:962):
 * steepest descent along F/m with an adaptive step (accepted moves lengthen it by 1.2, rejected ones are undone and
 * halve it), at most max_iters trial moves, list rebuilt on the same displacement criterion as mc_step.  Velocities
 * are preserved.  Returns the number of accepted moves and the potential energy before / after.  Single-GPU handles,
 * no constraints set. */
int mc_minimize_energy(mc_ctx *ctx, int max_iters, int *iters_accepted, double *e_initial, double *e_final);

/* ---- read-back (caller-allocated, original atom order) ---------------------------------- */
int mc_get_positions(mc_ctx *ctx, mc_float4 *out);
int mc_get_velocities(mc_ctx *ctx, mc_float4 *out);
int mc_get_forces(mc_ctx *ctx, mc_float4 *out);
int mc_get_energy(mc_ctx *ctx, mc_energy *out);
int mc_get_stats(mc_ctx *ctx, mc_stats *out);
/* SnapshotEnergyData.pressure (reference ui/panels/md_viewer.rs:202-256): P = (2 KE + W) / 3V in bar and the virial
 * W = sum r_ij . f_ij in kcal/mol over nonbonded pairs inside the cutoffs, scaled 1-4 pairs, bonded terms and the SPME
 * reciprocal sum with its excluded-pair correction.  One extra pass over the neighbour list, on demand only.
 * Periodic boxes; on a decomposed handle a collective call (all ranks, one all-reduce).  With rigid waters / constrained bonds the virial of the constraint forces of the LAST
 * step is included (mass x constraint displacement / dt^2 on the old positions), so at least one step must have been taken.
 * Either output may be NULL. */
int mc_get_pressure(mc_ctx *ctx, double *pressure_bar, double *virial);

/* MdConfig.barostat_cfg = BarostatCfg{pressure_target [bar], tau [ps]} (reference ui/panels/md.rs:517-556,
 * properties/crystal.rs:312, water_sol.rs:146).  Every every_n_steps steps the instantaneous pressure is measured and
 * box + coordinates are scaled about the origin: MC_BAROSTAT_BERENDSEN (weak coupling) or MC_BAROSTAT_CRESCALE
 * (stochastic cell rescaling, Bernetti & Bussi 2020, samples NPT together with a canonical thermostat -- call
 * mc_set_thermostat first, its temperature sets the noise; velocities are scaled by 1/mu).  compressibility_per_bar:
 * isothermal compressibility (water 4.5e-5).  The `dynamics` crate's own barostat algorithm is not in the reference
 * tree [EXTERNAL]; both kinds relax the box towards pressure_target with time constant tau.  Periodic single-GPU
 * handles; mc_get_box reports the current box. */
#define MC_BAROSTAT_NONE 0
#define MC_BAROSTAT_BERENDSEN 1
#define MC_BAROSTAT_CRESCALE 2
int mc_set_barostat(mc_ctx *ctx, int kind, float pressure_bar, float tau_ps, float compressibility_per_bar, int every_n_steps,
                    uint64_t seed);
int mc_get_box(mc_ctx *ctx, float lo[3], float hi[3]);

/* SnapshotEnergyData.energy_potential_between_mols (src/md/mod.rs:1242-1245): mol_id[n] assigns every atom to a molecule;
 * mc_get_energy_between_mols sums the nonbonded pair energies (LJ + the Coulomb form in force, within the cutoffs)
 * over the listed pairs whose atoms belong to different molecules.  On demand, not on the step path; excluded
 * pairs and the reciprocal part of SPME are not in it.  On a decomposed handle mol_id covers all n_global atoms and
 * the get is a collective call (one all-reduce); NULL clears the ids. */
int mc_set_molecule_ids(mc_ctx *ctx, const uint16_t *mol_id);
int mc_get_energy_between_mols(mc_ctx *ctx, double *out);

/* Asynchronous snapshot hand-off (the Snapshot queue of src/md/mod.rs:118-152): mc_snapshot_begin
 * stages the current positions on the device and starts their copy to `out_positions` on a second
 * stream, then returns; the caller may issue further mc_step calls at once.  mc_snapshot_wait blocks
 * until the OLDEST outstanding snapshot has landed (at most two may be in flight; host buffers should
 * be page-locked for the copy to overlap).  Single GPU: n_global entries in original order, out_ids
 * may be NULL.  Decomposed handle: the atoms this rank owns, *n_out of them, with their original
 * ids in out_ids (both buffers must hold the rank's capacity, see mc_comm_counts). */
int mc_snapshot_begin(mc_ctx *ctx, mc_float4 *out_positions, int32_t *out_ids, int64_t *n_out);
/* The same hand-off as packed x, y, z (Snapshot.atom_posits: Vec<Vec3F32>, src/md/trajectory.rs:160-204): 3 floats per atom.
 * *layout_epoch (may be NULL) counts the list builds: the ids a decomposed rank returns only change when it does.  It is read
 * and written: on entry the epoch whose ids the caller already holds (-1: none) -- out_ids is then only filled, and the ids
 * only travel, when the layout is another one -- on return the epoch of this snapshot.  (NULL: ids whenever out_ids is given.) */
int mc_snapshot_begin_xyz(mc_ctx *ctx, float *out_xyz, int32_t *out_ids, int64_t *n_out, int64_t *layout_epoch);
/* The same with velocities ({vx, vy, vz, 1/m}; Snapshot.atom_velocities, src/md/trajectory.rs:160-204).  Velocities are
 * only final once the step a pipelined mc_step may have left open is closed, so this call finishes it first. */
int mc_snapshot_begin_pv(mc_ctx *ctx, mc_float4 *out_positions, mc_float4 *out_velocities, int32_t *out_ids, int64_t *n_out);
int mc_snapshot_wait(mc_ctx *ctx);

/* Verlet list as CSR in original ids, rows ascending.  start: n+1 entries.  Two-call protocol:
 * idx == NULL or cap too small -> start[] is still filled, *total set, MC_E_CAPACITY returned
 * when idx != NULL. */
int mc_get_neighbors(mc_ctx *ctx, int64_t *start, int32_t *idx, int64_t cap, int64_t *total);

int mc_reset_timers(mc_ctx *ctx);

/* Times `reps` stand-alone launches of the pair-force kernel on the current state with CUDA
 * events on the handle's stream (3 untimed warm-ups first); the mean is returned by
 * mc_last_pair_kernel_ms.  flush_l2 != 0 writes a buffer twice the L2 size between launches. */
int mc_time_kernels(mc_ctx *ctx, int reps, int flush_l2);
double mc_last_pair_kernel_ms(mc_ctx *ctx);

/* ---- docking pose-energy scan (src/docking/legacy/mod.rs:210-383, :174-200) ------------- */

/* receptor: n_rec atoms (xyzq, type, hydrophobic flag); ligand: n_lig atoms in their reference
 * conformation + anchor point; ljtab: n_rec_types x n_lig_types (sigma, eps); poses: n_poses x
 * {ax, ay, az, qw, qx, qy, qz}.  out: n_poses x {score, vdw, hydrophobic, electrostatic,
 * coulomb_e}.  All host pointers; runs on the handle's stream and returns when done. */
int mc_dock_score(mc_ctx *ctx, int64_t n_rec, const mc_float4 *rec_xyzq, const uint16_t *rec_type,
                  const uint8_t *rec_hydrophobic, int64_t n_lig, const mc_float4 *lig_xyzq,
                  const uint16_t *lig_type, const uint8_t *lig_hydrophobic, const float lig_anchor[3],
                  int n_rec_types, int n_lig_types, const float *ljtab,
                  int64_t n_poses, const float *poses, float *out);
/* The same scan for a flexible ligand (ConformationType::AssignedTorsions, legacy/mod.rs:140-158): poses carry n_flex torsion
 * angles behind the 7 rigid numbers (mc_dock_make_poses_flex); before the rigid transform the atoms downstream of flexible
 * bond f (flex_mask[f * n_lig + a], mc_dock_flex_masks) are rotated by t_f about the axis a0 -> a1 through a1, bond after
 * bond in the order given (a later axis sees the atoms where the earlier rotations left them), in f64 like the transform. */
int mc_dock_score_flex(mc_ctx *ctx, int64_t n_rec, const mc_float4 *rec_xyzq, const uint16_t *rec_type, const uint8_t *rec_hydrophobic,
                       int64_t n_lig, const mc_float4 *lig_xyzq, const uint16_t *lig_type, const uint8_t *lig_hydrophobic,
                       const float lig_anchor[3], int n_rec_types, int n_lig_types, const float *ljtab, int n_flex,
                       const int32_t *flex_axis, const uint8_t *flex_mask, int64_t n_poses, const float *poses, float *out);
/* ---- pose set of the scan -- host side, usable without a GPU -------------------- */

/* make_posits_orientations + init_poses for a rigid ligand (src/docking/legacy/mod.rs:386-500): anchors on a
 * num_posits^3 grid of cell centres over the cube of half-width site_radius about site_center (x slowest), times
 * n_lats * n_lons * n_rolls orientations with n_lats = floor((num_orientations/2)^(1/3)), n_lons = n_rolls =
 * 2 n_lats (find_optimal_pose uses num_posits = 8, num_orientations = 60 -> 512 x 108 poses, :705-706).
 * out_poses: n x {ax, ay, az, qw, qx, qy, qz}, anchor-major.  out_poses == NULL only reports *n_out. */
int mc_dock_make_poses(const double site_center[3], double site_radius, int num_posits, int num_orientations,
                       float *out_poses, int64_t cap, int64_t *n_out);
int mc_dock_orientation_count(int num_orientations);
/* init_poses with flexible bonds (legacy/mod.rs:453-500, Torsion{bond, dihedral_angle} legacy/prep.rs:405-410): every rigid
 * pose times the cartesian product of angles_per_bond angles = linspace(0, TAU, angles_per_bond) per flexible bond (first
 * bond slowest).  out_poses: n x (7 + n_flex_bonds) floats {ax, ay, az, qw, qx, qy, qz, t_0 ..}.  n_flex_bonds <= MC_DOCK_MAX_FLEX. */
#define MC_DOCK_MAX_FLEX 12
int mc_dock_make_poses_flex(const double site_center[3], double site_radius, int num_posits, int num_orientations,
                            int n_flex_bonds, int angles_per_bond, float *out_poses, int64_t cap, int64_t *n_out);
/* The rotating side of every flexible bond: bonds[2 * n_bonds] atom pairs of the ligand, flex_bond_idx[n_flex_bonds] indices
 * into it; axis_out[2 f] = {a0, a1}, mask_out[f * n_lig + a] = 1 for the atoms downstream of a1 when the bond is cut.
 * MC_E_INVALID for a bond inside a ring. */
int mc_dock_flex_masks(int64_t n_lig, int64_t n_bonds, const int32_t *bonds, int n_flex_bonds, const int32_t *flex_bond_idx,
                       int32_t *axis_out, uint8_t *mask_out);
/* find_rec_atoms_near_site (legacy/prep.rs:506-532): receptor atoms within 1.4 x site_radius of the site centre
 * that are not hetero atoms; out_idx (capacity n_rec) may be NULL to count. */
int mc_dock_near_site(int64_t n_rec, const mc_float4 *rec_xyzq, const uint8_t *rec_hetero, const double site_center[3],
                      double site_radius, int32_t *out_idx, int64_t *n_out);
/* Clash pre-filter of process_poses (legacy/mod.rs:522-573): keep[p] = 0 when a sampled ligand carbon (index % 4 == 0)
 * of pose p lies within 1.1 x vdw_radius of a sampled receptor carbon (near-site index % 6 == 0).  rec_* describe
 * the near-site subset in its own order; the ligand is posed exactly as mc_dock_score poses it. */
int mc_dock_filter_poses(int64_t n_rec, const mc_float4 *rec_xyzq, const uint8_t *rec_is_carbon, int64_t n_lig,
                         const mc_float4 *lig_xyzq, const uint8_t *lig_is_carbon, const float lig_anchor[3],
                         float vdw_radius, int64_t n_poses, const float *poses, uint8_t *keep, int64_t *n_kept);
/* The same filter on the device (one thread per pose; at most 64 sampled ligand carbons): identical arguments and
 * result, for pose sets where the serial host loop would cost more than the scoring kernel itself. */
int mc_dock_filter_poses_gpu(mc_ctx *ctx, int64_t n_rec, const mc_float4 *rec_xyzq, const uint8_t *rec_is_carbon, int64_t n_lig,
                             const mc_float4 *lig_xyzq, const uint8_t *lig_is_carbon, const float lig_anchor[3],
                             float vdw_radius, int64_t n_poses, const float *poses, uint8_t *keep, int64_t *n_kept);
/* CUDA-event duration of the scan kernel of the last mc_dock_score call (profiling on). */
double mc_last_dock_kernel_ms(mc_ctx *ctx);

/* ---- domain decomposition: one handle per GPU / process ---------------------- */

/* 128-byte NCCL unique id, produced on rank 0 and distributed by the host (any transport). */
int mc_comm_unique_id(uint8_t id[128]);
/* Join the communicator; slabs along z.  Must precede mc_set_atoms; afterwards mc_set_atoms
 * takes the GLOBAL system on every rank and keeps the atoms this rank owns. */
int mc_comm_init(mc_ctx *ctx, const uint8_t id[128], int rank, int n_ranks);
/* The decomposition arithmetic on its own (host only, usable without a GPU): out = {ncx, ncy, ncz
 * global cell grid, kz0, kz1 owned z layers [kz0, kz1), ghost layer from prev, ghost layer from
 * next, next rank}.  r_list = max(rc_lj, rc_q) + skin. */
int mc_dd_plan(const float box_ext[3], float r_list, int rank, int n_ranks, int32_t out[8]);
/* How the per-step ghost refresh runs: *fused = 1 when the neighbours' position arrays are mapped into
 * this process (cudaIpc over NVLink) and the step kernels exchange ghosts themselves -- kick_drift
 * stores its boundary layers into the neighbours' ghost blocks and raises their flags, the pair
 * kernel of the boundary rows waits on them; 0 = ncclSend / ncclRecv between the two kernels (option
 * "halo_fused" = 0, or the mapping failed: `why` then holds the reason).  Valid after the first build. */
int mc_comm_halo_mode(mc_ctx *ctx, int *fused, char *why, int why_cap);
/* Rebuild schedule of a decomposed run.  Every rank must rebuild at the same step, so the decision cannot
 * be the local displacement flag: option "rebuild_every" = k > 0 fixes the interval; 0 (default) adapts
 * it at every build from the largest displacement any rank saw in the interval that just ended (the
 * number travels with the build's all-gathered layout table, so all ranks derive the same interval), aiming
 * at 75 % of skin/2.  *interval = steps between builds now in force, *last_disp_frac = that largest
 * displacement / (skin/2); mc_stats.n_list_violations counts intervals that overshot. */
int mc_comm_schedule(mc_ctx *ctx, int *interval, double *last_disp_frac);
/* Number of atoms this rank currently owns / holds as ghosts. */
int mc_comm_counts(mc_ctx *ctx, int64_t *n_owned, int64_t *n_ghost);
/* Gather global arrays (original ids, length n_global) -- valid on every rank. */
int mc_get_positions_global(mc_ctx *ctx, mc_float4 *out);
int mc_get_forces_global(mc_ctx *ctx, mc_float4 *out);

#ifdef __cplusplus
}
#endif
#endif /* MOLCHANICA_MD_H */
