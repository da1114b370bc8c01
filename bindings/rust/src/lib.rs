//! Raw FFI binding of `libmolchanica_md.so` (include/molchanica_md.h) for the reference's host side: what
//! `dynamics` / `src/md` would link in place of the PTX module of build.rs:10-16 (INTEGRATION.md says where each call
//! goes).  GENERATED from the header by tools/gen_rust_binding.py -- do not edit by hand; tests/test_abi.py keeps it
//! in step with the header.  There is no Rust toolchain in the build image: this file has not been compiled.
#![allow(non_camel_case_types, clippy::too_many_arguments)]
use std::os::raw::{c_char, c_int};

pub const MC_ABI_VERSION: c_int = 2;
pub const MC_OK: c_int = 0;
pub const MC_E_INVALID: c_int = -1;
pub const MC_E_CUDA: c_int = -2;
pub const MC_E_NODEVICE: c_int = -3;
pub const MC_E_CAPACITY: c_int = -4;
pub const MC_E_COMM: c_int = -5;
pub const MC_W_STALE_LIST: c_int = 1;
pub const MC_COULOMB_NONE: c_int = 0;
pub const MC_COULOMB_PLAIN: c_int = 1;
pub const MC_COULOMB_ERFC: c_int = 2;
pub const MC_THERMOSTAT_NONE: c_int = 0;
pub const MC_THERMOSTAT_LANGEVIN: c_int = 1;
pub const MC_THERMOSTAT_CSVR: c_int = 2;
pub const MC_FLAG_STATIC: u8 = 1;
pub const MC_BAROSTAT_NONE: c_int = 0;
pub const MC_BAROSTAT_BERENDSEN: c_int = 1;
pub const MC_BAROSTAT_CRESCALE: c_int = 2;
pub const MC_DOCK_MAX_FLEX: c_int = 12;

#[repr(C)]
pub struct McCtx {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct McFloat4 {
    pub x: f32,
    pub y: f32,
    pub z: f32,
    pub w: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct McEnergy {
    pub energy_potential: f64,
    pub energy_potential_nonbonded: f64,
    pub energy_potential_bonded: f64,
    pub energy_kinetic: f64,
    pub temperature: f64,
    pub energy_bond: f64,
    pub energy_angle: f64,
    pub energy_dihedral: f64,
    pub volume: f64,
    pub density: f64,
    pub energy_pme: f64,
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct McStats {
    pub n_atoms: i64,
    pub n_ghosts: i64,
    pub n_pairs_listed: i64,
    pub n_rebuilds: i64,
    pub n_steps: i64,
    pub n_kernel_launches: i64,
    pub n_cells: [i64; 3],
    pub pair_ms_sum: f64,
    pub pair_launches_timed: i64,
    pub build_ms_sum: f64,
    pub builds_timed: i64,
    pub integrate_ms_sum: f64,
    pub integrate_launches_timed: i64,
    pub halo_ms_sum: f64,
    pub halos_timed: i64,
    pub n_list_violations: i64,
    pub list_bytes: i64,
    pub ext_upload_bytes: i64,
}

#[link(name = "molchanica_md")]
extern "C" {
    pub fn mc_create(device: c_int, out: *mut *mut McCtx) -> c_int;
    pub fn mc_destroy(ctx: *mut McCtx) -> c_int;
    pub fn mc_last_error(ctx: *const McCtx) -> *const c_char;
    pub fn mc_abi_version() -> c_int;
    pub fn mc_struct_sizes(energy_bytes: *mut c_int, stats_bytes: *mut c_int) -> c_int;
    pub fn mc_set_box(ctx: *mut McCtx, lo: *const f32, hi: *const f32, periodic: c_int) -> c_int;
    pub fn mc_set_atoms(ctx: *mut McCtx, n: i64, xyzq: *const McFloat4, type_: *const u16, vel_invmass: *const McFloat4, flags: *const u8) -> c_int;
    pub fn mc_set_lj_table(ctx: *mut McCtx, n_types: c_int, sigma_eps: *const f32) -> c_int;
    pub fn mc_set_exclusions(ctx: *mut McCtx, start: *const i32, idx: *const i32) -> c_int;
    pub fn mc_set_pairs14(ctx: *mut McCtx, m: i64, pairs: *const i32, scale_lj: f32, scale_q: f32) -> c_int;
    pub fn mc_set_bonds(ctx: *mut McCtx, m: i64, pairs: *const i32, k_r0: *const f32) -> c_int;
    pub fn mc_set_angles(ctx: *mut McCtx, m: i64, triples: *const i32, k_theta0: *const f32) -> c_int;
    pub fn mc_set_dihedrals(ctx: *mut McCtx, m: i64, quads: *const i32, pk_n_phase: *const f32) -> c_int;
    pub fn mc_set_thermostat(ctx: *mut McCtx, kind: c_int, temperature_k: f32, gamma_per_ps: f32, seed: u64) -> c_int;
    pub fn mc_set_pme(ctx: *mut McCtx, k1: c_int, k2: c_int, k3: c_int) -> c_int;
    pub fn mc_pme_suggest(rc: f32, tol: f32, box_ext: *const f32, alpha: *mut f32, grid: *mut i32) -> c_int;
    pub fn mc_set_rigid_waters(ctx: *mut McCtx, m: i64, triples: *const i32, d_oh: f32, d_hh: f32, m_o: f32, m_h: f32) -> c_int;
    pub fn mc_set_hbond_constraints(ctx: *mut McCtx, m: i64, clusters: *const i32, lengths: *const f32) -> c_int;
    pub fn mc_set_virtual_sites(ctx: *mut McCtx, m: i64, quads: *const i32, a: f32, b: f32) -> c_int;
    pub fn mc_set_cutoffs(ctx: *mut McCtx, rc_lj: f32, rc_q: f32, skin: f32, coulomb_mode: c_int, alpha: f32) -> c_int;
    pub fn mc_set_overrides(ctx: *mut McCtx, lj_disabled: c_int, coulomb_disabled: c_int) -> c_int;
    pub fn mc_set_option(ctx: *mut McCtx, name: *const c_char, value: f64) -> c_int;
    pub fn mc_set_positions(ctx: *mut McCtx, xyzq: *const McFloat4) -> c_int;
    pub fn mc_set_velocities(ctx: *mut McCtx, vel_invmass: *const McFloat4) -> c_int;
    pub fn mc_build_neighbors(ctx: *mut McCtx) -> c_int;
    pub fn mc_compute_forces(ctx: *mut McCtx) -> c_int;
    pub fn mc_step(ctx: *mut McCtx, dt: f32, n_steps: c_int, ext_forces: *const f32) -> c_int;
    pub fn mc_last_step_ms(ctx: *mut McCtx) -> f64;
    pub fn mc_minimize_energy(ctx: *mut McCtx, max_iters: c_int, iters_accepted: *mut c_int, e_initial: *mut f64, e_final: *mut f64) -> c_int;
    pub fn mc_get_positions(ctx: *mut McCtx, out: *mut McFloat4) -> c_int;
    pub fn mc_get_velocities(ctx: *mut McCtx, out: *mut McFloat4) -> c_int;
    pub fn mc_get_forces(ctx: *mut McCtx, out: *mut McFloat4) -> c_int;
    pub fn mc_get_energy(ctx: *mut McCtx, out: *mut McEnergy) -> c_int;
    pub fn mc_get_stats(ctx: *mut McCtx, out: *mut McStats) -> c_int;
    pub fn mc_get_pressure(ctx: *mut McCtx, pressure_bar: *mut f64, virial: *mut f64) -> c_int;
    pub fn mc_set_barostat(ctx: *mut McCtx, kind: c_int, pressure_bar: f32, tau_ps: f32, compressibility_per_bar: f32, every_n_steps: c_int, seed: u64) -> c_int;
    pub fn mc_get_box(ctx: *mut McCtx, lo: *mut f32, hi: *mut f32) -> c_int;
    pub fn mc_set_molecule_ids(ctx: *mut McCtx, mol_id: *const u16) -> c_int;
    pub fn mc_get_energy_between_mols(ctx: *mut McCtx, out: *mut f64) -> c_int;
    pub fn mc_snapshot_begin(ctx: *mut McCtx, out_positions: *mut McFloat4, out_ids: *mut i32, n_out: *mut i64) -> c_int;
    pub fn mc_snapshot_begin_xyz(ctx: *mut McCtx, out_xyz: *mut f32, out_ids: *mut i32, n_out: *mut i64, layout_epoch: *mut i64) -> c_int;
    pub fn mc_snapshot_begin_pv(ctx: *mut McCtx, out_positions: *mut McFloat4, out_velocities: *mut McFloat4, out_ids: *mut i32, n_out: *mut i64) -> c_int;
    pub fn mc_snapshot_wait(ctx: *mut McCtx) -> c_int;
    pub fn mc_get_neighbors(ctx: *mut McCtx, start: *mut i64, idx: *mut i32, cap: i64, total: *mut i64) -> c_int;
    pub fn mc_reset_timers(ctx: *mut McCtx) -> c_int;
    pub fn mc_time_kernels(ctx: *mut McCtx, reps: c_int, flush_l2: c_int) -> c_int;
    pub fn mc_last_pair_kernel_ms(ctx: *mut McCtx) -> f64;
    pub fn mc_dock_score(ctx: *mut McCtx, n_rec: i64, rec_xyzq: *const McFloat4, rec_type: *const u16, rec_hydrophobic: *const u8, n_lig: i64, lig_xyzq: *const McFloat4, lig_type: *const u16, lig_hydrophobic: *const u8, lig_anchor: *const f32, n_rec_types: c_int, n_lig_types: c_int, ljtab: *const f32, n_poses: i64, poses: *const f32, out: *mut f32) -> c_int;
    pub fn mc_dock_score_flex(ctx: *mut McCtx, n_rec: i64, rec_xyzq: *const McFloat4, rec_type: *const u16, rec_hydrophobic: *const u8, n_lig: i64, lig_xyzq: *const McFloat4, lig_type: *const u16, lig_hydrophobic: *const u8, lig_anchor: *const f32, n_rec_types: c_int, n_lig_types: c_int, ljtab: *const f32, n_flex: c_int, flex_axis: *const i32, flex_mask: *const u8, n_poses: i64, poses: *const f32, out: *mut f32) -> c_int;
    pub fn mc_dock_make_poses(site_center: *const f64, site_radius: f64, num_posits: c_int, num_orientations: c_int, out_poses: *mut f32, cap: i64, n_out: *mut i64) -> c_int;
    pub fn mc_dock_orientation_count(num_orientations: c_int) -> c_int;
    pub fn mc_dock_make_poses_flex(site_center: *const f64, site_radius: f64, num_posits: c_int, num_orientations: c_int, n_flex_bonds: c_int, angles_per_bond: c_int, out_poses: *mut f32, cap: i64, n_out: *mut i64) -> c_int;
    pub fn mc_dock_flex_masks(n_lig: i64, n_bonds: i64, bonds: *const i32, n_flex_bonds: c_int, flex_bond_idx: *const i32, axis_out: *mut i32, mask_out: *mut u8) -> c_int;
    pub fn mc_dock_near_site(n_rec: i64, rec_xyzq: *const McFloat4, rec_hetero: *const u8, site_center: *const f64, site_radius: f64, out_idx: *mut i32, n_out: *mut i64) -> c_int;
    pub fn mc_dock_filter_poses(n_rec: i64, rec_xyzq: *const McFloat4, rec_is_carbon: *const u8, n_lig: i64, lig_xyzq: *const McFloat4, lig_is_carbon: *const u8, lig_anchor: *const f32, vdw_radius: f32, n_poses: i64, poses: *const f32, keep: *mut u8, n_kept: *mut i64) -> c_int;
    pub fn mc_dock_filter_poses_gpu(ctx: *mut McCtx, n_rec: i64, rec_xyzq: *const McFloat4, rec_is_carbon: *const u8, n_lig: i64, lig_xyzq: *const McFloat4, lig_is_carbon: *const u8, lig_anchor: *const f32, vdw_radius: f32, n_poses: i64, poses: *const f32, keep: *mut u8, n_kept: *mut i64) -> c_int;
    pub fn mc_last_dock_kernel_ms(ctx: *mut McCtx) -> f64;
    pub fn mc_comm_unique_id(id: *mut u8) -> c_int;
    pub fn mc_comm_init(ctx: *mut McCtx, id: *const u8, rank: c_int, n_ranks: c_int) -> c_int;
    pub fn mc_dd_plan(box_ext: *const f32, r_list: f32, rank: c_int, n_ranks: c_int, out: *mut i32) -> c_int;
    pub fn mc_comm_halo_mode(ctx: *mut McCtx, fused: *mut c_int, why: *mut c_char, why_cap: c_int) -> c_int;
    pub fn mc_comm_schedule(ctx: *mut McCtx, interval: *mut c_int, last_disp_frac: *mut f64) -> c_int;
    pub fn mc_comm_counts(ctx: *mut McCtx, n_owned: *mut i64, n_ghost: *mut i64) -> c_int;
    pub fn mc_get_positions_global(ctx: *mut McCtx, out: *mut McFloat4) -> c_int;
    pub fn mc_get_forces_global(ctx: *mut McCtx, out: *mut McFloat4) -> c_int;
}

/// Maps a status code to the error text `build_dynamics` propagates as `ParamError` (reference src/md/mod.rs:651).
pub fn check(ctx: *const McCtx, rc: c_int) -> Result<(), String> {
    if rc >= MC_OK {
        return Ok(()); // positive codes are warnings (MC_W_STALE_LIST): the call completed, mc_last_error has the text
    }
    let msg = unsafe { std::ffi::CStr::from_ptr(mc_last_error(ctx)) };
    Err(format!("molchanica_md error {}: {}", rc, msg.to_string_lossy()))
}
