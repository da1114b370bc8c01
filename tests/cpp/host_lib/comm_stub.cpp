// Test infrastructure: what the host build of the library (tests/cpp/host_lib/) links in place of comm.cu.  One process
// on one "device" has no peers: a handle is never decomposed, the collective entry points refuse.
#define MC_HOST_SHIM 1
#include "../../../molchanica_b200/csrc/engine.cuh"

int comm_set_atoms(mc_ctx *, int64_t, const mc_float4 *, const uint16_t *, const mc_float4 *, const uint8_t *) { return MC_E_COMM; }
int comm_rebuild(mc_ctx *) { return MC_E_COMM; }
int comm_halo_positions(mc_ctx *) { return MC_E_COMM; }
bool comm_peer_direct(const mc_ctx *) { return false; }
void comm_set_migrate(mc_ctx *, bool) {}
int comm_interval(const mc_ctx *) { return 1; }
void comm_step_descriptors(mc_ctx *, bool, HaloPush *, HaloSplit *) {}
int comm_agree_flag(mc_ctx *, bool *) { return MC_E_COMM; }
int comm_allreduce3(mc_ctx *, double[3]) { return MC_E_COMM; }
int comm_allreduce_f4(mc_ctx *, float4 *, int64_t) { return MC_E_COMM; }
void comm_destroy(mc_ctx *) {}

extern "C" {
int mc_comm_unique_id(uint8_t[128]) { return MC_E_COMM; }
int mc_comm_init(mc_ctx *, const uint8_t[128], int, int) { return MC_E_COMM; }
int mc_comm_counts(mc_ctx *, int64_t *, int64_t *) { return MC_E_COMM; }
int mc_get_positions_global(mc_ctx *c, mc_float4 *out) { return mc_get_positions(c, out); }
int mc_get_forces_global(mc_ctx *c, mc_float4 *out) { return mc_get_forces(c, out); }
int mc_dd_plan(const float[3], float, int, int, int32_t[8]) { return MC_E_COMM; }
int mc_comm_schedule(mc_ctx *, int *, double *) { return MC_E_COMM; }
int mc_comm_halo_mode(mc_ctx *, int *, char *, int) { return MC_E_COMM; }
}
