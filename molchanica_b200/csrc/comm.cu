// comm.cu -- slab domain decomposition + ghost-atom halo exchange.  Placeholder
// until the single-GPU path is parity-green: every entry point reports that the handle is not
// decomposed.
#include "engine.cuh"

struct CommState {};

static int not_yet(mc_ctx *c) {
    c->err = "domain decomposition is not built into this library yet";
    return MC_E_COMM;
}

int comm_set_atoms(mc_ctx *c, int64_t, const mc_float4 *, const uint16_t *, const mc_float4 *, const uint8_t *) { return not_yet(c); }
int comm_rebuild(mc_ctx *c) { return not_yet(c); }
int comm_halo_positions(mc_ctx *c) { return not_yet(c); }
int comm_agree_flag(mc_ctx *c, bool *) { return not_yet(c); }
int comm_allreduce3(mc_ctx *c, double *) { return not_yet(c); }
void comm_destroy(mc_ctx *) {}

extern "C" int mc_comm_unique_id(uint8_t *) { return MC_E_COMM; }
extern "C" int mc_comm_init(mc_ctx *c, const uint8_t *, int, int) { return c ? not_yet(c) : MC_E_INVALID; }
extern "C" int mc_comm_counts(mc_ctx *c, int64_t *n_owned, int64_t *n_ghost) {
    if (!c) return MC_E_INVALID;
    if (n_owned) *n_owned = c->n_rows;
    if (n_ghost) *n_ghost = c->n - c->n_rows;
    return MC_OK;
}
extern "C" int mc_get_positions_global(mc_ctx *c, mc_float4 *out) { return mc_get_positions(c, out); }
extern "C" int mc_get_forces_global(mc_ctx *c, mc_float4 *out) { return mc_get_forces(c, out); }
