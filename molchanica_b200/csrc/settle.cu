// settle.cu -- rigid three-site water on the device: one thread per molecule applies the
// analytic SETTLE of settle_terms.h to the positions kick_drift has just advanced and corrects the velocities
// by (x_constrained - x_unconstrained) / dt, the leap-frog form of the constraint (the merged kicks of the
// step loop make the velocity a half-step one).  The old positions are recovered as x' - v dt, and everything
// is done relative to the oxygen's old position, so no extra position array is kept and fp32 cancellation
// stays at the 1e-6 A level; each atom keeps its own periodic image (molecules may straddle the box edge).
// HBM-bound and tiny: 3 x (16 + 16) B read and written per molecule.
// After the closing half kick of every step the velocity stage of RATTLE (rattle_terms.h) removes the velocity components
// along the constrained bonds, so that kinetic energy / temperature / pressure read between steps are the constrained
// system's.  Confirmed on hardware at the end of round 1 (positions) / in round 2 (velocity stage).
#include "settle.cuh"
#include "settle_terms.h"
#include "vsite_terms.h"
#include "shake_terms.h"
#include "rattle_terms.h"

namespace {

// Adds a thread's share of the constraint virial to *virial (one fp64 atomic per warp).  The constraint force that
// moved atom i by delta_i within this step is 2 m_i delta_i / dt^2 (half kick + drift form); its virial is taken with the OLD
// positions relative to the molecule's first atom (the constraint forces of a molecule sum to zero), in kcal/mol.
__device__ __forceinline__ void add_constraint_virial(float wc, double *__restrict__ virial) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) wc += __shfl_xor_sync(MC_FULL_MASK, wc, d);
    if ((threadIdx.x & 31) == 0 && wc != 0.f) atomicAdd(virial, (double)wc);
}

__device__ __forceinline__ float settle_one(int w, const int4 *__restrict__ waters, const int *__restrict__ slot_of_orig,
                                            float4 *__restrict__ xyzq, float4 *__restrict__ vel, const SettleParams &sp,
                                            const NbParams &p, float dt) {
    const int4 ids = waters[w];
    const int so = slot_of_orig[ids.x], s1 = slot_of_orig[ids.y], s2 = slot_of_orig[ids.z];
    float4 xo = xyzq[so], x1 = xyzq[s1], x2 = xyzq[s2];
    float4 vo = vel[so], v1 = vel[s1], v2 = vel[s2];
    const float xo_[3] = {xo.x, xo.y, xo.z}, x1_[3] = {x1.x, x1.y, x1.z}, x2_[3] = {x2.x, x2.y, x2.z};
    const float vo_[3] = {vo.x, vo.y, vo.z}, v1_[3] = {v1.x, v1.y, v1.z}, v2_[3] = {v2.x, v2.y, v2.z};
    float b0[3], c0[3], a1[3], b1[3], c1[3], a3[3], b3[3], c3[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        // new relative vectors (minimum image), then back one drift: old H - old O
        float db = x1_[a] - xo_[a], dc = x2_[a] - xo_[a];
        if (p.periodic) {
            db -= rintf(db * p.inv_ext[a]) * p.ext[a];
            dc -= rintf(dc * p.inv_ext[a]) * p.ext[a];
        }
        a1[a] = vo_[a] * dt;
        b0[a] = db - (v1_[a] - vo_[a]) * dt;
        c0[a] = dc - (v2_[a] - vo_[a]) * dt;
        b1[a] = b0[a] + v1_[a] * dt;
        c1[a] = c0[a] + v2_[a] * dt;
    }
    mc_settle(sp, b0, c0, a1, b1, c1, a3, b3, c3);
    const float inv_dt = 1.f / dt;
    float da[3], db_[3], dc_[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) { da[a] = a3[a] - a1[a]; db_[a] = b3[a] - b1[a]; dc_[a] = c3[a] - c1[a]; }
    xo.x += da[0]; xo.y += da[1]; xo.z += da[2];
    x1.x += db_[0]; x1.y += db_[1]; x1.z += db_[2];
    x2.x += dc_[0]; x2.y += dc_[1]; x2.z += dc_[2];
    vo.x += da[0] * inv_dt; vo.y += da[1] * inv_dt; vo.z += da[2] * inv_dt;
    v1.x += db_[0] * inv_dt; v1.y += db_[1] * inv_dt; v1.z += db_[2] * inv_dt;
    v2.x += dc_[0] * inv_dt; v2.y += dc_[1] * inv_dt; v2.z += dc_[2] * inv_dt;
    xyzq[so] = xo; xyzq[s1] = x1; xyzq[s2] = x2;
    vel[so] = vo; vel[s1] = v1; vel[s2] = v2;
    // oxygen: old relative position 0, no contribution
    // The step is half kick + drift (the closing half kick and RATTLE's velocity stage follow the force evaluation): the
    // position stage's displacement is that of a constraint force acting through ONE half kick, delta = f_c dt^2 / (2 m).
    return 2.f * sp.m_h * (db_[0] * b0[0] + db_[1] * b0[1] + db_[2] * b0[2] + dc_[0] * c0[0] + dc_[1] * c0[1] + dc_[2] * c0[2]) * inv_dt * inv_dt *
           (1.f / (float)MC_ACCEL_CONV);
}

__global__ void __launch_bounds__(128) settle_kernel(int n_w, const int4 *__restrict__ waters, const int *__restrict__ slot_of_orig,
                                                      float4 *__restrict__ xyzq, float4 *__restrict__ vel, const SettleParams sp,
                                                      const NbParams p, float dt, double *__restrict__ virial) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const float wc = w < n_w ? settle_one(w, waters, slot_of_orig, xyzq, vel, sp, p, dt) : 0.f;
    if (virial) add_constraint_virial(wc, virial);
}

// Bonds to hydrogen (shake_terms.h): one thread per heavy atom with its <= 3 hydrogens, same bookkeeping as settle_kernel
// (old positions = x' - v dt, everything relative to the heavy atom's old position, velocities corrected by the
// position change / dt).  clusters: (heavy, h1, h2, h3) original ids, -1 = no such hydrogen; dist: 3 lengths per cluster.
__device__ __forceinline__ float shake_h_one(int c, const int4 *__restrict__ clusters, const float *__restrict__ dist,
                                             const int *__restrict__ slot_of_orig, float4 *__restrict__ xyzq,
                                             float4 *__restrict__ vel, const NbParams &p, float dt, float tol,
                                             int *__restrict__ not_converged) {
    const int4 ids = clusters[c];
    const int hid[3] = {ids.y, ids.z, ids.w};
    const int s0 = slot_of_orig[ids.x];
    float4 x0 = xyzq[s0], v0 = vel[s0];
    int sh[MC_SHAKE_MAX_H], nh = 0;
    float4 xh[MC_SHAKE_MAX_H], vh[MC_SHAKE_MAX_H];
    float r0[MC_SHAKE_MAX_H][3], pp[MC_SHAKE_MAX_H][3], q1[MC_SHAKE_MAX_H][3], inv_m[MC_SHAKE_MAX_H], d[MC_SHAKE_MAX_H];
#pragma unroll
    for (int k = 0; k < MC_SHAKE_MAX_H; ++k) {
        if (hid[k] < 0) continue;
        sh[nh] = slot_of_orig[hid[k]];
        xh[nh] = xyzq[sh[nh]];
        vh[nh] = vel[sh[nh]];
        inv_m[nh] = vh[nh].w;
        d[nh] = dist[3 * c + k];
        ++nh;
    }
    const float x0_[3] = {x0.x, x0.y, x0.z}, v0_[3] = {v0.x, v0.y, v0.z};
    float p0[3], a1[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) { a1[a] = v0_[a] * dt; p0[a] = a1[a]; }
    for (int k = 0; k < nh; ++k) {
        const float xk[3] = {xh[k].x, xh[k].y, xh[k].z}, vk[3] = {vh[k].x, vh[k].y, vh[k].z};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float db = xk[a] - x0_[a];
            if (p.periodic) db -= rintf(db * p.inv_ext[a]) * p.ext[a];
            r0[k][a] = db - (vk[a] - v0_[a]) * dt;   // old heavy -> hydrogen vector
            q1[k][a] = r0[k][a] + vk[a] * dt;        // unconstrained new hydrogen, relative to the old heavy atom
            pp[k][a] = q1[k][a];
        }
    }
    const int sweeps = mc_shake_cluster(nh, r0, p0, pp, v0.w, inv_m, d, tol, 64);
    if (sweeps > 64) atomicAdd(not_converged, 1);
    const float inv_dt = 1.f / dt;
    {
        const float da[3] = {p0[0] - a1[0], p0[1] - a1[1], p0[2] - a1[2]};
        x0.x += da[0]; x0.y += da[1]; x0.z += da[2];
        v0.x += da[0] * inv_dt; v0.y += da[1] * inv_dt; v0.z += da[2] * inv_dt;
        xyzq[s0] = x0; vel[s0] = v0;
    }
    float wc = 0.f;  // heavy atom: old relative position 0
    for (int k = 0; k < nh; ++k) {
        const float dk[3] = {pp[k][0] - q1[k][0], pp[k][1] - q1[k][1], pp[k][2] - q1[k][2]};
        xh[k].x += dk[0]; xh[k].y += dk[1]; xh[k].z += dk[2];
        vh[k].x += dk[0] * inv_dt; vh[k].y += dk[1] * inv_dt; vh[k].z += dk[2] * inv_dt;
        xyzq[sh[k]] = xh[k]; vel[sh[k]] = vh[k];
        wc += (dk[0] * r0[k][0] + dk[1] * r0[k][1] + dk[2] * r0[k][2]) / inv_m[k];
    }
    return 2.f * wc * inv_dt * inv_dt * (1.f / (float)MC_ACCEL_CONV);  // delta = f_c dt^2 / (2 m), see settle_one
}

__global__ void __launch_bounds__(128) shake_h_kernel(int n_c, const int4 *__restrict__ clusters, const float *__restrict__ dist,
                                                       const int *__restrict__ slot_of_orig, float4 *__restrict__ xyzq,
                                                       float4 *__restrict__ vel, const NbParams p, float dt, float tol,
                                                       int *__restrict__ not_converged, double *__restrict__ virial) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const float wc = c < n_c ? shake_h_one(c, clusters, dist, slot_of_orig, xyzq, vel, p, dt, tol, not_converged) : 0.f;
    if (virial) add_constraint_virial(wc, virial);
}

// Virtual sites (vsite_terms.h): one thread per site.  construct: after the parents have their final positions of
// the step; spread: after every force of the evaluation has been accumulated, before the next kick.  The site is a
// static atom to the integrator (inverse mass 0), carries charge / LJ like any atom in the pair kernel.
__global__ void __launch_bounds__(128) vsite_construct_kernel(int n_v, const int4 *__restrict__ sites, const int *__restrict__ slot_of_orig,
                                                               float4 *__restrict__ xyzq, float a, float b, const NbParams p) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_v) return;
    const int4 ids = sites[v];
    const int sm = slot_of_orig[ids.x], so = slot_of_orig[ids.y], s1 = slot_of_orig[ids.z], s2 = slot_of_orig[ids.w];
    const float4 xo = xyzq[so], x1 = xyzq[s1], x2 = xyzq[s2];
    const float o[3] = {xo.x, xo.y, xo.z};
    float d1[3] = {x1.x - xo.x, x1.y - xo.y, x1.z - xo.z}, d2[3] = {x2.x - xo.x, x2.y - xo.y, x2.z - xo.z}, m[3];
    if (p.periodic) {
#pragma unroll
        for (int x = 0; x < 3; ++x) {
            d1[x] -= rintf(d1[x] * p.inv_ext[x]) * p.ext[x];
            d2[x] -= rintf(d2[x] * p.inv_ext[x]) * p.ext[x];
        }
    }
    mc_vsite_position(o, d1, d2, a, b, m);
    float4 xm = xyzq[sm];
    xm.x = m[0]; xm.y = m[1]; xm.z = m[2];
    xyzq[sm] = xm;
}

__global__ void __launch_bounds__(128) vsite_spread_kernel(int n_v, const int4 *__restrict__ sites, const int *__restrict__ slot_of_orig,
                                                            float4 *__restrict__ force, float a, float b) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_v) return;
    const int4 ids = sites[v];
    const int sm = slot_of_orig[ids.x], so = slot_of_orig[ids.y], s1 = slot_of_orig[ids.z], s2 = slot_of_orig[ids.w];
    float4 fm4 = force[sm];
    const float fm[3] = {fm4.x, fm4.y, fm4.z};
    float fo[3], f1[3], f2[3];
    mc_vsite_spread(fm, a, b, fo, f1, f2);
    // every parent belongs to exactly one site: plain read-modify-write
    float4 t = force[so]; t.x += fo[0]; t.y += fo[1]; t.z += fo[2]; force[so] = t;
    t = force[s1]; t.x += f1[0]; t.y += f1[1]; t.z += f1[2]; force[s1] = t;
    t = force[s2]; t.x += f2[0]; t.y += f2[1]; t.z += f2[2]; force[s2] = t;
    fm4.x = fm4.y = fm4.z = 0.f;   // the energy row sum in .w stays
    force[sm] = fm4;
}

// Velocity stage of RATTLE: one thread per rigid water / per hydrogen cluster (rattle_terms.h).
__global__ void __launch_bounds__(128) rattle_waters_kernel(int n_w, const int4 *__restrict__ waters, const int *__restrict__ slot_of_orig,
                                                             const float4 *__restrict__ xyzq, float4 *__restrict__ vel, const NbParams p) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_w) return;
    const int4 ids = waters[w];
    const int sl[3] = {slot_of_orig[ids.x], slot_of_orig[ids.y], slot_of_orig[ids.z]};
    float r[4][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}}, v[4][3], inv_m[4] = {0.f, 0.f, 0.f, 0.f};
    float4 vv[3];
    const float4 x0 = xyzq[sl[0]];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float4 x = xyzq[sl[k]];
        vv[k] = vel[sl[k]];
        float d[3] = {x.x - x0.x, x.y - x0.y, x.z - x0.z};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (p.periodic) d[a] -= rintf(d[a] * p.inv_ext[a]) * p.ext[a];
            r[k][a] = d[a];
        }
        v[k][0] = vv[k].x; v[k][1] = vv[k].y; v[k][2] = vv[k].z;
        inv_m[k] = vv[k].w;
    }
    v[3][0] = v[3][1] = v[3][2] = 0.f;
    const int ci[3] = {0, 0, 1}, cj[3] = {1, 2, 2};
    mc_rattle_velocity<float>(3, ci, cj, r, inv_m, v);
#pragma unroll
    for (int k = 0; k < 3; ++k) vel[sl[k]] = make_float4(v[k][0], v[k][1], v[k][2], vv[k].w);
}

__global__ void __launch_bounds__(128) rattle_h_kernel(int n_c, const int4 *__restrict__ clusters, const int *__restrict__ slot_of_orig,
                                                        const float4 *__restrict__ xyzq, float4 *__restrict__ vel, const NbParams p) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_c) return;
    const int4 ids = clusters[c];
    const int id[4] = {ids.x, ids.y, ids.z, ids.w};
    int sl[4], nat = 0, ci[3] = {0, 0, 0}, cj[3] = {1, 2, 3};
    float r[4][3], v[4][3], inv_m[4], w4[4];
    float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < 4; ++k) {
        r[k][0] = r[k][1] = r[k][2] = v[k][0] = v[k][1] = v[k][2] = 0.f;
        inv_m[k] = 0.f;
        if (id[k] < 0) continue;
        sl[nat] = slot_of_orig[id[k]];
        const float4 x = xyzq[sl[nat]], vv = vel[sl[nat]];
        if (nat == 0) x0 = x;
        float d[3] = {x.x - x0.x, x.y - x0.y, x.z - x0.z};
        for (int a = 0; a < 3; ++a) {
            if (p.periodic) d[a] -= rintf(d[a] * p.inv_ext[a]) * p.ext[a];
            r[nat][a] = d[a];
        }
        v[nat][0] = vv.x; v[nat][1] = vv.y; v[nat][2] = vv.z;
        inv_m[nat] = vv.w; w4[nat] = vv.w;
        ++nat;
    }
    if (nat < 2) return;
    mc_rattle_velocity<float>(nat - 1, ci, cj, r, inv_m, v);
    for (int k = 0; k < nat; ++k) vel[sl[k]] = make_float4(v[k][0], v[k][1], v[k][2], w4[k]);
}

}  // namespace

#ifdef MC_HAVE_LAUNCH  // the serial stand-in of tests/cpp/shim/ has no launcher
void launch_rattle_velocities(int n_w, const int4 *waters, int n_c, const int4 *clusters, const int *slot_of_orig, const float4 *xyzq,
                              float4 *vel, const NbParams &p, cudaStream_t st, int64_t *launches) {
    if (n_w > 0) {
        MC_LAUNCH(rattle_waters_kernel, div_up((size_t)n_w, 128), 128, 0, st, n_w, waters, slot_of_orig, xyzq, vel, p);
        *launches += 1;
    }
    if (n_c > 0) {
        MC_LAUNCH(rattle_h_kernel, div_up((size_t)n_c, 128), 128, 0, st, n_c, clusters, slot_of_orig, xyzq, vel, p);
        *launches += 1;
    }
}

void launch_vsite_construct(int n_v, const int4 *sites, const int *slot_of_orig, float4 *xyzq, float a, float b, const NbParams &p,
                            cudaStream_t st, int64_t *launches) {
    if (n_v <= 0) return;
    MC_LAUNCH(vsite_construct_kernel, div_up((size_t)n_v, 128), 128, 0, st, n_v, sites, slot_of_orig, xyzq, a, b, p);
    *launches += 1;
}

void launch_vsite_spread(int n_v, const int4 *sites, const int *slot_of_orig, float4 *force, float a, float b, cudaStream_t st,
                         int64_t *launches) {
    if (n_v <= 0) return;
    MC_LAUNCH(vsite_spread_kernel, div_up((size_t)n_v, 128), 128, 0, st, n_v, sites, slot_of_orig, force, a, b);
    *launches += 1;
}

void launch_shake_h(int n_c, const int4 *clusters, const float *dist, const int *slot_of_orig, float4 *xyzq, float4 *vel,
                    const NbParams &p, float dt, float tol, int *not_converged, double *virial, cudaStream_t st, int64_t *launches) {
    if (n_c <= 0) return;
    MC_LAUNCH(shake_h_kernel, div_up((size_t)n_c, 128), 128, 0, st, n_c, clusters, dist, slot_of_orig, xyzq, vel, p, dt, tol, not_converged,
              virial);
    *launches += 1;
}

void launch_settle(int n_w, const int4 *waters, const int *slot_of_orig, float4 *xyzq, float4 *vel, float m_o, float m_h,
                   float d_oh, float d_hh, const NbParams &p, float dt, double *virial, cudaStream_t st, int64_t *launches) {
    if (n_w <= 0) return;
    const SettleParams sp = mc_settle_params(m_o, m_h, d_oh, d_hh);
    MC_LAUNCH(settle_kernel, div_up((size_t)n_w, 128), 128, 0, st, n_w, waters, slot_of_orig, xyzq, vel, sp, p, dt, virial);
    *launches += 1;
}
#endif
