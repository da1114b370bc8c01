// bonded_device.cuh -- one bonded term per call (device side), shared by bonded.cu (one thread per term, one launch per
// evaluation) and md_fused.cu (the persistent multi-step kernel of small systems).  Arithmetic: bonded_terms.h.
#pragma once
#include "bonded.cuh"
#include "bonded_terms.h"

namespace {

__device__ __forceinline__ void min_image3(float d[3], const NbParams &p) {
    if (p.periodic) {
#pragma unroll
        for (int a = 0; a < 3; ++a) d[a] -= rintf(d[a] * p.inv_ext[a]) * p.ext[a];
    }
}

__device__ __forceinline__ void add_force(float4 *force, int slot, const float f[3]) {
    atomicAdd(&force[slot].x, f[0]);
    atomicAdd(&force[slot].y, f[1]);
    atomicAdd(&force[slot].z, f[2]);
}

// Ownership of a term's atoms on this rank (BondedTerms::own0 / own1): returns the number of owned atoms, -1 when an owned
// atom has a partner that is not held here (flagged), 0 when the term is somebody else's.
template <int N>
__device__ __forceinline__ int term_owned(const BondedTerms &t, const int (&s)[N], bool (&own)[N]) {
    int n_own = 0;
    bool held = true;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        own[k] = s[k] >= t.own0 && s[k] < t.own1;
        n_own += own[k] ? 1 : 0;
        held = held && s[k] >= 0;
    }
    if (n_own > 0 && !held) {
        if (t.missing) *t.missing = 1;
        return -1;
    }
    return n_own;
}

// Term `tid` of the concatenated list bonds | angles | dihedrals: adds its forces to `force` (fp32 atomics), returns its
// energy in e, its share of the virial in w and its kind (0 / 1 / 2, -1 = no such term).
__device__ __forceinline__ void bonded_term_apply(int tid, const BondedTerms &t, const int *__restrict__ slot_of_orig,
                                                  const float4 *xyzq, const NbParams &p, float4 *force, float &e, float &w, int &kind) {
    e = 0.f; w = 0.f; kind = -1;
    if (tid < t.n_bonds) {
        kind = 0;
        const int2 ij = t.bonds[tid];
        const float2 kr = t.bond_kr0[tid];
        const int si = slot_of_orig[ij.x], sj = slot_of_orig[ij.y];
        const int sl_[2] = {si, sj};
        bool own[2];
        const int n_own = term_owned(t, sl_, own);
        if (n_own <= 0) return;
        const float4 xi = xyzq[si], xj = xyzq[sj];
        float d[3] = {xi.x - xj.x, xi.y - xj.y, xi.z - xj.z}, fi[3];
        min_image3(d, p);
        const float share = (float)n_own * 0.5f;
        e = mc_bond_term(d, kr.x, kr.y, fi) * share;
        w = (d[0] * fi[0] + d[1] * fi[1] + d[2] * fi[2]) * share;
        const float fj[3] = {-fi[0], -fi[1], -fi[2]};
        if (own[0]) add_force(force, si, fi);
        if (own[1]) add_force(force, sj, fj);
    } else if (tid < t.n_bonds + t.n_angles) {
        kind = 1;
        const int a_ = tid - t.n_bonds;
        const int4 ijk = t.angles[a_];
        const float2 kt = t.angle_kt0[a_];
        const int si = slot_of_orig[ijk.x], sj = slot_of_orig[ijk.y], sk = slot_of_orig[ijk.z];
        const int sl_[3] = {si, sj, sk};
        bool own[3];
        const int n_own = term_owned(t, sl_, own);
        if (n_own <= 0) return;
        const float share = (float)n_own * (1.f / 3.f);
        const float4 xi = xyzq[si], xj = xyzq[sj], xk = xyzq[sk];
        float a[3] = {xi.x - xj.x, xi.y - xj.y, xi.z - xj.z}, b[3] = {xk.x - xj.x, xk.y - xj.y, xk.z - xj.z}, fi[3], fk[3];
        min_image3(a, p);
        min_image3(b, p);
        e = mc_angle_term(a, b, kt.x, kt.y, fi, fk) * share;
        w = (a[0] * fi[0] + a[1] * fi[1] + a[2] * fi[2] + b[0] * fk[0] + b[1] * fk[1] + b[2] * fk[2]) * share;  // zero up to rounding
        const float fj[3] = {-(fi[0] + fk[0]), -(fi[1] + fk[1]), -(fi[2] + fk[2])};
        if (own[0]) add_force(force, si, fi);
        if (own[1]) add_force(force, sj, fj);
        if (own[2]) add_force(force, sk, fk);
    } else if (tid < t.n_bonds + t.n_angles + t.n_dihedrals) {
        kind = 2;
        const int d_ = tid - t.n_bonds - t.n_angles;
        const int4 q = t.dihedrals[d_];
        const float4 prm = t.dihedral_prm[d_];  // pk, periodicity, phase
        const int si = slot_of_orig[q.x], sj = slot_of_orig[q.y], sk = slot_of_orig[q.z], sl = slot_of_orig[q.w];
        const int sl_[4] = {si, sj, sk, sl};
        bool own[4];
        const int n_own = term_owned(t, sl_, own);
        if (n_own <= 0) return;
        const float share = (float)n_own * 0.25f;
        const float4 xi = xyzq[si], xj = xyzq[sj], xk = xyzq[sk], xl = xyzq[sl];
        float rij[3] = {xi.x - xj.x, xi.y - xj.y, xi.z - xj.z}, rkj[3] = {xk.x - xj.x, xk.y - xj.y, xk.z - xj.z},
              rkl[3] = {xk.x - xl.x, xk.y - xl.y, xk.z - xl.z}, fi[3], fj[3], fk[3], fl[3];
        min_image3(rij, p);
        min_image3(rkj, p);
        min_image3(rkl, p);
        e = mc_dihedral_term(rij, rkj, rkl, prm.x, prm.y, prm.z, fi, fj, fk, fl) * share;
        // relative to atom j: r_i - r_j = rij, r_k - r_j = rkj, r_l - r_j = rkj - rkl (zero up to rounding as well)
        w = (rij[0] * fi[0] + rij[1] * fi[1] + rij[2] * fi[2] + rkj[0] * fk[0] + rkj[1] * fk[1] + rkj[2] * fk[2] +
             (rkj[0] - rkl[0]) * fl[0] + (rkj[1] - rkl[1]) * fl[1] + (rkj[2] - rkl[2]) * fl[2]) * share;
        if (own[0]) add_force(force, si, fi);
        if (own[1]) add_force(force, sj, fj);
        if (own[2]) add_force(force, sk, fk);
        if (own[3]) add_force(force, sl, fl);
    }
}

}  // namespace
