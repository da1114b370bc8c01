/*
 * md_oracle.c -- CPU restatement of the Molchanica MD / docking hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library, and only as the checker / CPU baseline.  The product path (molchanica_b200/csrc)
 * never links or calls it.
 *
 * PARITY STATUS.  The engine arithmetic of the reference lives in the un-vendored crate
 * `dynamics` 0.2.2 (reference Cargo.toml:25,79) whose source is absent, and the reference holds
 * no MD test vectors (src/tests.rs:3-4 is empty).  What IS in-tree is the pair arithmetic of
 * src/cuda/util.cu; that file is host-compiled unmodified into oracle/_ref/libref_cuda.so
 * (see oracle/Makefile, oracle/ref_shim/) and this restatement is pinned against it in
 * tests/test_oracle_vs_ref.py and against tests/golden/ vectors generated from it.
 * Everything beyond those helper functions (neighbour list, exclusions, 1-4 scaling,
 * velocity Verlet, docking score) is "parity unpinned": a restatement of published
 * conventions, each citing the reference call site it serves.
 *
 * Conventions (each one mirrored bit-for-bit by the CUDA path where integers/indices are
 * concerned):
 *   - min image        d -= rintf(d / ext) * ext            (util.cu:65-71; host twin md/mod.rs:278-296)
 *   - squared distance ((dx*dx) + (dy*dy)) + (dz*dz), fp32, NO fma contraction (the reference
 *                      CPU path is Rust, which never contracts); compile with -ffp-contract=off
 *   - list membership  r2 < (r_cut + skin)^2 strictly, j != i, j not excluded; rows ascending
 *   - cutoff mask      r2 < rc^2 strictly, decided in fp32 with the expression above, also
 *                      when the pair arithmetic itself is evaluated in fp64 (truth mode)
 *   - LJ 12-6          sr = sigma/r; F = dir * 24 eps (2 sr^12 - sr^6)/r; E = 4 eps (sr^12 - sr^6);
 *                      dir = (r_tgt - r_src)/r  (util.cu:93-139)
 *   - Coulomb (plain)  F = dir * q_s q_t / (r^2 + 1e-6)  (util.cu:54-63); E = q_s q_t / r;
 *                      charges arrive pre-scaled by sqrt(332.0522)
 *   - Coulomb (erfc)   Ewald real-space: E = qq erfc(a r)/r,
 *                      F = dir * qq (erfc(a r)/r^2 + 2a/sqrt(pi) exp(-a^2 r^2)/r)  (INV_SQRT_PI, util.cu:15-18)
 *   - velocity Verlet  v += F/m * dt/2 * 418.4; x += v dt; F(x); v += F/m * dt/2 * 418.4
 *                      (units A, ps, amu, kcal/mol)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_SOFTENING_SQ 0.000001f          /* util.cu:9-10 */
#define ORC_INV_SQRT_PI 0.5641895835477563  /* util.cu:15-18 */
#define ORC_ACCEL_CONV 418.4f               /* kcal/mol/A/amu -> A/ps^2 */

#define ORC_COUL_NONE 0
#define ORC_COUL_PLAIN 1
#define ORC_COUL_ERFC 2

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* util.cu:65-71 */
static inline float min_image_f32(float d, float ext) {
    return d - rintf(d / ext) * ext;
}

float orc_min_image(float d, float ext) { return min_image_f32(d, ext); }

/* The one fp32 distance expression shared by list build and cutoff mask. */
static inline float dist2_f32(const float *a, const float *b, const float *ext, int periodic,
                              float *d /* out: a - b, min-imaged */) {
    float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    if (periodic) {
        dx = min_image_f32(dx, ext[0]);
        dy = min_image_f32(dy, ext[1]);
        dz = min_image_f32(dz, ext[2]);
    }
    d[0] = dx; d[1] = dy; d[2] = dz;
    return ((dx * dx) + (dy * dy)) + (dz * dz);
}

float orc_dist2(const float *a, const float *b, const float *ext, int periodic) {
    float d[3];
    return dist2_f32(a, b, ext, periodic, d);
}

/* ---- single-pair helpers, exported so tests can pin them against oracle/_ref ---------- */

/* util.cu:119-139 (lj_force) + :93-115 (lj_force_v2). out = {fx,fy,fz,energy}, force on tgt. */
void orc_pair_lj(const float *tgt, const float *src, float sigma, float eps, float *out) {
    float dx = tgt[0] - src[0], dy = tgt[1] - src[1], dz = tgt[2] - src[2];
    float r_sq = dx * dx + dy * dy + dz * dz;
    float r = sqrtf(r_sq);
    float inv_r = 1.0f / r;
    float sr = sigma * inv_r;
    float sr2 = sr * sr, sr4 = sr2 * sr2, sr6 = sr4 * sr2, sr12 = sr6 * sr6;
    float mag = 24.0f * eps * fmaf(2.f, sr12, -sr6) * inv_r;
    out[0] = dx * inv_r * mag;
    out[1] = dy * inv_r * mag;
    out[2] = dz * inv_r * mag;
    out[3] = 4.f * eps * (sr12 - sr6);
}

/* util.cu:54-63. out = {fx,fy,fz}, force on tgt. */
void orc_pair_coulomb(const float *tgt, const float *src, float q_src, float q_tgt, float *out) {
    float dx = tgt[0] - src[0], dy = tgt[1] - src[1], dz = tgt[2] - src[2];
    float dist = sqrtf(dx * dx + dy * dy + dz * dz);
    float mag = q_src * q_tgt / (dist * dist + ORC_SOFTENING_SQ);
    out[0] = dx / dist * mag;
    out[1] = dy / dist * mag;
    out[2] = dz / dist * mag;
}

/* ---- neighbour list --------------------------------------------------------------------- */

static int is_excluded(const int32_t *excl_start, const int32_t *excl_idx, int i, int j) {
    if (!excl_start) return 0;
    for (int32_t k = excl_start[i]; k < excl_start[i + 1]; ++k)
        if (excl_idx[k] == j) return 1;
    return 0;
}

/*
 * Brute-force O(N^2) Verlet list (the definition).  xyzq: n*4 floats.  Rows ascending.
 * Two-call protocol: out_idx == NULL -> only out_start (n+1 prefix) is filled.
 * Returns total entries, or -1 if cap is too small.
 */
int64_t orc_neighbors_brute(int n, const float *xyzq, const float *ext, int periodic, float r_list,
                            const int32_t *excl_start, const int32_t *excl_idx,
                            int64_t *out_start, int32_t *out_idx, int64_t cap) {
    const float rl2 = r_list * r_list;
    int64_t *cnt = (int64_t *)calloc((size_t)n + 1, sizeof(int64_t));
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < n; ++i) {
        int64_t c = 0;
        float d[3];
        for (int j = 0; j < n; ++j) {
            if (j == i) continue;
            if (dist2_f32(xyzq + 4 * i, xyzq + 4 * j, ext, periodic, d) < rl2 &&
                !is_excluded(excl_start, excl_idx, i, j))
                ++c;
        }
        cnt[i] = c;
    }
    int64_t tot = 0;
    for (int i = 0; i < n; ++i) { out_start[i] = tot; tot += cnt[i]; }
    out_start[n] = tot;
    free(cnt);
    if (!out_idx) return tot;
    if (tot > cap) return -1;
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < n; ++i) {
        int64_t p = out_start[i];
        float d[3];
        for (int j = 0; j < n; ++j) {
            if (j == i) continue;
            if (dist2_f32(xyzq + 4 * i, xyzq + 4 * j, ext, periodic, d) < rl2 &&
                !is_excluded(excl_start, excl_idx, i, j))
                out_idx[p++] = j;
        }
    }
    return tot;
}

static int cmp_i32(const void *a, const void *b) {
    int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}

/*
 * Cell-list Verlet list: same set, same order (rows sorted ascending) as the brute-force
 * definition, O(N).  Cells are an acceleration structure only; the accept test is the shared
 * fp32 expression, so the result is bit-identical to orc_neighbors_brute.
 * lo: box origin (periodic) or any lower bound of the coordinates (non-periodic; ext then is
 * the bounding extent).
 */
int64_t orc_neighbors_cell(int n, const float *xyzq, const float *lo, const float *ext, int periodic,
                           float r_list, const int32_t *excl_start, const int32_t *excl_idx,
                           int64_t *out_start, int32_t *out_idx, int64_t cap) {
    const float rl2 = r_list * r_list;
    int nc[3];
    double cw[3];
    for (int a = 0; a < 3; ++a) {
        /* margin: cells slightly larger than r_list so fp32 rounding of the cell assignment can
           never hide a pair that the fp32 distance test accepts */
        int m = (int)floor((double)ext[a] / ((double)r_list * 1.001 + 1e-3));
        if (m < 1) m = 1;
        if (m > 1024) m = 1024;
        nc[a] = m;
        cw[a] = (double)ext[a] / m;
    }
    const int64_t ncell = (int64_t)nc[0] * nc[1] * nc[2];
    int32_t *cell_of = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    int32_t *cstart = (int32_t *)calloc((size_t)ncell + 1, sizeof(int32_t));
    int32_t *order = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    for (int i = 0; i < n; ++i) {
        int c[3];
        for (int a = 0; a < 3; ++a) {
            double u = ((double)xyzq[4 * i + a] - (double)lo[a]) / cw[a];
            int64_t k = (int64_t)floor(u);
            if (periodic) { k %= nc[a]; if (k < 0) k += nc[a]; }
            else { if (k < 0) k = 0; if (k >= nc[a]) k = nc[a] - 1; }
            c[a] = (int)k;
        }
        cell_of[i] = (c[2] * nc[1] + c[1]) * nc[0] + c[0];
        cstart[cell_of[i] + 1]++;
    }
    for (int64_t c = 0; c < ncell; ++c) cstart[c + 1] += cstart[c];
    int32_t *cursor = (int32_t *)malloc(sizeof(int32_t) * (size_t)ncell);
    memcpy(cursor, cstart, sizeof(int32_t) * (size_t)ncell);
    for (int i = 0; i < n; ++i) order[cursor[cell_of[i]]++] = i;
    free(cursor);

    /* per-dimension offset sets, de-duplicated for tiny periodic grids */
    int noff[3], off[3][3];
    for (int a = 0; a < 3; ++a) {
        if (nc[a] >= 3 || !periodic) { noff[a] = 3; off[a][0] = -1; off[a][1] = 0; off[a][2] = 1; }
        else if (nc[a] == 2) { noff[a] = 2; off[a][0] = 0; off[a][1] = 1; }
        else { noff[a] = 1; off[a][0] = 0; }
    }

    int64_t *cnt = (int64_t *)calloc((size_t)n + 1, sizeof(int64_t));
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) {
            int64_t tot = 0;
            for (int i = 0; i < n; ++i) { out_start[i] = tot; tot += cnt[i]; }
            out_start[n] = tot;
            if (!out_idx) break;
            if (tot > cap) { free(cnt); free(cell_of); free(cstart); free(order); return -1; }
        }
#pragma omp parallel for schedule(dynamic, 64)
        for (int i = 0; i < n; ++i) {
            int ci = cell_of[i];
            int c0 = ci % nc[0], c1 = (ci / nc[0]) % nc[1], c2 = ci / (nc[0] * nc[1]);
            int64_t c = 0, p = pass ? out_start[i] : 0;
            float d[3];
            for (int iz = 0; iz < noff[2]; ++iz)
            for (int iy = 0; iy < noff[1]; ++iy)
            for (int ix = 0; ix < noff[0]; ++ix) {
                int k0 = c0 + off[0][ix], k1 = c1 + off[1][iy], k2 = c2 + off[2][iz];
                if (periodic) {
                    k0 = (k0 + nc[0]) % nc[0]; k1 = (k1 + nc[1]) % nc[1]; k2 = (k2 + nc[2]) % nc[2];
                } else if (k0 < 0 || k0 >= nc[0] || k1 < 0 || k1 >= nc[1] || k2 < 0 || k2 >= nc[2]) {
                    continue;
                }
                int cj = (k2 * nc[1] + k1) * nc[0] + k0;
                for (int32_t s = cstart[cj]; s < cstart[cj + 1]; ++s) {
                    int j = order[s];
                    if (j == i) continue;
                    if (dist2_f32(xyzq + 4 * i, xyzq + 4 * j, ext, periodic, d) < rl2 &&
                        !is_excluded(excl_start, excl_idx, i, j)) {
                        if (pass) out_idx[p++] = j;
                        ++c;
                    }
                }
            }
            if (!pass) cnt[i] = c;
            else qsort(out_idx + out_start[i], (size_t)c, sizeof(int32_t), cmp_i32);
        }
    }
    int64_t tot = out_start[n];
    free(cnt); free(cell_of); free(cstart); free(order);
    return tot;
}

/* ---- nonbonded forces over a list ------------------------------------------------------- */

typedef struct {
    float rc_lj, rc_q;     /* cutoffs (A) */
    int coul_mode;         /* ORC_COUL_* */
    float alpha;           /* Ewald splitting parameter (1/A), erfc mode */
    int lj_on, coul_on;    /* MdOverrides lj_disabled / coulomb_disabled, md/mod.rs:671-686 */
} orc_nb_params;

/* fp32 pair arithmetic, cutoff masks decided in fp32. */
static inline void pair_f32(const float *d, float r2, float sigma, float eps, float qq,
                            const orc_nb_params *p, float *f, float *e_lj, float *e_q, float *fabs_) {
    float fr = 0.f; /* |F|/r: multiply by d to get the vector */
    if (p->lj_on && eps != 0.f && r2 < p->rc_lj * p->rc_lj) {
        float r = sqrtf(r2), inv_r = 1.0f / r;
        float sr = sigma * inv_r, sr2 = sr * sr, sr4 = sr2 * sr2, sr6 = sr4 * sr2, sr12 = sr6 * sr6;
        float mag = 24.0f * eps * fmaf(2.f, sr12, -sr6) * inv_r;
        fr += mag * inv_r;
        *e_lj += 4.f * eps * (sr12 - sr6);
    }
    if (p->coul_on && p->coul_mode != ORC_COUL_NONE && qq != 0.f && r2 < p->rc_q * p->rc_q) {
        float r = sqrtf(r2);
        if (p->coul_mode == ORC_COUL_PLAIN) {
            float mag = qq / (r2 + ORC_SOFTENING_SQ);
            fr += mag / r;
            *e_q += qq / r;
        } else {
            float ar = p->alpha * r;
            float erfc_ar = erfcf(ar);
            float mag = qq * (erfc_ar / r2 + 2.f * p->alpha * (float)ORC_INV_SQRT_PI * expf(-ar * ar) / r);
            fr += mag / r;
            *e_q += qq * erfc_ar / r;
        }
    }
    f[0] += d[0] * fr; f[1] += d[1] * fr; f[2] += d[2] * fr;
    fabs_[0] += fabsf(fr) * sqrtf(r2);
    fabs_[1] += fabsf(fr) * sqrtf(r2);
}

/* fp64 pair arithmetic from the fp32 positions; the cutoff masks still use the fp32 r2. */
static inline void pair_f64(const double *d, float r2_f32, double sigma, double eps, double qq,
                            const orc_nb_params *p, double *f, double *e_lj, double *e_q, double *fabs_) {
    double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
    double fr = 0.0, fa = 0.0, ft = 0.0;
    if (p->lj_on && eps != 0.0 && r2_f32 < p->rc_lj * p->rc_lj) {
        double r = sqrt(r2), sr = sigma / r, sr2 = sr * sr, sr6 = sr2 * sr2 * sr2, sr12 = sr6 * sr6;
        double mag = 24.0 * eps * (2.0 * sr12 - sr6) / r;
        fr += mag / r; fa += fabs(mag);
        ft += 24.0 * fabs(eps) * (2.0 * sr12 + sr6) / r; /* |repulsive term| + |attractive term| */
        *e_lj += 4.0 * eps * (sr12 - sr6);
    }
    if (p->coul_on && p->coul_mode != ORC_COUL_NONE && qq != 0.0 && r2_f32 < p->rc_q * p->rc_q) {
        double r = sqrt(r2), mag;
        if (p->coul_mode == ORC_COUL_PLAIN) {
            mag = qq / (r2 + (double)ORC_SOFTENING_SQ);
            *e_q += qq / r;
        } else {
            double a = (double)p->alpha, ar = a * r;
            mag = qq * (erfc(ar) / r2 + 2.0 * a * ORC_INV_SQRT_PI * exp(-ar * ar) / r);
            *e_q += qq * erfc(ar) / r;
        }
        fr += mag / r; fa += fabs(mag); ft += fabs(mag);
    }
    f[0] += d[0] * fr; f[1] += d[1] * fr; f[2] += d[2] * fr;
    fabs_[0] += fa;
    fabs_[1] += ft;
}

/*
 * Forces on every atom from its (full) neighbour row; energy per atom = sum over the row of
 * the pair energy (so the system energy is half the total).  ljtab: T*T pairs (sigma, eps).
 * out_f: n*4 floats (fx, fy, fz, e_i).  out_sumabs: n floats, sum_j |f_ij| of the NET pair
 * forces; out_sumterms: n floats, sum_j (|LJ repulsive term| + |LJ attractive term| + |Coulomb
 * term|) -- the scale an fp32 evaluation can be held to (the 12 and 6 terms cancel near the LJ
 * minimum); either may be NULL.  out_energy: {E_lj, E_coulomb} system totals (already halved), fp64.
 * precision: 32 -> fp32 arithmetic in row order; 64 -> fp64 arithmetic (truth); 6432 -> fp64 arithmetic on the
 * reference's fp32 minimum-image differences.
 */
void orc_forces(int n, const float *xyzq, const uint16_t *type, int T, const float *ljtab,
                const float *ext, int periodic, const orc_nb_params *p,
                const int64_t *nbr_start, const int32_t *nbr_idx, int precision,
                float *out_f, float *out_sumabs, float *out_sumterms, double *out_energy) {
    double e_lj_tot = 0.0, e_q_tot = 0.0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : e_lj_tot, e_q_tot)
    for (int i = 0; i < n; ++i) {
        const float *xi = xyzq + 4 * i;
        const int ti = type ? type[i] : 0;
        if (precision == 32) {
            float f[3] = {0, 0, 0}, e_lj = 0, e_q = 0, fa[2] = {0, 0}, d[3];
            for (int64_t k = nbr_start[i]; k < nbr_start[i + 1]; ++k) {
                int j = nbr_idx[k];
                const float *xj = xyzq + 4 * j;
                float r2 = dist2_f32(xi, xj, ext, periodic, d);
                const float *lj = ljtab + 2 * ((size_t)ti * T + (type ? type[j] : 0));
                pair_f32(d, r2, lj[0], lj[1], xi[3] * xj[3], p, f, &e_lj, &e_q, fa);
            }
            out_f[4 * i] = f[0]; out_f[4 * i + 1] = f[1]; out_f[4 * i + 2] = f[2];
            out_f[4 * i + 3] = e_lj + e_q;
            if (out_sumabs) out_sumabs[i] = fa[0];
            if (out_sumterms) out_sumterms[i] = fa[1];
            e_lj_tot += e_lj; e_q_tot += e_q;
        } else {
            double f[3] = {0, 0, 0}, e_lj = 0, e_q = 0, fa[2] = {0, 0}, d[3];
            float df[3];
            for (int64_t k = nbr_start[i]; k < nbr_start[i + 1]; ++k) {
                int j = nbr_idx[k];
                const float *xj = xyzq + 4 * j;
                float r2 = dist2_f32(xi, xj, ext, periodic, df);
                for (int a = 0; a < 3; ++a) {
                    double dd = (double)xi[a] - (double)xj[a];
                    if (periodic) dd -= rint(dd / (double)ext[a]) * (double)ext[a];
                    /* precision 6432: fp64 arithmetic on the fp32 minimum-image difference the reference itself forms
                     * (float3 diff = posit_tgt - posit_src, util.cu:65-71 / cuda.cu:85-91): isolates the arithmetic from
                     * the ulp(L) rounding of a difference taken across the periodic seam, which both sides share */
                    d[a] = precision == 6432 ? (double)df[a] : dd;
                }
                const float *lj = ljtab + 2 * ((size_t)ti * T + (type ? type[j] : 0));
                pair_f64(d, r2, lj[0], lj[1], (double)xi[3] * (double)xj[3], p, f, &e_lj, &e_q, fa);
            }
            out_f[4 * i] = (float)f[0]; out_f[4 * i + 1] = (float)f[1]; out_f[4 * i + 2] = (float)f[2];
            out_f[4 * i + 3] = (float)(e_lj + e_q);
            if (out_sumabs) out_sumabs[i] = (float)fa[0];
            if (out_sumterms) out_sumterms[i] = (float)fa[1];
            e_lj_tot += e_lj; e_q_tot += e_q;
        }
    }
    if (out_energy) { out_energy[0] = 0.5 * e_lj_tot; out_energy[1] = 0.5 * e_q_tot; }
}

/*
 * Amber 1-4 pairs: listed explicitly, excluded from the Verlet list, evaluated without cutoff
 * with LJ scaled by scale_lj (0.5) and Coulomb by scale_q (1/1.2)  (SURVEY 8c).
 * pairs: npairs*2 atom ids.  Accumulates into f (n*4 floats) and energy[2] (fp64).
 */
void orc_pairs14(int npairs, const int32_t *pairs, const float *xyzq, const uint16_t *type, int T,
                 const float *ljtab, const float *ext, int periodic, float scale_lj, float scale_q,
                 int lj_on, int coul_on, float *f, double *energy, float *sumabs, float *sumterms) {
    for (int k = 0; k < npairs; ++k) {
        int i = pairs[2 * k], j = pairs[2 * k + 1];
        double d[3], r2 = 0;
        for (int a = 0; a < 3; ++a) {
            double dd = (double)xyzq[4 * i + a] - (double)xyzq[4 * j + a];
            if (periodic) dd -= rint(dd / (double)ext[a]) * (double)ext[a];
            d[a] = dd; r2 += dd * dd;
        }
        const float *lj = ljtab + 2 * ((size_t)(type ? type[i] : 0) * T + (type ? type[j] : 0));
        double r = sqrt(r2), fr = 0, e = 0, fa = 0, ft = 0;
        if (lj_on && lj[1] != 0.f) {
            double sr = lj[0] / r, sr6 = pow(sr, 6), sr12 = sr6 * sr6;
            fr += scale_lj * 24.0 * lj[1] * (2.0 * sr12 - sr6) / r2;
            fa += fabs(scale_lj * 24.0 * lj[1] * (2.0 * sr12 - sr6) / r);
            ft += fabs(scale_lj * 24.0 * lj[1]) * (2.0 * sr12 + sr6) / r;
            e += scale_lj * 4.0 * lj[1] * (sr12 - sr6);
            if (energy) energy[0] += scale_lj * 4.0 * lj[1] * (sr12 - sr6);
        }
        if (coul_on) {
            double qq = (double)xyzq[4 * i + 3] * (double)xyzq[4 * j + 3];
            fr += scale_q * qq / (r2 + (double)ORC_SOFTENING_SQ) / r;
            fa += fabs(scale_q * qq / (r2 + (double)ORC_SOFTENING_SQ));
            ft += fabs(scale_q * qq / (r2 + (double)ORC_SOFTENING_SQ));
            e += scale_q * qq / r;
            if (energy) energy[1] += scale_q * qq / r;
        }
        for (int a = 0; a < 3; ++a) {
            f[4 * i + a] += (float)(d[a] * fr);
            f[4 * j + a] -= (float)(d[a] * fr);
        }
        /* per-atom energy keeps the "row sum" convention: each end carries the full pair energy */
        f[4 * i + 3] += (float)e;
        f[4 * j + 3] += (float)e;
        if (sumabs) { sumabs[i] += (float)fa; sumabs[j] += (float)fa; }
        if (sumterms) { sumterms[i] += (float)ft; sumterms[j] += (float)ft; }
    }
}

/* Harmonic bonds E = k (r - r0)^2 (Amber convention, no 1/2).  Oracle-only helper so the C1
 * flexible-water NVE plumbing run is well posed; not part of the CUDA path. */
void orc_bonds(int nb, const int32_t *bonds, const float *kr0, const float *xyzq, const float *ext,
               int periodic, float *f, double *energy) {
    for (int b = 0; b < nb; ++b) {
        int i = bonds[2 * b], j = bonds[2 * b + 1];
        double d[3], r2 = 0;
        for (int a = 0; a < 3; ++a) {
            double dd = (double)xyzq[4 * i + a] - (double)xyzq[4 * j + a];
            if (periodic) dd -= rint(dd / (double)ext[a]) * (double)ext[a];
            d[a] = dd; r2 += dd * dd;
        }
        double r = sqrt(r2), k = kr0[2 * b], r0 = kr0[2 * b + 1];
        double fr = -2.0 * k * (r - r0) / r;
        for (int a = 0; a < 3; ++a) {
            f[4 * i + a] += (float)(d[a] * fr);
            f[4 * j + a] -= (float)(d[a] * fr);
        }
        if (energy) *energy += k * (r - r0) * (r - r0);
    }
}

/* ---- bonded terms beyond bonds, fp64 ------------------------------------------
 * Amber functional forms (the reference's force field, README.md:234-241; the evaluating code is in the
 * un-vendored `dynamics` crate -> parity unpinned, known-answer + finite-difference tests pin the
 * arithmetic):  angle E = k (theta - theta0)^2, dihedral E = pk (1 + cos(n phi - phase)), IUPAC phi.
 * Written independently of molchanica_b200/csrc/bonded_terms.h: gradients of theta and phi by the
 * chain rule on cos(theta) and on atan2, not the GROMACS vector form the device code uses. */
static void v_sub(const float *x, int i, int j, const float *ext, int periodic, double *d) {
    for (int a = 0; a < 3; ++a) {
        double dd = (double)x[4 * i + a] - (double)x[4 * j + a];
        if (periodic) dd -= rint(dd / (double)ext[a]) * (double)ext[a];
        d[a] = dd;
    }
}
static void v_cross(const double *a, const double *b, double *c) {
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
static double v_dot(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* forces f64: n*3 doubles accumulated; energy3: {bond, angle, dihedral} accumulated */
void orc_bonded64(int nb, const int32_t *bonds, const float *kr0, int na, const int32_t *angles, const float *kt0,
                  int nd, const int32_t *dih, const float *prm, const float *xyzq, const float *ext, int periodic,
                  double *f, double *energy3) {
    for (int t = 0; t < nb; ++t) {
        int i = bonds[2 * t], j = bonds[2 * t + 1];
        double d[3];
        v_sub(xyzq, i, j, ext, periodic, d);
        double r = sqrt(v_dot(d, d)), k = kr0[2 * t], r0 = kr0[2 * t + 1];
        double fr = -2.0 * k * (r - r0) / r;
        for (int a = 0; a < 3; ++a) { f[3 * i + a] += d[a] * fr; f[3 * j + a] -= d[a] * fr; }
        energy3[0] += k * (r - r0) * (r - r0);
    }
    for (int t = 0; t < na; ++t) {
        int i = angles[3 * t], j = angles[3 * t + 1], k = angles[3 * t + 2];
        double a[3], b[3];
        v_sub(xyzq, i, j, ext, periodic, a);
        v_sub(xyzq, k, j, ext, periodic, b);
        double la = sqrt(v_dot(a, a)), lb = sqrt(v_dot(b, b));
        double c = v_dot(a, b) / (la * lb);
        if (c > 1.0) c = 1.0;
        if (c < -1.0) c = -1.0;
        double th = acos(c), kk = kt0[2 * t], th0 = kt0[2 * t + 1];
        double dE = 2.0 * kk * (th - th0);            /* dE/dtheta */
        double s = sqrt(1.0 - c * c);
        if (s < 1e-9) s = 1e-9;
        double dth_dc = -1.0 / s;                     /* dtheta/dcos */
        for (int x = 0; x < 3; ++x) {
            double dc_da = b[x] / (la * lb) - c * a[x] / (la * la);
            double dc_db = a[x] / (la * lb) - c * b[x] / (lb * lb);
            double fi = -dE * dth_dc * dc_da, fk = -dE * dth_dc * dc_db;
            f[3 * i + x] += fi; f[3 * k + x] += fk; f[3 * j + x] -= fi + fk;
        }
        energy3[1] += kk * (th - th0) * (th - th0);
    }
    for (int t = 0; t < nd; ++t) {
        int i = dih[4 * t], j = dih[4 * t + 1], k = dih[4 * t + 2], l = dih[4 * t + 3];
        double pk = prm[3 * t], per = prm[3 * t + 1], ph = prm[3 * t + 2];
        /* numerical-free analytic gradient through y = |r_kj| r_ij.n, x = m.n is long; use the textbook
         * projection form (Blondel & Karplus 1996): F_i = -dE/dphi * (-|G| / |A|^2) A with F = r_i - r_j,
         * G = r_j - r_k, H = r_l - r_k, A = F x G, B = H x G */
        double F[3], G[3], H[3], A[3], B[3];
        v_sub(xyzq, i, j, ext, periodic, F);
        v_sub(xyzq, j, k, ext, periodic, G);
        v_sub(xyzq, l, k, ext, periodic, H);
        v_cross(F, G, A);
        v_cross(H, G, B);
        double A2 = v_dot(A, A), B2 = v_dot(B, B), g = sqrt(v_dot(G, G));
        double cosphi = v_dot(A, B) / sqrt(A2 * B2);
        double BxA[3];
        v_cross(B, A, BxA);
        double sinphi = v_dot(BxA, G) / (sqrt(A2 * B2) * g);
        double phi = atan2(sinphi, cosphi);
        double dE = -pk * per * sin(per * phi - ph);  /* dE/dphi */
        double FG = v_dot(F, G), HG = v_dot(H, G);
        for (int x = 0; x < 3; ++x) {
            double dphi_di = -g / A2 * A[x];
            double dphi_dl = g / B2 * B[x];
            double dphi_dj = g / A2 * A[x] + FG / (A2 * g) * A[x] - HG / (B2 * g) * B[x];
            double dphi_dk = -g / B2 * B[x] - FG / (A2 * g) * A[x] + HG / (B2 * g) * B[x];
            f[3 * i + x] -= dE * dphi_di; f[3 * j + x] -= dE * dphi_dj;
            f[3 * k + x] -= dE * dphi_dk; f[3 * l + x] -= dE * dphi_dl;
        }
        energy3[2] += pk * (1.0 + cos(per * phi - ph));
    }
}

/* ---- velocity Verlet -------------------------------------------------------------------- */

/* v += F * inv_mass * (dt/2) * 418.4 ; vel: n*4 (vx,vy,vz,inv_mass); static atoms: inv_mass 0 */
void orc_kick(int n, float *vel, const float *f, float half_dt) {
    for (int i = 0; i < n; ++i) {
        float s = vel[4 * i + 3] * half_dt * ORC_ACCEL_CONV;
        vel[4 * i] += f[4 * i] * s; vel[4 * i + 1] += f[4 * i + 1] * s; vel[4 * i + 2] += f[4 * i + 2] * s;
    }
}

/* x += v dt; returns the largest squared displacement from xref (n*4) */
float orc_drift(int n, float *xyzq, const float *vel, float dt, const float *xref) {
    float worst = 0.f;
    for (int i = 0; i < n; ++i) {
        float d2 = 0.f;
        for (int a = 0; a < 3; ++a) {
            xyzq[4 * i + a] += vel[4 * i + a] * dt;
            if (xref) { float d = xyzq[4 * i + a] - xref[4 * i + a]; d2 += d * d; }
        }
        if (d2 > worst) worst = d2;
    }
    return worst;
}

/* kinetic energy in kcal/mol: 1/2 m v^2 / 418.4 */
double orc_kinetic(int n, const float *vel) {
    double ke = 0;
    for (int i = 0; i < n; ++i) {
        if (vel[4 * i + 3] == 0.f) continue;
        double m = 1.0 / vel[4 * i + 3];
        ke += 0.5 * m * ((double)vel[4 * i] * vel[4 * i] + (double)vel[4 * i + 1] * vel[4 * i + 1] +
                         (double)vel[4 * i + 2] * vel[4 * i + 2]);
    }
    return ke / (double)ORC_ACCEL_CONV;
}

/* ---- rigid three-site water, fp64 SHAKE ------------------------------------------
 * The reference keeps water rigid with SETTLE (README.md:239; the code is in the un-vendored `dynamics` crate ->
 * parity unpinned).  The oracle solves the same constraint equations iteratively (SHAKE, Ryckaert et al. 1977,
 * displacements along the OLD bond vectors, converged to 1e-13) -- an independent route to the positions the
 * analytic SETTLE of the CUDA path produces.  Set once with orc_set_rigid_waters; orc_md_run then constrains
 * after every drift and corrects the velocities by the position change / dt (leap-frog form). */
static int g_nw = 0;
static const int32_t *g_waters = NULL;
static double g_doh = 0, g_dhh = 0;
void orc_set_rigid_waters(int n, const int32_t *triples, float d_oh, float d_hh) {
    g_nw = n; g_waters = triples; g_doh = d_oh; g_dhh = d_hh;
}

/* xold / xnew: n*4 floats; constrains xnew in place, adds (x'' - x')/dt to vel */
void orc_shake_waters(const float *xold, float *xnew, float *vel, const float *ext, int periodic, float dt) {
    for (int w = 0; w < g_nw; ++w) {
        const int id[3] = {g_waters[3 * w], g_waters[3 * w + 1], g_waters[3 * w + 2]};
        double r0[3][3], p[3][3], m[3];
        for (int k = 0; k < 3; ++k) {
            m[k] = 1.0 / (double)vel[4 * id[k] + 3];
            for (int a = 0; a < 3; ++a) {
                /* everything relative to the old oxygen, minimum image per atom */
                double d0 = (double)xold[4 * id[k] + a] - (double)xold[4 * id[0] + a];
                double d1 = (double)xnew[4 * id[k] + a] - (double)xold[4 * id[0] + a];
                if (periodic) { d0 -= rint(d0 / (double)ext[a]) * (double)ext[a]; d1 -= rint(d1 / (double)ext[a]) * (double)ext[a]; }
                r0[k][a] = d0; p[k][a] = d1;
            }
        }
        double q[3][3];
        memcpy(q, p, sizeof(q));
        const int ci[3] = {0, 0, 1}, cj[3] = {1, 2, 2};
        const double cd[3] = {g_doh, g_doh, g_dhh};
        for (int it = 0; it < 1000; ++it) {
            double worst = 0;
            for (int c = 0; c < 3; ++c) {
                int i = ci[c], j = cj[c];
                double s[3], r[3], ss = 0, sr = 0;
                for (int a = 0; a < 3; ++a) { s[a] = q[i][a] - q[j][a]; r[a] = r0[i][a] - r0[j][a]; ss += s[a] * s[a]; sr += s[a] * r[a]; }
                double diff = cd[c] * cd[c] - ss;
                if (fabs(diff) / (cd[c] * cd[c]) > worst) worst = fabs(diff) / (cd[c] * cd[c]);
                double g = diff / (2.0 * sr * (1.0 / m[i] + 1.0 / m[j]));
                for (int a = 0; a < 3; ++a) { q[i][a] += g * r[a] / m[i]; q[j][a] -= g * r[a] / m[j]; }
            }
            if (worst < 1e-13) break;
        }
        for (int k = 0; k < 3; ++k)
            for (int a = 0; a < 3; ++a) {
                double d = q[k][a] - p[k][a];
                xnew[4 * id[k] + a] = (float)((double)xnew[4 * id[k] + a] + d);
                vel[4 * id[k] + a] = (float)((double)vel[4 * id[k] + a] + d / (double)dt);
            }
    }
}

/* ---- Langevin thermostat, fp64 O step with counter-based noise -------------------------
 * v <- c1 v + sqrt(1 - c1^2) sqrt(kT/m) xi after the drift (and constraints) of every step.  xi comes from
 * Philox4x32-10 (Salmon et al., SC'11; this is an implementation of the published algorithm written
 * independently of molchanica_b200/csrc/langevin_terms.h and pinned by its known-answer vectors) keyed by
 * (seed, atom id, step), Box-Muller in double on the same 24-bit uniforms the CUDA path uses. */
static int g_lgv = 0;
static double g_lgv_kT = 0, g_lgv_gamma = 0;
static uint64_t g_lgv_seed = 0, g_lgv_step = 0;
void orc_set_langevin(int on, float temperature_k, float gamma_per_ps, uint64_t seed) {
    g_lgv = on; g_lgv_kT = 0.0019872041 * (double)temperature_k; g_lgv_gamma = gamma_per_ps; g_lgv_seed = seed; g_lgv_step = 0;
}

void orc_philox4x32_10(const uint32_t *ctr, const uint32_t *key, uint32_t *out) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]}, k[2] = {key[0], key[1]};
    for (int round = 0; round < 10; ++round) {
        uint64_t lo_prod = (uint64_t)c[0] * 0xD2511F53ull, hi_prod = (uint64_t)c[2] * 0xCD9E8D57ull;
        uint32_t t[4];
        t[0] = (uint32_t)(hi_prod >> 32) ^ c[1] ^ k[0];
        t[1] = (uint32_t)(hi_prod & 0xffffffffull);
        t[2] = (uint32_t)(lo_prod >> 32) ^ c[3] ^ k[1];
        t[3] = (uint32_t)(lo_prod & 0xffffffffull);
        memcpy(c, t, sizeof(c));
        k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
    }
    memcpy(out, c, 4 * sizeof(uint32_t));
}

void orc_langevin_normals(uint64_t seed, uint32_t atom, uint64_t step, double *xi) {
    uint32_t ctr[4] = {atom, (uint32_t)(step & 0xffffffffull), (uint32_t)(step >> 32), 0}, key[2] = {(uint32_t)(seed & 0xffffffffull), (uint32_t)(seed >> 32)}, r[4];
    orc_philox4x32_10(ctr, key, r);
    double u[4];
    for (int j = 0; j < 4; ++j) u[j] = ((double)(r[j] >> 8) + 1.0) / 16777216.0;
    double ra = sqrt(-2.0 * log(u[0])), rb = sqrt(-2.0 * log(u[2]));
    xi[0] = ra * cos(6.283185307179586 * u[1]);
    xi[1] = ra * sin(6.283185307179586 * u[1]);
    xi[2] = rb * cos(6.283185307179586 * u[3]);
}

static void orc_langevin_step(int n, float *vel, float dt) {
    const double c1 = exp(-g_lgv_gamma * (double)dt), c2 = sqrt(1.0 - c1 * c1);
    for (int i = 0; i < n; ++i) {
        const double im = vel[4 * i + 3];
        if (im <= 0) continue;
        double xi[3];
        orc_langevin_normals(g_lgv_seed, (uint32_t)i, g_lgv_step, xi);
        const double s = c2 * sqrt(g_lgv_kT * im * 418.4);
        for (int a = 0; a < 3; ++a) vel[4 * i + a] = (float)(c1 * (double)vel[4 * i + a] + s * xi[a]);
    }
    ++g_lgv_step;
}

/* ---- velocity stage of RATTLE (Andersen 1983) for the clusters above, fp64: after every closing half kick the velocity
 * components along the constrained bonds are removed by solving the <= 3 x 3 linear system of the cluster (Gaussian
 * elimination here; molchanica_b200/csrc/rattle_terms.h uses Cramer's rule in fp32). */
static void orc_rattle_cluster(int nat, const int *id, int nc, const int *ci, const int *cj, const float *xyzq, float *vel,
                               const float *ext, int periodic) {
    double r[4][3], rk[3][3], A[3][4], lam[3];
    for (int k = 0; k < nat; ++k)
        for (int a = 0; a < 3; ++a) {
            double d = (double)xyzq[4 * id[k] + a] - (double)xyzq[4 * id[0] + a];
            if (periodic) d -= rint(d / (double)ext[a]) * (double)ext[a];
            r[k][a] = d;
        }
    for (int k = 0; k < nc; ++k) {
        A[k][3] = 0;
        for (int a = 0; a < 3; ++a) {
            rk[k][a] = r[ci[k]][a] - r[cj[k]][a];
            A[k][3] += rk[k][a] * ((double)vel[4 * id[ci[k]] + a] - (double)vel[4 * id[cj[k]] + a]);
        }
    }
    for (int k = 0; k < nc; ++k)
        for (int l = 0; l < nc; ++l) {
            double si = (ci[k] == ci[l]) - (ci[k] == cj[l]), sj = (cj[k] == ci[l]) - (cj[k] == cj[l]);
            A[k][l] = (si * (double)vel[4 * id[ci[k]] + 3] - sj * (double)vel[4 * id[cj[k]] + 3]) *
                      (rk[l][0] * rk[k][0] + rk[l][1] * rk[k][1] + rk[l][2] * rk[k][2]);
        }
    for (int p = 0; p < nc; ++p) {  /* elimination with partial pivoting */
        int best = p;
        for (int q = p + 1; q < nc; ++q) if (fabs(A[q][p]) > fabs(A[best][p])) best = q;
        for (int c = 0; c < 4; ++c) { double t = A[p][c]; A[p][c] = A[best][c]; A[best][c] = t; }
        for (int q = p + 1; q < nc; ++q) {
            double f = A[q][p] / A[p][p];
            for (int c = p; c < 4; ++c) A[q][c] -= f * A[p][c];
        }
    }
    for (int p = nc - 1; p >= 0; --p) {
        double t = A[p][3];
        for (int c = p + 1; c < nc; ++c) t -= A[p][c] * lam[c];
        lam[p] = t / A[p][p];
    }
    for (int l = 0; l < nc; ++l)
        for (int a = 0; a < 3; ++a) {
            vel[4 * id[ci[l]] + a] = (float)((double)vel[4 * id[ci[l]] + a] - (double)vel[4 * id[ci[l]] + 3] * lam[l] * rk[l][a]);
            vel[4 * id[cj[l]] + a] = (float)((double)vel[4 * id[cj[l]] + a] + (double)vel[4 * id[cj[l]] + 3] * lam[l] * rk[l][a]);
        }
}

static const int32_t *g_hclusters_fwd(void);
static int g_nhc_fwd(void);
void orc_rattle_velocities(const float *xyzq, float *vel, const float *ext, int periodic) {
    const int wi[3] = {0, 0, 1}, wj[3] = {1, 2, 2}, hi[3] = {0, 0, 0}, hj[3] = {1, 2, 3};
    for (int w = 0; w < g_nw; ++w) {
        const int id[4] = {g_waters[3 * w], g_waters[3 * w + 1], g_waters[3 * w + 2], 0};
        orc_rattle_cluster(3, id, 3, wi, wj, xyzq, vel, ext, periodic);
    }
    const int32_t *hc = g_hclusters_fwd();
    for (int c = 0; c < g_nhc_fwd(); ++c) {
        int id[4], nat = 0;
        for (int k = 0; k < 4; ++k) if (hc[4 * c + k] >= 0) id[nat++] = hc[4 * c + k];
        if (nat >= 2) orc_rattle_cluster(nat, id, nat - 1, hi, hj, xyzq, vel, ext, periodic);
    }
}

/* ---- SHAKE for bonds to hydrogen, fp64 ------------------------------------------------------
 * Clusters (heavy, h1, h2, h3; -1 = unused) with one length per hydrogen; the reference constrains bonds to hydrogen
 * at 2 fs (ui/panels/md.rs:362-371; code in the un-vendored `dynamics` crate -> parity unpinned).  Same equations as
 * molchanica_b200/csrc/shake_terms.h, solved here in double, sweeping the constraints in the opposite order and to
 * 1e-13: the fixed point does not depend on the sweep order. */
static int g_nhc = 0;
static const int32_t *g_hclusters = NULL;
static const float *g_hdist = NULL;
void orc_set_hbond_constraints(int n, const int32_t *clusters, const float *lengths) { g_nhc = n; g_hclusters = clusters; g_hdist = lengths; }
static const int32_t *g_hclusters_fwd(void) { return g_hclusters; }
static int g_nhc_fwd(void) { return g_nhc; }

void orc_shake_h(const float *xold, float *xnew, float *vel, const float *ext, int periodic, float dt) {
    for (int c = 0; c < g_nhc; ++c) {
        const int hv = g_hclusters[4 * c];
        int hid[3], nh = 0;
        double dl[3];
        for (int k = 0; k < 3; ++k)
            if (g_hclusters[4 * c + 1 + k] >= 0) { hid[nh] = g_hclusters[4 * c + 1 + k]; dl[nh] = g_hdist[3 * c + k]; ++nh; }
        double r0[3][3], p[3][3], q[3][3], p0[3], q0[3], im0 = vel[4 * hv + 3], im[3];
        for (int a = 0; a < 3; ++a) {
            double d = (double)xnew[4 * hv + a] - (double)xold[4 * hv + a];
            if (periodic) d -= rint(d / (double)ext[a]) * (double)ext[a];
            p0[a] = q0[a] = d;
        }
        for (int k = 0; k < nh; ++k) {
            im[k] = vel[4 * hid[k] + 3];
            for (int a = 0; a < 3; ++a) {
                double d0 = (double)xold[4 * hid[k] + a] - (double)xold[4 * hv + a], d1 = (double)xnew[4 * hid[k] + a] - (double)xold[4 * hv + a];
                if (periodic) { d0 -= rint(d0 / (double)ext[a]) * (double)ext[a]; d1 -= rint(d1 / (double)ext[a]) * (double)ext[a]; }
                r0[k][a] = d0; p[k][a] = q[k][a] = d1;
            }
        }
        for (int it = 0; it < 2000; ++it) {
            double worst = 0;
            for (int k = nh - 1; k >= 0; --k) {
                double s[3], ss = 0, sr = 0;
                for (int a = 0; a < 3; ++a) { s[a] = p[k][a] - p0[a]; ss += s[a] * s[a]; sr += s[a] * r0[k][a]; }
                double diff = dl[k] * dl[k] - ss;
                if (fabs(diff) / (dl[k] * dl[k]) > worst) worst = fabs(diff) / (dl[k] * dl[k]);
                double g = diff / (2.0 * sr * (im0 + im[k]));
                for (int a = 0; a < 3; ++a) { p[k][a] += g * r0[k][a] * im[k]; p0[a] -= g * r0[k][a] * im0; }
            }
            if (worst < 1e-13) break;
        }
        for (int a = 0; a < 3; ++a) {
            double d = p0[a] - q0[a];
            xnew[4 * hv + a] = (float)((double)xnew[4 * hv + a] + d);
            vel[4 * hv + a] = (float)((double)vel[4 * hv + a] + d / (double)dt);
        }
        for (int k = 0; k < nh; ++k)
            for (int a = 0; a < 3; ++a) {
                double d = p[k][a] - q[k][a];
                xnew[4 * hid[k] + a] = (float)((double)xnew[4 * hid[k] + a] + d);
                vel[4 * hid[k] + a] = (float)((double)vel[4 * hid[k] + a] + d / (double)dt);
            }
    }
}

/* ---- virtual sites of four-site water: M = O + a (H1 - O) + b (H2 - O) ------------------
 * placed after every drift (+ constraints), its force handed to the parents after every evaluation
 * (the reference's md.water {o, h0, h1, m}, properties/sol_shrinking_box.rs:605-613; parity unpinned). */
static int g_nv = 0;
static const int32_t *g_vsites = NULL;
static double g_va = 0, g_vb = 0;
void orc_set_virtual_sites(int n, const int32_t *quads, float a, float b) { g_nv = n; g_vsites = quads; g_va = a; g_vb = b; }

void orc_vsite_construct(float *xyzq, const float *ext, int periodic) {
    for (int v = 0; v < g_nv; ++v) {
        const int m = g_vsites[4 * v], o = g_vsites[4 * v + 1], h1 = g_vsites[4 * v + 2], h2 = g_vsites[4 * v + 3];
        for (int a = 0; a < 3; ++a) {
            double d1 = (double)xyzq[4 * h1 + a] - (double)xyzq[4 * o + a], d2 = (double)xyzq[4 * h2 + a] - (double)xyzq[4 * o + a];
            if (periodic) { d1 -= rint(d1 / (double)ext[a]) * (double)ext[a]; d2 -= rint(d2 / (double)ext[a]) * (double)ext[a]; }
            xyzq[4 * m + a] = (float)((double)xyzq[4 * o + a] + g_va * d1 + g_vb * d2);
        }
    }
}

void orc_vsite_spread(float *f) {
    for (int v = 0; v < g_nv; ++v) {
        const int m = g_vsites[4 * v], o = g_vsites[4 * v + 1], h1 = g_vsites[4 * v + 2], h2 = g_vsites[4 * v + 3];
        for (int a = 0; a < 3; ++a) {
            const double fm = f[4 * m + a];
            f[4 * o + a] = (float)((double)f[4 * o + a] + (1.0 - g_va - g_vb) * fm);
            f[4 * h1 + a] = (float)((double)f[4 * h1 + a] + g_va * fm);
            f[4 * h2 + a] = (float)((double)f[4 * h2 + a] + g_vb * fm);
            f[4 * m + a] = 0.f;
        }
    }
}

/* ---- CSVR thermostat: Bussi, Donadio & Parrinello, J. Chem. Phys. 126, 014101 (2007) ------
 * One scale factor per step from the kinetic energy, K' = K + (1-c)(Kbar (R1^2 + S)/Nf - K) + 2 R1 sqrt(K Kbar/Nf (1-c) c).
 * The draw sequence follows the specification in molchanica_b200/csrc/csvr_terms.h (Philox counter = (draw, step,
 * 0xC5A1), 53-bit uniforms, Box-Muller, Marsaglia-Tsang gamma) so that both sides get the same factor; the code is
 * written independently on top of this file's own Philox. */
static int g_csvr = 0;
static double g_csvr_kT = 0, g_csvr_inv_tau = 0, g_csvr_dof_removed = 0;
static uint64_t g_csvr_seed = 0, g_csvr_step = 0;
void orc_set_csvr(int on, float temperature_k, float inv_tau, uint64_t seed, double dof_removed) {
    g_csvr = on; g_csvr_kT = 0.0019872041 * (double)temperature_k; g_csvr_inv_tau = inv_tau; g_csvr_seed = seed; g_csvr_step = 0;
    g_csvr_dof_removed = dof_removed;
}

typedef struct { uint64_t seed, step; uint32_t draw; } orc_rng;
static void rng_words(orc_rng *g, uint32_t *r) {
    uint32_t ctr[4] = {g->draw++, (uint32_t)(g->step & 0xffffffffull), (uint32_t)(g->step >> 32), 0xC5A1u};
    uint32_t key[2] = {(uint32_t)(g->seed & 0xffffffffull), (uint32_t)(g->seed >> 32)};
    orc_philox4x32_10(ctr, key, r);
}
static double rng_u53(uint32_t hi, uint32_t lo) { return ((double)((((uint64_t)hi) << 21) ^ (uint64_t)(lo >> 11)) + 0.5) / 9007199254740992.0; }
static void rng_normals(orc_rng *g, double *a, double *b) {
    uint32_t r[4];
    rng_words(g, r);
    double u0 = rng_u53(r[0], r[1]), u1 = rng_u53(r[2], r[3]), rad = sqrt(-2.0 * log(u0));
    *a = rad * cos(6.283185307179586 * u1); *b = rad * sin(6.283185307179586 * u1);
}
static double rng_gamma(orc_rng *g, double k) {
    double d = k - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * d);
    for (int trial = 0; trial < 1000; ++trial) {
        double x, y;
        rng_normals(g, &x, &y);
        double t = 1.0 + cc * x;
        if (t <= 0.0) continue;
        double v = t * t * t;
        uint32_t r[4];
        rng_words(g, r);
        if (log(rng_u53(r[0], r[1])) < 0.5 * x * x + d - d * v + d * log(v)) return d * v;
    }
    return d;
}
double orc_csvr_lambda(double kinetic, double kT, double nf, double c, uint64_t seed, uint64_t step) {
    if (!(kinetic > 0.0) || nf < 1.0) return 1.0;
    orc_rng g = {seed, step, 0};
    double r1, y, s = 0.0;
    rng_normals(&g, &r1, &y);
    if (nf > 1.0) {
        if (nf - 1.0 >= 2.0) s = 2.0 * rng_gamma(&g, 0.5 * (nf - 1.0));
        else { double a, b; rng_normals(&g, &a, &b); s = a * a; }
    }
    double kbar = 0.5 * nf * kT;
    double knew = kinetic + (1.0 - c) * (kbar * (r1 * r1 + s) / nf - kinetic) + 2.0 * r1 * sqrt(kinetic * kbar / nf * (1.0 - c) * c);
    if (knew < 0.0) knew = 0.0;
    return sqrt(knew / kinetic);
}

static void orc_csvr_step(int n, float *vel, float dt) {
    double ke = 0, mobile = 0;
    for (int i = 0; i < n; ++i) {
        if (vel[4 * i + 3] <= 0.f) continue;
        double m = 1.0 / (double)vel[4 * i + 3];
        ke += 0.5 * m * ((double)vel[4 * i] * vel[4 * i] + (double)vel[4 * i + 1] * vel[4 * i + 1] + (double)vel[4 * i + 2] * vel[4 * i + 2]);
        mobile += 1.0;
    }
    double lam = orc_csvr_lambda(ke / (double)ORC_ACCEL_CONV, g_csvr_kT, 3.0 * mobile - g_csvr_dof_removed,
                                 exp(-g_csvr_inv_tau * (double)dt), g_csvr_seed, g_csvr_step++);
    float l = (float)lam;
    for (int i = 0; i < n; ++i) { vel[4 * i] *= l; vel[4 * i + 1] *= l; vel[4 * i + 2] *= l; }
}

/*
 * Whole MD loop on the CPU (the CPU baseline and the C1 plumbing run): n_steps of velocity
 * Verlet with a Verlet list rebuilt when the largest displacement since the last build exceeds
 * skin/2.  Positions are NOT wrapped (min-image handles drift).  Returns the number of list
 * rebuilds, <0 on error.  energies_out (may be NULL): per step {E_lj, E_coul, E_bond, KE}.
 */
int orc_md_run(int n, float *xyzq, float *vel, const uint16_t *type, int T, const float *ljtab,
               const float *lo, const float *ext, int periodic, const orc_nb_params *p, float skin,
               const int32_t *excl_start, const int32_t *excl_idx,
               int npairs14, const int32_t *pairs14, float scale14_lj, float scale14_q,
               int nbonds, const int32_t *bonds, const float *bond_kr0,
               const float *ext_force /* n*3 or NULL */,
               float dt, int n_steps, int precision, double *energies_out, float *forces_out) {
    float rmax = p->rc_lj > p->rc_q ? p->rc_lj : p->rc_q;
    float r_list = rmax + skin;
    int64_t cap = 0;
    int64_t *nstart = (int64_t *)malloc(sizeof(int64_t) * ((size_t)n + 1));
    int32_t *nidx = NULL;
    float *xref = (float *)malloc(sizeof(float) * 4 * (size_t)n);
    float *f = (float *)calloc(4 * (size_t)n, sizeof(float));
    int rebuilds = 0, need = 1;
    double en[3];
    for (int step = 0; step <= n_steps; ++step) {
        if (need) {
            int64_t tot = orc_neighbors_cell(n, xyzq, lo, ext, periodic, r_list, excl_start, excl_idx,
                                             nstart, NULL, 0);
            if (tot > cap) { cap = tot + tot / 8 + 1024; free(nidx); nidx = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap); }
            if (orc_neighbors_cell(n, xyzq, lo, ext, periodic, r_list, excl_start, excl_idx, nstart, nidx, cap) < 0) {
                free(nstart); free(nidx); free(xref); free(f); return -1;
            }
            memcpy(xref, xyzq, sizeof(float) * 4 * (size_t)n);
            ++rebuilds; need = 0;
        }
        en[0] = en[1] = en[2] = 0;
        double e2[2] = {0, 0};
        orc_forces(n, xyzq, type, T, ljtab, ext, periodic, p, nstart, nidx, precision, f, NULL, NULL, e2);
        en[0] = e2[0]; en[1] = e2[1];
        if (npairs14) orc_pairs14(npairs14, pairs14, xyzq, type, T, ljtab, ext, periodic, scale14_lj, scale14_q,
                                  p->lj_on, p->coul_on, f, en, NULL, NULL);
        if (nbonds) orc_bonds(nbonds, bonds, bond_kr0, xyzq, ext, periodic, f, &en[2]);
        if (g_nv) orc_vsite_spread(f);
        if (ext_force)
            for (int i = 0; i < n; ++i)
                for (int a = 0; a < 3; ++a) f[4 * i + a] += ext_force[3 * i + a];
        if (step > 0) {
            orc_kick(n, vel, f, 0.5f * dt);
            if (g_nw || g_nhc) orc_rattle_velocities(xyzq, vel, ext, periodic);  /* on-step velocities of the constrained system */
        }
        if (energies_out) {
            energies_out[4 * step] = en[0]; energies_out[4 * step + 1] = en[1];
            energies_out[4 * step + 2] = en[2]; energies_out[4 * step + 3] = orc_kinetic(n, vel);
        }
        if (step == n_steps) break;
        orc_kick(n, vel, f, 0.5f * dt);
        float *xprev = NULL;
        if (g_nw || g_nhc) { xprev = (float *)malloc(sizeof(float) * 4 * (size_t)n); memcpy(xprev, xyzq, sizeof(float) * 4 * (size_t)n); }
        float worst = orc_drift(n, xyzq, vel, dt, xref);
        if (g_nw) orc_shake_waters(xprev, xyzq, vel, ext, periodic, dt);
        if (g_nhc) orc_shake_h(xprev, xyzq, vel, ext, periodic, dt);
        free(xprev);
        if (g_nv) orc_vsite_construct(xyzq, ext, periodic);
        if (g_lgv) orc_langevin_step(n, vel, dt);
        if (g_csvr) orc_csvr_step(n, vel, dt);
        if (worst > 0.25f * skin * skin) need = 1;
    }
    if (forces_out) memcpy(forces_out, f, sizeof(float) * 4 * (size_t)n);
    free(nstart); free(nidx); free(xref); free(f);
    return rebuilds;
}

/* ---- docking pose score ----------------------------------------------------------------- */

#define ORC_HYDROPHOBIC_CUTOFF 4.25f /* docking/legacy/mod.rs:70 */

/* rotate v by unit quaternion q = (w, x, y, z):  v' = v + 2 w (u x v) + 2 u x (u x v) */
static inline void quat_rot(const double *q, const double *v, double *o) {
    double ux = q[1], uy = q[2], uz = q[3], w = q[0];
    double cx = uy * v[2] - uz * v[1], cy = uz * v[0] - ux * v[2], cz = ux * v[1] - uy * v[0];
    double dx = uy * cz - uz * cy, dy = uz * cx - ux * cz, dz = ux * cy - uy * cx;
    o[0] = v[0] + 2.0 * (w * cx + dx);
    o[1] = v[1] + 2.0 * (w * cy + dy);
    o[2] = v[2] + 2.0 * (w * cz + dz);
}

/*
 * Pose score after docking/legacy/mod.rs:210-383 (calc_binding_energy) and :174-200
 * (BindingEnergy::new), rigid poses:
 *   ligand atom a at pose p:  x = anchor_p + R(q_p) * (lig_a - lig_anchor)      (legacy/mod.rs:149-158)
 *   vdw          = sum_{rec,lig} 4 eps ((sigma/r)^12 - (sigma/r)^6)             (:235-262, lj_V util.cu:74-90)
 *   hydrophobic  = sum over pairs with both flags set and r < 4.25 of -0.2 (1 - r/4.25)   (:305-321)
 *   electrostatic= | sum_{rec,lig} coulomb_force(rec -> lig) |   softening 1e-6, exact direct
 *                  sum (the reference approximates the same sum with Barnes-Hut, :332-375)
 *   coulomb_e    = sum q_r q_l / r    (raw extra output)
 *   score        = 1*vdw + (-1.2)*n_hbond + 1*hydrophobic + 10*electrostatic   (:174-200);
 *                  the H-bond finder lives in the external crate mol_defs -> n_hbond = 0 here.
 * rec: R*4 xyzq, rec_type R, rec_hphob R (0/1); lig: L*4 xyzq (reference conformation),
 * lig_type, lig_hphob; lig_anchor[3]; poses: P*7 doubles-as-float (ax,ay,az,qw,qx,qy,qz);
 * ljtab: Trec*Tlig (sigma, eps).  out: P*5 floats {score, vdw, hydrophobic, electrostatic, coulomb_e};
 * out_abs (may be NULL): P*3 floats {sum|vdw terms|, sum|coulomb pair forces|, sum|q q / r|}.
 * precision 32: fp32 arithmetic as the reference (:221-229); 64: fp64 truth.
 */
void orc_dock_score(int R, const float *rec, const uint16_t *rec_type, const uint8_t *rec_hphob,
                    int L, const float *lig, const uint16_t *lig_type, const uint8_t *lig_hphob,
                    const float *lig_anchor, int Tlig, const float *ljtab,
                    int P, const float *poses, int precision, float *out, float *out_abs) {
#pragma omp parallel for schedule(dynamic, 8)
    for (int p = 0; p < P; ++p) {
        const float *ps = poses + 7 * p;
        double q[4] = {ps[3], ps[4], ps[5], ps[6]};
        double qn = sqrt(((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3]);
        for (int a = 0; a < 4; ++a) q[a] /= qn;
        /* the pose transform runs in f64 and is rounded ONCE to f32, as the reference does
           (Pose{anchor_posit, orientation} are f64, lig_posits: &[Vec3F32], legacy/mod.rs:149-158,210-214);
           the CUDA kernel performs the identical f64 operations, so both sides score the same points */
        double *lp = (double *)malloc(sizeof(double) * 3 * (size_t)L);
        for (int a = 0; a < L; ++a) {
            double v[3] = {(double)lig[4 * a] - lig_anchor[0], (double)lig[4 * a + 1] - lig_anchor[1],
                           (double)lig[4 * a + 2] - lig_anchor[2]}, o[3];
            quat_rot(q, v, o);
            lp[3 * a] = (double)(float)(o[0] + ps[0]);
            lp[3 * a + 1] = (double)(float)(o[1] + ps[1]);
            lp[3 * a + 2] = (double)(float)(o[2] + ps[2]);
        }
        double vdw = 0, hyd = 0, ecoul = 0, fe[3] = {0, 0, 0}, a_vdw = 0, a_f = 0, a_e = 0;
        for (int r = 0; r < R; ++r) {
            for (int a = 0; a < L; ++a) {
                const float *lj = ljtab + 2 * ((size_t)rec_type[r] * Tlig + lig_type[a]);
                double qq = (double)rec[4 * r + 3] * (double)lig[4 * a + 3];
                if (precision == 32) {
                    float dx = (float)lp[3 * a] - rec[4 * r], dy = (float)lp[3 * a + 1] - rec[4 * r + 1],
                          dz = (float)lp[3 * a + 2] - rec[4 * r + 2];
                    float r2 = dx * dx + dy * dy + dz * dz, rr = sqrtf(r2);
                    float sr = lj[0] / rr, sr6 = sr * sr * sr * sr * sr * sr;
                    vdw += 4.f * lj[1] * (sr6 * sr6 - sr6);
                    a_vdw += fabsf(4.f * lj[1] * (sr6 * sr6 - sr6));
                    if (rec_hphob[r] && lig_hphob[a] && rr < ORC_HYDROPHOBIC_CUTOFF)
                        hyd += -0.2f * fmaxf(1.0f - rr / ORC_HYDROPHOBIC_CUTOFF, 0.f);
                    float mag = (float)qq / (r2 + ORC_SOFTENING_SQ) / rr;
                    fe[0] += dx * mag; fe[1] += dy * mag; fe[2] += dz * mag;
                    ecoul += (float)qq / rr;
                    a_f += fabsf(mag) * rr; a_e += fabsf((float)qq / rr);
                } else {
                    double dx = lp[3 * a] - rec[4 * r], dy = lp[3 * a + 1] - rec[4 * r + 1],
                           dz = lp[3 * a + 2] - rec[4 * r + 2];
                    double r2 = dx * dx + dy * dy + dz * dz, rr = sqrt(r2);
                    double sr = lj[0] / rr, sr6 = pow(sr, 6);
                    vdw += 4.0 * lj[1] * (sr6 * sr6 - sr6);
                    a_vdw += fabs(4.0 * lj[1] * (sr6 * sr6 - sr6));
                    if (rec_hphob[r] && lig_hphob[a] && rr < (double)ORC_HYDROPHOBIC_CUTOFF)
                        hyd += -0.2 * fmax(1.0 - rr / (double)ORC_HYDROPHOBIC_CUTOFF, 0.0);
                    double mag = qq / (r2 + (double)ORC_SOFTENING_SQ) / rr;
                    fe[0] += dx * mag; fe[1] += dy * mag; fe[2] += dz * mag;
                    ecoul += qq / rr;
                    a_f += fabs(mag) * rr; a_e += fabs(qq / rr);
                }
            }
        }
        double es = sqrt(fe[0] * fe[0] + fe[1] * fe[1] + fe[2] * fe[2]);
        out[5 * p] = (float)(1.0 * vdw + 0.0 + 1.0 * hyd + 10.0 * es);
        out[5 * p + 1] = (float)vdw;
        out[5 * p + 2] = (float)hyd;
        out[5 * p + 3] = (float)es;
        out[5 * p + 4] = (float)ecoul;
        if (out_abs) { /* sums of term magnitudes: the scale fp32 summation error is judged against */
            out_abs[3 * p] = (float)a_vdw; out_abs[3 * p + 1] = (float)a_f; out_abs[3 * p + 2] = (float)a_e;
        }
        free(lp);
    }
}
