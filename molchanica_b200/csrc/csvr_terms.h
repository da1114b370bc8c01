// csvr_terms.h -- arithmetic of the canonical-sampling velocity-rescaling thermostat (Bussi, Donadio & Parrinello,
// J. Chem. Phys. 126, 014101 (2007); the reference's default NVT thermostat "CSVR", README.md:238,
// ui/panels/md.rs:296-306): one scalar per step,
//   K' = K + (1 - c) (Kbar (R1^2 + S) / Nf - K) + 2 R1 sqrt(K Kbar / Nf (1 - c) c),   c = exp(-dt / tau),
//   Kbar = Nf kT / 2,  R1 ~ N(0, 1),  S ~ chi^2(Nf - 1),   lambda = sqrt(K' / K),   v <- lambda v.
// Random numbers: Philox4x32-10 keyed by the seed, counter = (draw index, step) -- a pure function of (seed, step),
// so the CPU oracle draws the same lambda.  Evaluated by ONE device thread in fp64 (thermostat.cu) and by the host
// tests (tests/test_csvr_cpu.py) from this same header.
#pragma once
#include "langevin_terms.h"

struct CsvrRng {
    uint64_t seed, step;
    uint32_t draw;
};

MC_LGV_HD void mc_csvr_words(CsvrRng &g, uint32_t r[4]) {
    const uint32_t ctr[4] = {g.draw++, (uint32_t)g.step, (uint32_t)(g.step >> 32), 0xC5A1u};
    const uint32_t key[2] = {(uint32_t)g.seed, (uint32_t)(g.seed >> 32)};
    mc_philox4x32_10(ctr, key, r);
}

MC_LGV_HD double mc_csvr_u01(uint32_t hi, uint32_t lo) {  // (0, 1) from 53 random bits
    return ((double)(((uint64_t)hi << 21) ^ (uint64_t)(lo >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
}

// two independent standard normals and two uniforms per Philox call
MC_LGV_HD void mc_csvr_normal_pair(CsvrRng &g, double *n0, double *n1) {
    uint32_t r[4];
    mc_csvr_words(g, r);
    const double u0 = mc_csvr_u01(r[0], r[1]), u1 = mc_csvr_u01(r[2], r[3]);
    const double rad = sqrt(-2.0 * log(u0));
    *n0 = rad * cos(6.283185307179586 * u1);
    *n1 = rad * sin(6.283185307179586 * u1);
}

// Gamma(shape k >= 1, scale 1) by Marsaglia & Tsang (ACM TOMS 26, 363 (2000))
MC_LGV_HD double mc_csvr_gamma(CsvrRng &g, double k) {
    const double d = k - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * d);
    for (int trial = 0; trial < 1000; ++trial) {
        double x, unused;
        mc_csvr_normal_pair(g, &x, &unused);
        const double t = 1.0 + cc * x;
        if (t <= 0.0) continue;
        const double v = t * t * t;
        uint32_t r[4];
        mc_csvr_words(g, r);
        const double u = mc_csvr_u01(r[0], r[1]);
        if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) return d * v;
    }
    return d;  // not reached in practice (acceptance > 95 %)
}

// chi^2 with n >= 1 degrees of freedom
MC_LGV_HD double mc_csvr_chi2(CsvrRng &g, double n) {
    if (n >= 2.0) return 2.0 * mc_csvr_gamma(g, 0.5 * n);
    double a, b;                     // n = 1 (a system with two degrees of freedom): one squared normal
    mc_csvr_normal_pair(g, &a, &b);
    return a * a;
}

// The velocity scale factor of one step.  kinetic, kT in the same energy unit; nf = degrees of freedom.
MC_LGV_HD double mc_csvr_lambda(double kinetic, double kT, double nf, double c, uint64_t seed, uint64_t step) {
    if (!(kinetic > 0.0) || nf < 1.0) return 1.0;
    CsvrRng g = {seed, step, 0u};
    double r1, unused;
    mc_csvr_normal_pair(g, &r1, &unused);
    const double s = nf > 1.0 ? mc_csvr_chi2(g, nf - 1.0) : 0.0;
    const double kbar = 0.5 * nf * kT;
    double knew = kinetic + (1.0 - c) * (kbar * (r1 * r1 + s) / nf - kinetic) + 2.0 * r1 * sqrt(kinetic * kbar / nf * (1.0 - c) * c);
    if (knew < 0.0) knew = 0.0;
    return sqrt(knew / kinetic);
}
