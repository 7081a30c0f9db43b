/* TEST INFRASTRUCTURE ONLY -- plain-C restatement ("port") of the reference's 2B+3B hot path.
 *
 * Nothing in the product links, imports or calls this file; it exists so that tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg have an independent CPU statement of the algorithm to check the CUDA path against.
 * Parity of THIS file is pinned: tests/test_oracle.py checks it against the golden vectors of the reference's own
 * tests (platforms/reference/tests/TestReferenceMBPolTwoBodyForce.cpp:99-127, TestReferenceMBPolThreeBodyForce.cpp:
 * 97-138, TestReferenceThreeNeighborList.cpp:42-78) and against the reference's own code compiled unmodified
 * (oracle/_ref, oracle/Makefile).
 *
 * What is restated, and from where:
 *   tuple lists        ReferenceThreeNeighborList.cpp:14-33 (periodicDifference, compPairDistanceSquared),
 *                      :158-162 (d2 <= c2, j < i), :215-225 (chain triplets I > J > K)            -> port_pairs/_triplets
 *   PBC imaging        MBPolReferenceTwoBodyForce.cpp:66-109 (imageParticles, imageMolecules)      -> image_to()
 *   2B                 MBPolReferenceTwoBodyForce.cpp:110-296 + mbpol_2body_constants.cpp:10-111    -> port_dimer()
 *   3B                 MBPolReferenceThreeBodyForce.cpp:67-293                                     -> port_trimer()
 *   polynomials        poly-2b-v6x.cpp / poly-3b-v2x.cpp are Maple straight-line programs; restated here by their
 *                      mathematical definition  E = sum_m a[c_m] prod_{v in m} x_v  (monomial tables extracted by
 *                      tools/extract_pip.py, passed in by the caller) with the analytic gradient of each monomial.
 *                      This is deliberately NOT the evaluation scheme the CUDA kernels use (adjoint Horner).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PAD 255

typedef struct {
    int n_vars, n_mono;
    const double* coef;          /* [n_coef] */
    const uint16_t* mono_coef;   /* [n_mono] */
    const uint8_t* mono_vars;    /* [n_mono][4], padded with 255 */
} pip_table;

static pip_table T2, T3;

void port_set_tables(int order, int n_vars, int n_mono, const double* coef, const uint16_t* mono_coef,
                     const uint8_t* mono_vars) {
    pip_table* t = order == 2 ? &T2 : &T3;
    t->n_vars = n_vars; t->n_mono = n_mono; t->coef = coef; t->mono_coef = mono_coef; t->mono_vars = mono_vars;
}

/* E and dE/dx from the monomial definition */
double port_poly(int order, const double* x, double* g) {
    const pip_table* t = order == 2 ? &T2 : &T3;
    double e = 0.0;
    for (int i = 0; i < t->n_vars; ++i) g[i] = 0.0;
    for (int m = 0; m < t->n_mono; ++m) {
        const uint8_t* v = t->mono_vars + 4 * m;
        const double a = t->coef[t->mono_coef[m]];
        int deg = 0;
        double f[4];
        while (deg < 4 && v[deg] != PAD) { f[deg] = x[v[deg]]; ++deg; }
        double p = a;
        for (int s = 0; s < deg; ++s) p *= f[s];
        e += p;
        for (int s = 0; s < deg; ++s) {
            double q = a;
            for (int r = 0; r < deg; ++r) if (r != s) q *= f[r];
            g[v[s]] += q;
        }
    }
    return e;
}

/* ---- tuple lists ---------------------------------------------------------------------------------------------- */
static double periodic_difference(double v1, double v2, double period) {
    double diff = v1 - v2;
    double base = floor(diff / period + 0.5) * period;
    return diff - base;
}
static double pair_d2(const double* p1, const double* p2, const double* box, int pbc) {
    double dx, dy, dz;
    if (!pbc) { dx = p2[0] - p1[0]; dy = p2[1] - p1[1]; dz = p2[2] - p1[2]; }
    else {
        dx = periodic_difference(p2[0], p1[0], box[0]);
        dy = periodic_difference(p2[1], p1[1], box[1]);
        dz = periodic_difference(p2[2], p1[2], box[2]);
    }
    return dx * dx + dy * dy + dz * dz;
}

/* nearby[i] = { j < i : d2 <= c2 }, ascending j; returns CSR (malloc'ed) */
static void nearby_lists(const double* posO, int n, const double* box, int pbc, double cutoff, long** start_out,
                         int** idx_out) {
    const double c2 = cutoff * cutoff;
    long* start = (long*)malloc(sizeof(long) * (n + 1));
    long cap = 1024, m = 0;
    int* idx = (int*)malloc(sizeof(int) * cap);
    for (int i = 0; i < n; ++i) {
        start[i] = m;
        for (int j = 0; j < i; ++j) {
            if (pair_d2(posO + 3 * i, posO + 3 * j, box, pbc) > c2) continue;
            if (m == cap) { cap *= 2; idx = (int*)realloc(idx, sizeof(int) * cap); }
            idx[m++] = j;
        }
    }
    start[n] = m;
    *start_out = start; *idx_out = idx;
}

/* canonical (ascending I, J) pair list; returns the count, writes at most cap pairs.  O(n^2): small systems only */
long port_pairs(const double* posO, int n, const double* box, int pbc, double cutoff, int* out, long cap) {
    long* st; int* idx;
    nearby_lists(posO, n, box, pbc, cutoff, &st, &idx);
    long m = 0;
    for (int i = 0; i < n; ++i)
        for (long q = st[i]; q < st[i + 1]; ++q, ++m)
            if (m < cap) { out[2 * m] = i; out[2 * m + 1] = idx[q]; }
    free(st); free(idx);
    return m;
}

long port_triplets(const double* posO, int n, const double* box, int pbc, double cutoff, int* out, long cap) {
    long* st; int* idx;
    nearby_lists(posO, n, box, pbc, cutoff, &st, &idx);
    long m = 0;
    for (int i = 0; i < n; ++i)
        for (long q = st[i]; q < st[i + 1]; ++q) {
            const int j = idx[q];
            for (long r = st[j]; r < st[j + 1]; ++r, ++m)
                if (m < cap) { out[3 * m] = i; out[3 * m + 1] = j; out[3 * m + 2] = idx[r]; }
        }
    free(st); free(idx);
    return m;
}

/* ---- geometry helpers -------------------------------------------------------------------------------------------- */
typedef struct { double v[3]; } vec3;
static vec3 sub(vec3 a, vec3 b) { vec3 r = {{a.v[0] - b.v[0], a.v[1] - b.v[1], a.v[2] - b.v[2]}}; return r; }
static vec3 add(vec3 a, vec3 b) { vec3 r = {{a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2]}}; return r; }
static vec3 scl(vec3 a, double s) { vec3 r = {{a.v[0] * s, a.v[1] * s, a.v[2] * s}}; return r; }
static double dot(vec3 a, vec3 b) { return a.v[0] * b.v[0] + a.v[1] * b.v[1] + a.v[2] * b.v[2]; }
static vec3 cross(vec3 a, vec3 b) {
    vec3 r = {{a.v[1] * b.v[2] - a.v[2] * b.v[1], a.v[2] * b.v[0] - a.v[0] * b.v[2], a.v[0] * b.v[1] - a.v[1] * b.v[0]}};
    return r;
}
static void image_to(const double* box, const vec3* ref, vec3* p) {
    for (int i = 0; i < 3; ++i) {
        double distance = ref->v[i] - p->v[i];
        double factor = floor(distance / box[i] + 0.5);
        p->v[i] += box[i] * factor;
    }
}
static vec3 load(const double* pos, int particle, double scale) {
    vec3 r = {{pos[3 * particle] * scale, pos[3 * particle + 1] * scale, pos[3 * particle + 2] * scale}};
    return r;
}
static double cos_switch(double r, double ri, double rf, double* g) {
    if (r > rf) { *g = 0.0; return 0.0; }
    if (r > ri) {
        const double t1 = M_PI / (rf - ri);
        const double x = (r - ri) * t1;
        *g = -sin(x) * t1 / 2.0;
        return (1.0 + cos(x)) / 2.0;
    }
    *g = 0.0;
    return 1.0;
}

/* ---- 2B ---------------------------------------------------------------------------------------------------------- */
static const double k_HH_intra = -6.480884773303821e-01, k_OH_intra = 1.674518993682975e+00;
static const double k_HH_coul = 1.148231864355956e+00, k_OH_coul = 1.205989761123099e+00, k_OO_coul = 1.395357065790959e+00;
static const double k_XH_main = 7.347036852042255e-01, k_XO_main = 7.998249864422826e-01, k_XX_main = 7.960663960630585e-01;
static const double in_plane_gamma = -9.721486914088159e-02, out_of_plane_gamma = 9.859272078406150e-02;
static const double r2i = 4.5, r2f = 6.5;

/* variable table: kind (0 v_exp, 1 v_coul), k, r0, site a, site b; sites 0..5 atoms, 6..9 = Xa1,Xa2,Xb1,Xb2 */
typedef struct { int kind; double k, r0; int a, b; } var2;
static void vars2(var2* t) {
    int n = 0;
#define V(kind_, k_, r0_, a_, b_) t[n].kind = kind_, t[n].k = k_, t[n].r0 = r0_, t[n].a = a_, t[n].b = b_, ++n
    V(0, k_HH_intra, 1.0, 1, 2); V(0, k_HH_intra, 1.0, 4, 5);
    V(0, k_OH_intra, 1.0, 0, 1); V(0, k_OH_intra, 1.0, 0, 2); V(0, k_OH_intra, 1.0, 3, 4); V(0, k_OH_intra, 1.0, 3, 5);
    V(1, k_HH_coul, 4.0, 1, 4); V(1, k_HH_coul, 4.0, 1, 5); V(1, k_HH_coul, 4.0, 2, 4); V(1, k_HH_coul, 4.0, 2, 5);
    V(1, k_OH_coul, 4.0, 0, 4); V(1, k_OH_coul, 4.0, 0, 5); V(1, k_OH_coul, 4.0, 3, 1); V(1, k_OH_coul, 4.0, 3, 2);
    V(1, k_OO_coul, 4.0, 0, 3);
    V(0, k_XH_main, 4.0, 6, 4); V(0, k_XH_main, 4.0, 6, 5); V(0, k_XH_main, 4.0, 7, 4); V(0, k_XH_main, 4.0, 7, 5);
    V(0, k_XH_main, 4.0, 8, 1); V(0, k_XH_main, 4.0, 8, 2); V(0, k_XH_main, 4.0, 9, 1); V(0, k_XH_main, 4.0, 9, 2);
    V(0, k_XO_main, 4.0, 0, 8); V(0, k_XO_main, 4.0, 0, 9); V(0, k_XO_main, 4.0, 3, 6); V(0, k_XO_main, 4.0, 3, 7);
    V(0, k_XX_main, 4.0, 6, 8); V(0, k_XX_main, 4.0, 6, 9); V(0, k_XX_main, 4.0, 7, 8); V(0, k_XX_main, 4.0, 7, 9);
#undef V
}

/* energy (kJ/mol) of one dimer; f[6][3] receives the force increments (kJ/mol/nm) on Oa,Ha1,Ha2,Ob,Hb1,Hb2 */
double port_dimer(const double* pos, const int* ia, const int* ib, const double* box_nm, int pbc, double* f) {
    vec3 s[10], F[10];
    for (int k = 0; k < 3; ++k) { s[k] = load(pos, ia[k], 10.0); s[3 + k] = load(pos, ib[k], 10.0); }
    memset(F, 0, sizeof F);
    memset(f, 0, sizeof(double) * 18);
    if (pbc) {
        double box[3] = {box_nm[0] * 10.0, box_nm[1] * 10.0, box_nm[2] * 10.0};
        image_to(box, &s[0], &s[1]); image_to(box, &s[0], &s[2]); image_to(box, &s[0], &s[3]);
        image_to(box, &s[3], &s[4]); image_to(box, &s[3], &s[5]);
    }
    vec3 dOO = sub(s[0], s[3]);
    const double rOO = sqrt(dot(dOO, dOO));
    if (rOO > r2f || rOO < 2.0) return 0.0;
    vec3 oh1[2], oh2[2];
    for (int m = 0; m < 2; ++m) {                       /* monomer::setup */
        oh1[m] = sub(s[3 * m + 1], s[3 * m]); oh2[m] = sub(s[3 * m + 2], s[3 * m]);
        vec3 v = cross(oh1[m], oh2[m]);
        vec3 in_plane = add(s[3 * m], scl(add(oh1[m], oh2[m]), 0.5 * in_plane_gamma));
        vec3 oop = scl(v, out_of_plane_gamma);
        s[6 + 2 * m] = add(in_plane, oop); s[7 + 2 * m] = sub(in_plane, oop);
    }
    var2 t[31];
    vars2(t);
    double x[31], g[31];
    vec3 gv[31];
    for (int i = 0; i < 31; ++i) {
        vec3 d = sub(s[t[i].a], s[t[i].b]);
        const double r = sqrt(dot(d, d));
        const double e = exp(t[i].k * (t[i].r0 - r));
        if (t[i].kind == 0) { x[i] = e; gv[i] = scl(d, -t[i].k * e / r); }
        else { const double rinv = 1.0 / r, val = e * rinv; x[i] = val; gv[i] = scl(d, -(t[i].k + rinv) * val * rinv); }
    }
    const double E_poly = port_poly(2, x, g);
    for (int i = 0; i < 31; ++i) {
        vec3 d = scl(gv[i], g[i]);
        F[t[i].a] = add(F[t[i].a], d); F[t[i].b] = sub(F[t[i].b], d);
    }
    for (int m = 0; m < 2; ++m) {                       /* monomer::grads */
        vec3 g1 = F[6 + 2 * m], g2 = F[7 + 2 * m];
        vec3 gm = sub(g1, g2), t1 = cross(oh2[m], gm), t2 = cross(oh1[m], gm), gsum = add(g1, g2);
        vec3 in_plane = scl(gsum, 0.5 * in_plane_gamma);
        vec3 gh1 = add(in_plane, scl(t1, out_of_plane_gamma)), gh2 = sub(in_plane, scl(t2, out_of_plane_gamma));
        F[3 * m] = add(F[3 * m], sub(gsum, add(gh1, gh2)));
        F[3 * m + 1] = add(F[3 * m + 1], gh1); F[3 * m + 2] = add(F[3 * m + 2], gh2);
    }
    double gsw;
    const double sw = cos_switch(rOO, r2i, r2f, &gsw);
    const double cal2joule = 4.184;
    for (int a = 0; a < 6; ++a) for (int c = 0; c < 3; ++c) f[3 * a + c] = F[a].v[c] * sw * cal2joule * -10.;
    gsw *= E_poly / rOO;
    for (int c = 0; c < 3; ++c) {
        const double d = gsw * dOO.v[c];
        f[c] += d * cal2joule * -10.; f[9 + c] -= d * cal2joule * -10.;
    }
    return sw * E_poly * cal2joule;
}

/* ---- 3B ---------------------------------------------------------------------------------------------------------- */
static const double kHH_intra = -2.254303257872797e+00, kOH_intra = -2.564771901404151e-01;
static const double kHH = 4.247920544074333e-01, kOH = 8.128985941165371e-01, kOO = 3.749457984616480e-02;
static const double dHH_intra = 1.690594510379166e+00, dOH_intra = -2.999999868517452e+00;
static const double dHH = 3.499031358429095e+00, dOH = 4.854042963514281e+00, dOO_ = 4.816312044947604e-08;
static const double r3i = 0.0, r3f = 4.5;
static const int pair3[36][2] = {{1,2},{4,5},{7,8},{0,1},{0,2},{3,4},{3,5},{6,7},{6,8},
    {1,4},{1,5},{1,7},{1,8},{2,4},{2,5},{2,7},{2,8},{4,7},{4,8},{5,7},{5,8},
    {0,4},{0,5},{0,7},{0,8},{3,1},{3,2},{3,7},{3,8},{6,1},{6,2},{6,4},{6,5},{0,3},{0,6},{3,6}};

double port_trimer(const double* pos, const int* ia, const int* ib, const int* ic, const double* box, int pbc, double* f) {
    vec3 s[9], F[9];
    for (int k = 0; k < 3; ++k) { s[k] = load(pos, ia[k], 1.0); s[3 + k] = load(pos, ib[k], 1.0); s[6 + k] = load(pos, ic[k], 1.0); }
    memset(F, 0, sizeof F);
    memset(f, 0, sizeof(double) * 27);
    if (pbc) {
        image_to(box, &s[0], &s[1]); image_to(box, &s[0], &s[2]); image_to(box, &s[0], &s[3]);
        image_to(box, &s[3], &s[4]); image_to(box, &s[3], &s[5]); image_to(box, &s[0], &s[6]);
        image_to(box, &s[6], &s[7]); image_to(box, &s[6], &s[8]);
    }
    vec3 rab = scl(sub(s[0], s[3]), 10.0), rac = scl(sub(s[0], s[6]), 10.0), rbc = scl(sub(s[3], s[6]), 10.0);
    const double drab = sqrt(dot(rab, rab)), drac = sqrt(dot(rac, rac)), drbc = sqrt(dot(rbc, rbc));
    if (drab < 2 || drac < 2 || drbc < 2) return 0.0;
    double x[36], g[36], kk[36], d0[36];
    for (int i = 0; i < 36; ++i) {
        kk[i] = i < 3 ? kHH_intra : i < 9 ? kOH_intra : i < 21 ? kHH : i < 33 ? kOH : kOO;
        d0[i] = i < 3 ? dHH_intra : i < 9 ? dOH_intra : i < 21 ? dHH : i < 33 ? dOH : dOO_;
        vec3 dx = scl(sub(s[pair3[i][0]], s[pair3[i][1]]), 10.0);
        x[i] = exp(-kk[i] * (sqrt(dot(dx, dx)) - d0[i]));
    }
    double retval = port_poly(3, x, g);
    double gab, gac, gbc;
    const double sab = cos_switch(drab, r3i, r3f, &gab), sac = cos_switch(drac, r3i, r3f, &gac),
                 sbc = cos_switch(drbc, r3i, r3f, &gbc);
    const double sw = sab * sac + sab * sbc + sac * sbc;
    const double cal2joule = 4.184;
    for (int i = 0; i < 36; ++i) {
        vec3 dx = scl(sub(s[pair3[i][0]], s[pair3[i][1]]), 10.0);
        const double d = sqrt(dot(dx, dx));
        double gg = -kk[i] * (g[i] * sw) * exp(-kk[i] * (d - d0[i])) / d;
        gg *= cal2joule * 10. * -1;
        F[pair3[i][0]] = add(F[pair3[i][0]], scl(dx, gg));
        F[pair3[i][1]] = sub(F[pair3[i][1]], scl(dx, gg));
    }
    gab *= (sac + sbc) * retval / drab;
    gac *= (sab + sbc) * retval / drac;
    gbc *= (sab + sac) * retval / drbc;
    retval *= sw;
    for (int c = 0; c < 3; ++c) {
        F[0].v[c] += (gab * rab.v[c] + gac * rac.v[c]) * cal2joule * -10.0;
        F[3].v[c] += (gbc * rbc.v[c] - gab * rab.v[c]) * cal2joule * -10.0;
        F[6].v[c] -= (gac * rac.v[c] + gbc * rbc.v[c]) * cal2joule * -10.0;
    }
    for (int a = 0; a < 9; ++a) for (int c = 0; c < 3; ++c) f[3 * a + c] = F[a].v[c];
    return retval * cal2joule;
}

/* ---- drivers: serial loops over the lists, forces accumulated in list order like the Reference platform ----------- */
double port_eval_2b(const double* pos, double* forces, const int* idx, const int* pairs, long n_pairs, const double* box,
                    int pbc, double* e_each, double* f_each) {
    double energy = 0.0, f[18];
    for (long p = 0; p < n_pairs; ++p) {
        const int* ia = idx + 3 * pairs[2 * p];
        const int* ib = idx + 3 * pairs[2 * p + 1];
        const double e = port_dimer(pos, ia, ib, box, pbc, f);
        energy += e;
        if (e_each) e_each[p] = e;
        if (f_each) memcpy(f_each + 18 * p, f, sizeof f);
        if (forces)
            for (int k = 0; k < 3; ++k) for (int c = 0; c < 3; ++c) {
                forces[3 * ia[k] + c] += f[3 * k + c];
                forces[3 * ib[k] + c] += f[9 + 3 * k + c];
            }
    }
    return energy;
}

double port_eval_3b(const double* pos, double* forces, const int* idx, const int* trips, long n_trips, const double* box,
                    int pbc, double* e_each, double* f_each) {
    double energy = 0.0, f[27];
    for (long t = 0; t < n_trips; ++t) {
        const int* im[3] = {idx + 3 * trips[3 * t], idx + 3 * trips[3 * t + 1], idx + 3 * trips[3 * t + 2]};
        const double e = port_trimer(pos, im[0], im[1], im[2], box, pbc, f);
        energy += e;
        if (e_each) e_each[t] = e;
        if (f_each) memcpy(f_each + 27 * t, f, sizeof f);
        if (forces)
            for (int m = 0; m < 3; ++m) for (int k = 0; k < 3; ++k) for (int c = 0; c < 3; ++c)
                forces[3 * im[m][k] + c] += f[9 * m + 3 * k + c];
    }
    return energy;
}

/* ---- one-body term ---------------------------------------------------------------------------
 * Restates MBPolReferenceOneBodyForce::calculateOneBodyIxn (MBPolReferenceOneBodyForce.cpp:70-187) and
 * calculateForceAndEnergy (:191-210).  The parameter tables (mbpol_interaction_constants.h:24-424) are handed in by
 * refbind.py from mbpol_openmm_plugin_b200/data/onebody.npz. */
static struct {
    double c5z[245], reoh, b1, roh, alphaoh, deohA, phh1A, phh2, f5z, cm1_kcalmol;
    int i1[245], i2[245], i3[245];
} ps;

void port_set_onebody(const double* c5z, const int* i1, const int* i2, const int* i3, const double* scalars) {
    memcpy(ps.c5z, c5z, sizeof ps.c5z);
    memcpy(ps.i1, i1, sizeof ps.i1);
    memcpy(ps.i2, i2, sizeof ps.i2);
    memcpy(ps.i3, i3, sizeof ps.i3);
    ps.reoh = scalars[0]; ps.b1 = scalars[1]; ps.roh = scalars[2]; ps.alphaoh = scalars[3]; ps.deohA = scalars[4];
    ps.phh1A = scalars[5]; ps.phh2 = scalars[6]; ps.f5z = scalars[7]; ps.cm1_kcalmol = scalars[8];
}

static double port_monomer(vec3 O, vec3 H1, vec3 H2, double* f /* 3x3: O,H1,H2 increments */) {
    vec3 r1 = scl(sub(H1, O), 10.0), r2 = scl(sub(H2, O), 10.0), rh = scl(sub(H1, H2), 10.0);
    const double d1 = sqrt(dot(r1, r1)), d2 = sqrt(dot(r2, r2)), dh = sqrt(dot(rh, rh));
    const double costh = dot(r1, r2) / (d1 * d2);
    const double deoh = ps.f5z * ps.deohA, phh1 = ps.f5z * ps.phh1A * exp(ps.phh2);
    const double costhe = -.24780227221366464506;
    const double e1 = exp(-ps.alphaoh * (d1 - ps.roh)), e2 = exp(-ps.alphaoh * (d2 - ps.roh));
    const double Va = deoh * (e1 * (e1 - 2.0) + e2 * (e2 - 2.0));
    const double Vb = phh1 * exp(-ps.phh2 * dh);
    const double dVa1 = 2.0 * ps.alphaoh * deoh * e1 * (1.0 - e1) / d1;
    const double dVa2 = 2.0 * ps.alphaoh * deoh * e2 * (1.0 - e2) / d2;
    const double dVb = -ps.phh2 * Vb / dh;
    const double x[3] = {(d1 - ps.reoh) / ps.reoh, (d2 - ps.reoh) / ps.reoh, costh - costhe};
    double P[3][16];
    for (int a = 0; a < 3; ++a) {
        P[a][0] = 0.0;
        P[a][1] = 1.0;
        for (int j = 2; j < 16; ++j) P[a][j] = P[a][j - 1] * x[a];
    }
    const double efac = exp(-ps.b1 * ((d1 - ps.reoh) * (d1 - ps.reoh) + (d2 - ps.reoh) * (d2 - ps.reoh)));
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int j = 1; j < 245; ++j) {
        const int I = ps.i1[j], J = ps.i2[j], K = ps.i3[j];
        s0 += ps.c5z[j] * (P[0][I] * P[1][J] + P[0][J] * P[1][I]) * P[2][K];
        s1 += ps.c5z[j] * ((I - 1) * P[0][I - 1] * P[1][J] + (J - 1) * P[0][J - 1] * P[1][I]) * P[2][K];
        s2 += ps.c5z[j] * ((J - 1) * P[0][I] * P[1][J - 1] + (I - 1) * P[0][J] * P[1][I - 1]) * P[2][K];
        s3 += ps.c5z[j] * (P[0][I] * P[1][J] + P[0][J] * P[1][I]) * (K - 1) * P[2][K - 1];
    }
    double e = Va + Vb + (2 * ps.c5z[0] + efac * s0);
    e += 0.44739574026257;
    e *= ps.cm1_kcalmol;
    const double dr1 = (-2 * ps.b1 * efac * (d1 - ps.reoh) * s0 + efac * s1 / ps.reoh) / d1;
    const double dr2 = (-2 * ps.b1 * efac * (d2 - ps.reoh) * s0 + efac * s2 / ps.reoh) / d2;
    const double dct = efac * s3, unit = ps.cm1_kcalmol * 4.184 * 10;
    for (int c = 0; c < 3; ++c) {
        const double f1 = (dVa1 * r1.v[c] + dVb * rh.v[c] + dr1 * r1.v[c] +
                           dct * (r2.v[c] / (d1 * d2) - costh * r1.v[c] / (d1 * d1))) * unit;
        const double f2 = (dVa2 * r2.v[c] - dVb * rh.v[c] + dr2 * r2.v[c] +
                           dct * (r1.v[c] / (d1 * d2) - costh * r2.v[c] / (d2 * d2))) * unit;
        f[3 + c] = -f1;
        f[6 + c] = -f2;
        f[c] = f1 + f2;
    }
    return e * 4.184;
}

double port_eval_1b(const double* pos, double* forces, const int* idx, int n_mol, const double* box, int pbc) {
    double energy = 0.0, f[9];
    for (int m = 0; m < n_mol; ++m) {
        const int* im = idx + 3 * m;
        vec3 O = load(pos, im[0], 1.0), H1 = load(pos, im[1], 1.0), H2 = load(pos, im[2], 1.0);
        if (pbc) {
            image_to(box, &O, &H1);
            image_to(box, &O, &H2);
        }
        energy += port_monomer(O, H1, H2, f);
        if (forces)
            for (int k = 0; k < 3; ++k) for (int c = 0; c < 3; ++c) forces[3 * im[k] + c] += f[3 * k + c];
    }
    return energy;
}
