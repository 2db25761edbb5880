/*
 * mc3d_oracle.c -- TEST INFRASTRUCTURE ONLY.  Scalar fp64 CPU restatement of the reference's photon walk.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.  It is
 * the checker the CUDA path is compared against, never a fallback: the product package (monte_carlompi_b200)
 * must not import or link it.
 *
 * Parity pinned: replay mode below is checked bit-for-bit (discrete columns) / to 1e-9 relative (float columns)
 * against records produced by the unmodified reference run under oracle/ref_shim.py with a seeded np.random
 * (fixtures in tests/golden/, generator oracle/make_golden.py).
 *
 * Follows (reference = /root/reference/monte_carloMPI/monte_carlo3D.py, "MC3D"):
 *   walk                 MC3D:1111-1490
 *   per-event draws      MC3D:885-921, 1010-1025 (populate_pdfs, sphere/HG branch)
 *   first-step draws     MC3D:1027-1044 (initial_pdfs)
 *   HG inverse CDF       MC3D:790-800 (Henyey_Greenstein2)
 *   wavelength draw      MC3D:1515-1520
 *   derived quantities   MC3D:1575-1588, 1612
 * numpy semantics kept on purpose: `x**2` on a numpy *scalar* is libm pow(x, 2.0), on an *array* it is x*x;
 * no fused multiply-add (compile with -ffp-contract=off).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TWO_PIE (2.0 * M_PI)

typedef struct {
    double theta0_rad, tau_tot, rho_snw, r_lambert, wvl0_um, sigma_um;
    int32_t k_first;
    uint32_t flags; /* 1 = Lambertian bottom, 2 = Lambertian surface */
    int32_t n_theta_bins;
    int32_t reserved;
} oracle_params;

typedef struct {
    double wvl_um, ssa_ice, ssa_imp, g, ext_cff_mss, p_ext_imp;
} oracle_row;

/* ---- source of uniforms: either the reference's recorded stream or Philox -------------------------------- */

/* event kinds: which part of the photon's random stream an event draws from (production mode; see the layout in
 * monte_carlompi_b200/csrc/mc3d_device.cuh and DESIGN.md "random-number layout") */
enum { EV_FIRST = 0, EV_WALK = 1, EV_REFLECT = 2 };

typedef struct {
    /* replay */
    const double *init3;   /* 3 first-step uniforms of this photon */
    const double *stream;  /* this photon's walk segment */
    int64_t n_stream, pos;
    int exhausted;
    /* philox */
    int use_philox;
    uint32_t key[2];
    uint64_t pid;
    int impurity_on;
    uint32_t blk;          /* next block of the walk stream (3 x groups started) */
    int slot;              /* next event slot of the current group; 4 = start a new group */
    uint32_t grp[12];      /* the twelve words of the current group */
    uint32_t first[4];     /* the TAG_FIRST block (wavelength + first event) */
    uint32_t key16;        /* coarse absorption variate of the last event drawn */
    uint32_t t16[2], t24[2];   /* absorption thresholds T40 = t16 << 24 | t24 of (ice, impurity / surface) */
} draw_src;

#define PHILOX_ROUNDS 7

static void philox4x32(int rounds, const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4])
{
    /* Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3" (SC'11), Philox-4x32, R rounds */
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < rounds; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* stream tags (counter word 1, low byte) */
enum { TAG_WALK = 0, TAG_SPECIES = 1, TAG_LAMBERT = 2, TAG_FIRST = 3, TAG_FINE = 4 };
#define RENORM_KEY 0xffc0u

static void philox_block(const draw_src *s, uint32_t c0, uint32_t c1, uint32_t w[4])
{
    uint32_t ctr[4] = {c0, c1, (uint32_t)s->pid, (uint32_t)(s->pid >> 32)};
    philox4x32(PHILOX_ROUNDS, ctr, s->key, w);
}

static inline double u32_to_unit(uint32_t w) { return ((double)w + 0.5) * (1.0 / 4294967296.0); }

/* T40 = ceil(ssa 2^40 - 1/2) clamped to [0, 2^40]: (K40 + 1/2) 2^-40 >= ssa  <=>  K40 >= T40 */
static void threshold40(double ssa, uint32_t *t16, uint32_t *t24)
{
    double t = ceil(ldexp(ssa, 40) - 0.5);
    if (!(t > 0.0)) t = 0.0;
    if (t >= 1099511627776.0) t = 1099511627776.0;
    uint64_t T = (uint64_t)t;
    *t16 = (uint32_t)(T >> 24);
    *t24 = (uint32_t)(T & 0xffffffu);
}

static double replay_next(draw_src *s)
{
    if (s->pos >= s->n_stream) { s->exhausted = 1; s->pos++; return 0.5; }
    return s->stream[s->pos++];
}

typedef struct { double r1, u_phi, u_tau, u_ssa, u_ext; } event_draws;

/* the 5 (3 on the first step) uniforms of event i, in the reference's order MC3D:915-921, 1014-1023, 1036-1038.
 * `species` selects whose albedo the 40-bit variate will be compared with (0 ice, 1 impurity / Lambertian surface):
 * its low 24 bits are only drawn when the top 16 do not decide the comparison. */
static void draw_event(draw_src *s, int64_t i, int kind, event_draws *d)
{
    if (!s->use_philox) {
        if (i == 1) {
            d->r1 = d->u_phi = 0.0;
            d->u_tau = s->init3[0]; d->u_ssa = s->init3[1]; d->u_ext = s->init3[2];
        } else {
            d->r1 = replay_next(s); d->u_phi = replay_next(s); d->u_tau = replay_next(s);
            d->u_ssa = replay_next(s); d->u_ext = replay_next(s);
        }
        return;
    }
    if (kind == EV_FIRST) {                       /* TAG_FIRST: w2 free path, top half of w3 = key16 */
        d->r1 = d->u_phi = 0.0;
        d->u_tau = u32_to_unit(s->first[2]);
        s->key16 = s->first[3] >> 16;
    } else if (kind == EV_REFLECT) {              /* TAG_LAMBERT sub-block 0 of event i: w1 azimuth, w2 free path, w3 key */
        uint32_t w[4];
        philox_block(s, (uint32_t)i, TAG_LAMBERT, w);
        d->r1 = 0.0;                              /* the HG deflection is not used by a Lambertian reflection */
        d->u_phi = u32_to_unit(w[1]);
        d->u_tau = u32_to_unit(w[2]);
        s->key16 = w[3] >> 16;
    } else {                                      /* the walk stream: groups of 4 events on 3 blocks */
        if (s->slot >= 4) {
            for (int b = 0; b < 3; ++b) philox_block(s, s->blk + b, TAG_WALK, s->grp + 4 * b);
            s->blk += 3;
            s->slot = 0;
        }
        const uint32_t *w = s->grp + 3 * s->slot;
        s->slot += 1;
        d->r1 = u32_to_unit(w[0]);
        d->u_phi = u32_to_unit(w[1]);
        d->u_tau = u32_to_unit(w[2]);
        s->key16 = ((w[0] & 0xFFu) << 8) | (w[2] & 0xFFu);
    }
    if (s->impurity_on) {
        uint32_t v[4];
        philox_block(s, (uint32_t)(i >> 2), TAG_SPECIES, v);
        d->u_ext = u32_to_unit(v[i & 3]);
    } else {
        d->u_ext = 1.0; /* P_ext_imp == 0: "u > 0" always holds, MC3D:1375-1379 */
    }
    d->u_ssa = -1.0;    /* completed by draw_albedo() once the species is known */
}

/* the single-scatter-albedo uniform of event i, (K40 + 1/2) 2^-40 with K40 = key16 << 24 | fine24 */
static double draw_albedo(draw_src *s, int64_t i, int species)
{
    uint64_t k40 = (uint64_t)s->key16 << 24;
    if (s->key16 == s->t16[species]) {
        uint32_t f[4];
        philox_block(s, (uint32_t)i, TAG_FINE, f);
        k40 |= f[0] >> 8;
    }
    return ((double)k40 + 0.5) * (1.0 / 1099511627776.0);
}

static double draw_reflectance(draw_src *s, int64_t i)
{
    if (!s->use_philox) return replay_next(s);
    uint32_t w[4];
    philox_block(s, (uint32_t)i, TAG_LAMBERT, w);
    return u32_to_unit(w[0]);
}

/* attempt j of the cosine-law rejection loop of event i, MC3D:1244-1250 */
static void draw_lambert_pair(draw_src *s, int64_t i, int64_t j, double *u_theta, double *r1)
{
    if (!s->use_philox) { *u_theta = replay_next(s); *r1 = replay_next(s); return; }
    uint32_t w[4];
    philox_block(s, (uint32_t)i, TAG_LAMBERT | (uint32_t)((1 + (j >> 1)) << 8), w);
    *u_theta = u32_to_unit(w[2 * (j & 1)]);
    *r1 = u32_to_unit(w[2 * (j & 1) + 1]);
}

/* MC3D:790-800; g is a numpy scalar (pow), the bracket is an array (x*x) */
static double henyey_greenstein2(double g, double r)
{
    if (g == 0) return 1 - 2 * r;
    double g2 = pow(g, 2.0);
    double q = (1 - g2) / (1 - g + 2 * g * r);
    return (1. / (2. * g)) * (1 + g2 - q * q);
}

typedef struct {
    int32_t condition;
    double wvn, theta_n, phi_n, path_length, snow_depth;
    int64_t n_scat;
} photon_out;

/* one photon, MC3D:1111-1490 */
static void walk(const oracle_params *P, draw_src *src, double wvl, double ssa_ice, double ssa_imp, double g,
                 double ext_cff_mss, double p_ext_imp, photon_out *out)
{
    const int lambert_bottom = (P->flags & 1u) != 0, lambert_surface = (P->flags & 2u) != 0;
    double mux_0 = sin(P->theta0_rad), muy_0 = 0, muz_0 = -cos(P->theta0_rad); /* MC3D:1121-1123 */
    double muz2_0 = pow(muz_0, 2.0);                                            /* MC3D:1124 */
    double mux_n = 0, muy_n = 0, muz_n = 0;
    double z_prev = 0, z = 0, path_length = 0;
    int bottom_reflection = 0, condition = 0;
    int64_t i = 0;
    const double ext_cff = ext_cff_mss * P->rho_snw; /* MC3D:1355-1356 */

    while (condition == 0) {
        i += 1;
        event_draws d;
        const int kind = (i == 1) ? EV_FIRST : ((lambert_surface || bottom_reflection) ? EV_REFLECT : EV_WALK);
        draw_event(src, i, kind, &d);
        if (src->exhausted) break;
        double dtau = -log(d.u_tau);                 /* MC3D:1014, 1036, 1227 */
        if (lambert_surface && i == 1) dtau = 0;     /* MC3D:1228-1229 */

        double costheta, sintheta;
        if (i == 1) {                                /* MC3D:1232-1237 */
            costheta = 1; sintheta = 0;
        } else if (lambert_surface || bottom_reflection) { /* MC3D:1238-1250 */
            mux_0 = 0.; muy_0 = 0.; muz_0 = 1.;
            int64_t j = 0;
            for (;;) {
                double u_theta, r1;
                draw_lambert_pair(src, i, j++, &u_theta, &r1);
                if (src->exhausted) break;
                double theta_rand = 0.0 + (M_PI / 2 - 0.0) * u_theta; /* legacy np.random.uniform(0, pi/2) */
                if (r1 < 2 * sin(theta_rand) * cos(theta_rand)) {
                    costheta = cos(theta_rand);
                    sintheta = sqrt(1 - pow(costheta, 2.0));
                    break;
                }
            }
            if (src->exhausted) break;
        } else {                                     /* MC3D:1252-1253 */
            costheta = henyey_greenstein2(g, d.r1);
            sintheta = sqrt(1 - pow(costheta, 2.0));
        }

        if (i > 1) {                                 /* MC3D:1255-1285 */
            double phi = d.u_phi * TWO_PIE;          /* MC3D:921 */
            double cosphi = cos(phi), sinphi = sin(phi);
            if (muz_0 == 1) {
                mux_n = sintheta * cosphi; muy_n = sintheta * sinphi; muz_n = costheta;
            } else if (muz_0 == -1) {
                mux_n = sintheta * cosphi; muy_n = -sintheta * sinphi; muz_n = -costheta;
            } else {
                double den = sqrt(1 - muz2_0);
                mux_n = (sintheta * (mux_0 * muz_0 * cosphi - muy_0 * sinphi)) / den + mux_0 * costheta;
                muy_n = (sintheta * (muy_0 * muz_0 * cosphi + mux_0 * sinphi)) / den + muy_0 * costheta;
                muz_n = -den * sintheta * cosphi + muz_0 * costheta;
            }
            if (bottom_reflection) bottom_reflection = 0;
        } else {                                     /* MC3D:1344-1347 */
            mux_n = mux_0; muy_n = muy_0; muz_n = muz_0;
        }

        z_prev = z;
        z = z_prev + dtau * muz_n;                   /* MC3D:1352 */
        if (i > 1) {                                 /* MC3D:1364-1369 */
            mux_0 = mux_n; muy_0 = muy_n; muz_0 = muz_n;
            muz2_0 = pow(muz_0, 2.0);
        }
        path_length += dtau / ext_cff;               /* MC3D:1372 */

        int ext_state;
        double ssa_event;
        if (d.u_ext > p_ext_imp) { ext_state = 1; ssa_event = ssa_ice; } /* MC3D:1375-1383 */
        else { ext_state = 2; ssa_event = ssa_imp; }
        if (lambert_surface) ssa_event = P->r_lambert; /* MC3D:1385-1387 */
        int species = (ext_state == 2) ? 1 : 0;
        if (src->use_philox) {
            if (lambert_surface) { species = 1; threshold40(P->r_lambert, &src->t16[1], &src->t24[1]); }
            d.u_ssa = draw_albedo(src, i, species);
        }
        int attention = 0;  /* production stream layout: the event "needs attention" */

        if (z > 0) {                                 /* MC3D:1390-1397 */
            condition = 1;
            path_length += -((z * dtau) / ((z - z_prev) * ext_cff));
        } else if (z < -P->tau_tot) {                /* MC3D:1399-1459 (both branches share the arithmetic) */
            attention = 1;
            path_length += -(((z + P->tau_tot) * dtau) / ((z - z_prev) * ext_cff));
            double dtau_correction = -(((z + P->tau_tot) / (z_prev - z)) * dtau);
            z = z - (muz_n * dtau_correction);
            int exit_cond = (i == 1) ? 3 : 2;
            if (lambert_bottom) {
                double reflectance_rand = draw_reflectance(src, i);
                if (src->exhausted) break;
                if (reflectance_rand <= P->r_lambert) bottom_reflection = 1;
                else condition = exit_cond;
            } else {
                condition = exit_cond;
            }
        } else if (d.u_ssa >= ssa_event) {           /* MC3D:1461-1466 */
            condition = (ext_state == 1) ? 4 : 5;
        }
        if (src->use_philox && kind == EV_WALK) {
            /* a photon that survives an event needing attention (boundary hit, possible absorption, or due for the
             * GPU path's renormalisation: key16 >= min(t16, 0xffc0)) continues with the next group of its stream */
            uint32_t thot = src->t16[species] < RENORM_KEY ? src->t16[species] : RENORM_KEY;
            if (attention || src->key16 >= thot) src->slot = 4;
        }
    }

    out->condition = condition;
    out->wvn = 1. / wvl;                             /* MC3D:1468 */
    out->theta_n = acos(muz_0);                      /* MC3D:1469 */
    if (i == 1) out->phi_n = 0.;                     /* MC3D:1472-1485 */
    else if (mux_0 > 0 && muy_0 > 0) out->phi_n = atan(muy_0 / mux_0);
    else if (mux_0 < 0 && muy_0 > 0) out->phi_n = atan(muy_0 / mux_0) + M_PI;
    else if (mux_0 < 0 && muy_0 < 0) out->phi_n = atan(muy_0 / mux_0) + M_PI;
    else if (mux_0 > 0 && muy_0 < 0) out->phi_n = atan(muy_0 / mux_0) + TWO_PIE;
    else out->phi_n = NAN;                           /* the reference raises UnboundLocalError here */
    out->n_scat = i - 1;                             /* MC3D:1487 */
    out->path_length = path_length;
    out->snow_depth = P->tau_tot / (ext_cff_mss * P->rho_snw); /* MC3D:1612 */
}

/* ---- replay mode ------------------------------------------------------------------------------------------ */

int oracle_replay(const oracle_params *P, int64_t n_photon, const double *wvl, const double *ssa_ice,
                  const double *ssa_imp, const double *g, const double *ext_cff_mss, const double *p_ext_imp,
                  const double *init_draws, const int64_t *offsets, const double *stream, int32_t *condition,
                  double *wvn, double *theta_n, double *phi_n, int64_t *n_scat, double *path_length,
                  double *snow_depth, int64_t *consumed)
{
    int64_t mismatches = 0;
    for (int64_t p = 0; p < n_photon; ++p) {
        draw_src src;
        memset(&src, 0, sizeof src);
        src.init3 = init_draws + 3 * p;
        src.stream = stream + offsets[p];
        src.n_stream = offsets[p + 1] - offsets[p];
        photon_out o;
        walk(P, &src, wvl[p], ssa_ice[p], ssa_imp[p], g[p], ext_cff_mss[p], p_ext_imp[p], &o);
        condition[p] = o.condition; wvn[p] = o.wvn; theta_n[p] = o.theta_n; phi_n[p] = o.phi_n;
        n_scat[p] = o.n_scat; path_length[p] = o.path_length; snow_depth[p] = o.snow_depth;
        consumed[p] = src.pos;
        if (src.pos != src.n_stream) mismatches++;
    }
    return (int)(mismatches > 2147483647 ? 2147483647 : mismatches);
}

/* ---- production-mode restatement (same Philox draws as the CUDA kernel, reference arithmetic in fp64) ---- */

/* np.histogram(x, bins=n, range=(0, pi/2)) bin of one value, numpy/lib/_histograms_impl.py (uniform-bin path) */
static int histogram_bin(double x, int n_bins, const double *edges)
{
    double first = edges[0], last = edges[n_bins];
    if (!(x >= first && x <= last)) return -1;
    double f = ((x - first) / (last - first)) * n_bins;
    int idx = (int)f;
    if (idx == n_bins) idx -= 1;
    if (x < edges[idx]) idx -= 1;
    else if (x >= edges[idx + 1] && idx != n_bins - 1) idx += 1;
    return idx;
}

typedef struct {
    const oracle_params *P;
    const oracle_row *table;
    int n_rows;
    uint64_t seed, begin, n;
    int fp32_angles; /* bin float64((float)theta) like the device does */
    int32_t *condition; int16_t *wvl_row; double *theta_n, *phi_n, *path_length; int64_t *n_scat;
    uint64_t *tally; /* [n_rows][8 + n_theta_bins], private per thread */
    const double *edges;
    uint64_t n_events;
    int tid, n_threads;
} philox_job;

static int wavelength_row(const oracle_params *P, draw_src *src, int n_rows)
{
    /* MC3D:1515-1520: wvls = np.around(np.random.normal(wvl0, scale), 2); Box-Muller on Philox uniforms */
    uint32_t *w = src->first;
    philox_block(src, 0, TAG_FIRST, w);
    double u1 = u32_to_unit(w[0]), u2 = u32_to_unit(w[1]);
    double zn = sqrt(-2.0 * log(u1)) * cos(TWO_PIE * u2);
    double k = rint((P->wvl0_um + P->sigma_um * zn) * 100.0);
    int64_t row = (int64_t)k - P->k_first;
    if (row < 0) row = 0;
    if (row > n_rows - 1) row = n_rows - 1;
    return (int)row;
}

static void *philox_worker(void *arg)
{
    philox_job *J = (philox_job *)arg;
    const int stride = 8 + J->P->n_theta_bins;
    int impurity_on = 0;
    for (int r = 0; r < J->n_rows; ++r) if (J->table[r].p_ext_imp > 0) impurity_on = 1;
    /* contiguous chunk per thread (same boundaries as np.array_split) */
    uint64_t q = J->n / J->n_threads, rem = J->n % J->n_threads;
    uint64_t lo = J->tid * q + ((uint64_t)J->tid < rem ? J->tid : rem);
    uint64_t hi = lo + q + ((uint64_t)J->tid < rem ? 1 : 0);
    for (uint64_t idx = lo; idx < hi; ++idx) {
        draw_src src;
        memset(&src, 0, sizeof src);
        src.use_philox = 1;
        src.key[0] = (uint32_t)J->seed; src.key[1] = (uint32_t)(J->seed >> 32);
        src.pid = J->begin + idx;
        src.impurity_on = impurity_on;
        src.slot = 4;
        int row = wavelength_row(J->P, &src, J->n_rows);
        const oracle_row *R = &J->table[row];
        threshold40(R->ssa_ice, &src.t16[0], &src.t24[0]);
        threshold40(R->ssa_imp, &src.t16[1], &src.t24[1]);
        photon_out o;
        walk(J->P, &src, R->wvl_um, R->ssa_ice, R->ssa_imp, R->g, R->ext_cff_mss, R->p_ext_imp, &o);
        J->n_events += (uint64_t)(o.n_scat + 1);
        if (J->condition) J->condition[idx] = o.condition;
        if (J->wvl_row) J->wvl_row[idx] = (int16_t)row;
        if (J->theta_n) J->theta_n[idx] = o.theta_n;
        if (J->phi_n) J->phi_n[idx] = o.phi_n;
        if (J->path_length) J->path_length[idx] = o.path_length;
        if (J->n_scat) J->n_scat[idx] = o.n_scat;
        if (J->tally) {
            uint64_t *t = J->tally + (size_t)row * stride;
            t[0] += 1;
            t[o.condition] += 1;
            if (o.condition == 1 && J->P->n_theta_bins > 0) {
                double th = J->fp32_angles ? (double)(float)o.theta_n : o.theta_n;
                int b = histogram_bin(th, J->P->n_theta_bins, J->edges);
                if (b >= 0) t[8 + b] += 1;
            }
        }
    }
    return NULL;
}

/* Walk photon ids [begin, begin + n).  Any output pointer may be NULL.  tally: uint64[n_rows*(8+n_theta_bins)],
 * zeroed here.  edges: double[n_theta_bins + 1] = np.linspace(0, pi/2, n_theta_bins + 1).  Returns 0. */
int oracle_philox(const oracle_params *P, const oracle_row *table, int n_rows, uint64_t seed, uint64_t begin,
                  uint64_t n, int n_threads, int fp32_angles, const double *edges, int32_t *condition,
                  int16_t *wvl_row, double *theta_n, double *phi_n, int64_t *n_scat, double *path_length,
                  uint64_t *tally, uint64_t *n_events)
{
    if (n_threads < 1) n_threads = 1;
    const size_t tally_len = (size_t)n_rows * (8 + P->n_theta_bins);
    philox_job *jobs = (philox_job *)calloc(n_threads, sizeof *jobs);
    pthread_t *th = (pthread_t *)calloc(n_threads, sizeof *th);
    for (int t = 0; t < n_threads; ++t) {
        philox_job *J = &jobs[t];
        J->P = P; J->table = table; J->n_rows = n_rows; J->seed = seed; J->begin = begin; J->n = n;
        J->fp32_angles = fp32_angles; J->condition = condition; J->wvl_row = wvl_row; J->theta_n = theta_n;
        J->phi_n = phi_n; J->n_scat = n_scat; J->path_length = path_length; J->edges = edges;
        J->tally = tally ? (uint64_t *)calloc(tally_len, sizeof(uint64_t)) : NULL;
        J->tid = t; J->n_threads = n_threads;
        if (n_threads > 1) pthread_create(&th[t], NULL, philox_worker, J);
        else philox_worker(J);
    }
    uint64_t ev = 0;
    if (tally) memset(tally, 0, tally_len * sizeof(uint64_t));
    for (int t = 0; t < n_threads; ++t) {
        if (n_threads > 1) pthread_join(th[t], NULL);
        ev += jobs[t].n_events;
        if (tally) {
            for (size_t k = 0; k < tally_len; ++k) tally[k] += jobs[t].tally[k];
            free(jobs[t].tally);
        }
    }
    if (n_events) *n_events = ev;
    free(jobs); free(th);
    return 0;
}

/* exposed for unit tests */
void oracle_philox4x32(int rounds, const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox4x32(rounds, ctr, key, out); }
double oracle_henyey_greenstein2(double g, double r) { return henyey_greenstein2(g, r); }
int oracle_histogram_bin(double x, int n_bins, const double *edges) { return histogram_bin(x, n_bins, edges); }
