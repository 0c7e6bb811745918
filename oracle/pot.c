/*
 * oracle/pot.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE. See pot.h for provenance and the
 * "parity unpinned" statement (libpot v0.1.0 is absent from /root/reference).
 */
#include "pot.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * Table construction: restates LAMMPS PairEAM::array2spline, the algorithm libpot's
 * eam::interpolateFile() follows (call site: reference src/simulation.cpp:131).
 * Row layout: [0..2] derivative coefficients (already divided by dx), [3..6] value coefficients.
 * ---------------------------------------------------------------------------------------------- */
static void table_build(pot_table *t, int n, double dx, const double *data /* n values, 0-based */) {
    t->n = n;
    t->dx = dx;
    t->inv_dx = 1.0 / dx;
    t->values = (double *)calloc((size_t)n + 1, sizeof(double));
    t->spline = (double *)calloc(((size_t)n + 1) * 7, sizeof(double));
    for (int i = 0; i < n; i++) t->values[i + 1] = data[i];
#define S(m, k) t->spline[(size_t)(m) * 7 + (k)]
    for (int m = 1; m <= n; m++) S(m, 6) = t->values[m];
    S(1, 5) = S(2, 6) - S(1, 6);
    S(2, 5) = 0.5 * (S(3, 6) - S(1, 6));
    S(n - 1, 5) = 0.5 * (S(n, 6) - S(n - 2, 6));
    S(n, 5) = S(n, 6) - S(n - 1, 6);
    for (int m = 3; m <= n - 2; m++)
        S(m, 5) = ((S(m - 2, 6) - S(m + 2, 6)) + 8.0 * (S(m + 1, 6) - S(m - 1, 6))) / 12.0;
    for (int m = 1; m <= n - 1; m++) {
        S(m, 4) = 3.0 * (S(m + 1, 6) - S(m, 6)) - 2.0 * S(m, 5) - S(m + 1, 5);
        S(m, 3) = S(m, 5) + S(m + 1, 5) - 2.0 * (S(m + 1, 6) - S(m, 6));
    }
    S(n, 4) = 0.0;
    S(n, 3) = 0.0;
    for (int m = 1; m <= n; m++) {
        S(m, 2) = S(m, 5) / dx;
        S(m, 1) = 2.0 * S(m, 4) / dx;
        S(m, 0) = 3.0 * S(m, 3) / dx;
    }
#undef S
}

static void table_free(pot_table *t) {
    free(t->values);
    free(t->spline);
    t->values = NULL;
    t->spline = NULL;
}

/* libpot InterpolationObject::findSpline restated: p = x*invDx + 1; m = clamp(int(p), 1, n-1);
 * p = min(p - m, 1). Returns the row pointer and writes the fractional coordinate. */
static inline const double *table_find(const pot_table *t, double x, double *p_out) {
    double p = x * t->inv_dx + 1.0;
    int m = (int)p;
    if (m > t->n - 1) m = t->n - 1;
    if (m < 1) m = 1;
    p -= m;
    if (p > 1.0) p = 1.0;
    *p_out = p;
    return t->spline + (size_t)m * 7;
}

int pot_index_of_key(const pot_eam *pt, int key) {
    for (int i = 0; i < pt->n_ele; i++)
        if (pt->key[i] == key) return i;
    return -1;
}

double pot_charge_density(const pot_eam *pt, int key, double dist2) {
    const pot_table *t = &pt->elec[pot_index_of_key(pt, key)];
    const double r = sqrt(dist2);
    double p;
    const double *s = table_find(t, r, &p);
    return ((s[3] * p + s[4]) * p + s[5]) * p + s[6];
}

double pot_d_embed_energy(const pot_eam *pt, int key, double rho) {
    const pot_table *t = &pt->embed[pot_index_of_key(pt, key)];
    double p;
    const double *s = table_find(t, rho, &p);
    return (s[0] * p + s[1]) * p + s[2];
}

double pot_to_force(const pot_eam *pt, int key_from, int key_to, double dist2, double df_from, double df_to) {
    const int i = pot_index_of_key(pt, key_from), j = pot_index_of_key(pt, key_to);
    const double r = sqrt(dist2);
    double p;
    const double *s = table_find(&pt->phi[i][j], r, &p);
    const double z2 = ((s[3] * p + s[4]) * p + s[5]) * p + s[6];
    const double z2p = (s[0] * p + s[1]) * p + s[2];
    s = table_find(&pt->elec[i], r, &p);
    const double rho_p_from = (s[0] * p + s[1]) * p + s[2];
    s = table_find(&pt->elec[j], r, &p);
    const double rho_p_to = (s[0] * p + s[1]) * p + s[2];
    const double recip = 1.0 / r;
    const double phi = z2 * recip;
    const double phip = z2p * recip - phi * recip;
    const double psip = phip + (rho_p_from * df_to + rho_p_to * df_from);
    return -psip * recip;
}

double pot_embed_energy(const pot_eam *pt, int key, double rho) {
    const pot_table *t = &pt->embed[pot_index_of_key(pt, key)];
    double p;
    const double *s = table_find(t, rho, &p);
    return ((s[3] * p + s[4]) * p + s[5]) * p + s[6];
}

double pot_pair_energy(const pot_eam *pt, int key_from, int key_to, double dist2) {
    const int i = pot_index_of_key(pt, key_from), j = pot_index_of_key(pt, key_to);
    const double r = sqrt(dist2);
    double p;
    const double *s = table_find(&pt->phi[i][j], r, &p);
    const double z2 = ((s[3] * p + s[4]) * p + s[5]) * p + s[6];
    return z2 / r;
}

void pot_free(pot_eam *pt) {
    if (!pt) return;
    for (int i = 0; i < pt->n_ele; i++) {
        table_free(&pt->embed[i]);
        table_free(&pt->elec[i]);
        for (int j = 0; j <= i; j++) table_free(&pt->phi[i][j]);
    }
    free(pt);
}

/* ------------------------------------------------------------------------------------------------
 * setfl reader ("eam/alloy" format; the reference reads it through libpot's SetflParser:
 * reference src/simulation.cpp:105-117).
 * ---------------------------------------------------------------------------------------------- */
static int read_doubles(FILE *f, double *dst, int n) {
    for (int i = 0; i < n; i++)
        if (fscanf(f, "%lf", &dst[i]) != 1) return -1;
    return 0;
}

pot_eam *pot_read_setfl(const char *path) {
    FILE *f = fopen(path, "r");
    if (!f) return NULL;
    char line[4096];
    for (int i = 0; i < 3; i++)
        if (!fgets(line, sizeof line, f)) { fclose(f); return NULL; }
    pot_eam *pt = (pot_eam *)calloc(1, sizeof(pot_eam));
    if (fscanf(f, "%d", &pt->n_ele) != 1 || pt->n_ele < 1 || pt->n_ele > POT_MAX_ELE) goto fail;
    for (int i = 0; i < pt->n_ele; i++)
        if (fscanf(f, "%4095s", line) != 1) goto fail; /* element names */
    if (fscanf(f, "%d %lf %d %lf %lf", &pt->n_rho, &pt->d_rho, &pt->n_r, &pt->d_r, &pt->cutoff) != 5) goto fail;
    if (pt->n_rho < 5 || pt->n_r < 5) goto fail;
    {
        double *buf = (double *)malloc(sizeof(double) * (size_t)(pt->n_rho > pt->n_r ? pt->n_rho : pt->n_r));
        for (int i = 0; i < pt->n_ele; i++) {
            char lat_type[64];
            if (fscanf(f, "%d %lf %lf %63s", &pt->key[i], &pt->mass[i], &pt->lat_const[i], lat_type) != 4) { free(buf); goto fail; }
            if (read_doubles(f, buf, pt->n_rho)) { free(buf); goto fail; }
            table_build(&pt->embed[i], pt->n_rho, pt->d_rho, buf);
            if (read_doubles(f, buf, pt->n_r)) { free(buf); goto fail; }
            table_build(&pt->elec[i], pt->n_r, pt->d_r, buf);
        }
        for (int i = 0; i < pt->n_ele; i++)
            for (int j = 0; j <= i; j++) {
                if (read_doubles(f, buf, pt->n_r)) { free(buf); goto fail; }
                table_build(&pt->phi[i][j], pt->n_r, pt->d_r, buf);
                if (j != i) pt->phi[j][i] = pt->phi[i][j]; /* alias */
            }
        free(buf);
    }
    fclose(f);
    return pt;
fail:
    fclose(f);
    free(pt);
    return NULL;
}

/* ------------------------------------------------------------------------------------------------
 * Synthetic Fe-Cu-Ni setfl (BASELINE.json configs[0]: "FeCuNi.eam.alloy (or synthetic same-format
 * table)"). Analytic forms after Zhou, Johnson & Wadley (PRB 69, 144113): generalized-exponential pair
 * and density terms, one smooth embedding branch; parameters are approximate, the file is synthetic.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int z;
    double mass, a0;
    double re, fe, rhos, alpha, beta, A, B, kappa, lambda, Fe, eta;
} zhou_par;

static const zhou_par ZP[3] = {
    /* Fe */ {26, 55.845, 2.85532, 2.481987, 1.885957, 20.041463, 9.818270, 5.236411, 0.392811, 0.646243, 0.170306, 0.340613, -2.539945, 0.391750},
    /* Cu */ {29, 63.546, 3.615, 2.556162, 1.554485, 21.175395, 8.127620, 4.334731, 0.396620, 0.548085, 0.308782, 0.756515, -2.176490, 0.763905},
    /* Ni */ {28, 58.6934, 3.52, 2.488746, 2.007018, 27.930410, 8.383453, 4.471175, 0.429046, 0.633531, 0.443599, 0.820658, -2.700493, 0.469000},
};

static double sw(double r, double rc) { /* quintic switch 1 -> 0 on [rs, rc] */
    const double rs = rc - 0.6;
    if (r <= rs) return 1.0;
    if (r >= rc) return 0.0;
    const double t = (rc - r) / (rc - rs);
    return t * t * t * (10.0 - 15.0 * t + 6.0 * t * t);
}
static double z_rho(const zhou_par *q, double r, double rc) {
    const double x = r / q->re;
    return q->fe * exp(-q->beta * (x - 1.0)) / (1.0 + pow(x - q->lambda, 20.0)) * sw(r, rc);
}
static double z_phi(const zhou_par *q, double r, double rc) {
    const double x = r / q->re;
    return (q->A * exp(-q->alpha * (x - 1.0)) / (1.0 + pow(x - q->kappa, 20.0)) -
            q->B * exp(-q->beta * (x - 1.0)) / (1.0 + pow(x - q->lambda, 20.0))) * sw(r, rc);
}
static double z_F(const zhou_par *q, double rho) {
    if (rho <= 0.0) return 0.0;
    const double x = rho / q->rhos;
    return q->Fe * (1.0 - q->eta * log(x)) * pow(x, q->eta);
}

int pot_write_synthetic_setfl(const char *path, int n_rho, double d_rho, int n_r, double d_r, double cutoff) {
    FILE *f = fopen(path, "w");
    if (!f) return -1;
    fprintf(f, "Synthetic Fe-Cu-Ni EAM table in setfl (eam/alloy) format -- NOT a fitted potential.\n");
    fprintf(f, "Generated by misa-b200 oracle/pot.c (Zhou-Johnson-Wadley style analytic forms).\n");
    fprintf(f, "Stand-in for FeCuNi.eam.alloy which is absent from the reference tree.\n");
    fprintf(f, "3 Fe Cu Ni\n");
    fprintf(f, "%d %.17g %d %.17g %.17g\n", n_rho, d_rho, n_r, d_r, cutoff);
    for (int e = 0; e < 3; e++) {
        const zhou_par *q = &ZP[e];
        fprintf(f, "%d %.17g %.17g %s\n", q->z, q->mass, q->a0, e == 0 ? "bcc" : "fcc");
        for (int i = 0; i < n_rho; i++) fprintf(f, "%.17g%c", z_F(q, i * d_rho), (i % 5 == 4 || i == n_rho - 1) ? '\n' : ' ');
        for (int i = 0; i < n_r; i++) fprintf(f, "%.17g%c", z_rho(q, i * d_r, cutoff), (i % 5 == 4 || i == n_r - 1) ? '\n' : ' ');
    }
    for (int a = 0; a < 3; a++)
        for (int b = 0; b <= a; b++)
            for (int i = 0; i < n_r; i++) {
                const double r = i * d_r;
                double v;
                if (a == b) {
                    v = z_phi(&ZP[a], r, cutoff);
                } else {
                    const double ra = z_rho(&ZP[a], r, cutoff), rb = z_rho(&ZP[b], r, cutoff);
                    v = (ra > 0.0 && rb > 0.0)
                            ? 0.5 * (rb / ra * z_phi(&ZP[a], r, cutoff) + ra / rb * z_phi(&ZP[b], r, cutoff))
                            : 0.0;
                }
                fprintf(f, "%.17g%c", r * v, (i % 5 == 4 || i == n_r - 1) ? '\n' : ' ');
            }
    fclose(f);
    return 0;
}
