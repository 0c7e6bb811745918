/*
 * oracle/pot.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the EAM potential library the reference links against:
 *   git.hpcer.dev/HPCer/CrystalMD/potential  v0.1.0   (reference pkg.yaml:16, CMake target pot::pot,
 *   reference src/CMakeLists.txt:83).  Its sources are NOT under /root/reference (vendor/ is git-ignored,
 *   reference .gitignore:7-8), so the arithmetic here restates the published algorithm that library
 *   follows (LAMMPS pair_eam: file2array/array2spline 7-coefficient cubic rows, tables of r*phi), anchored
 *   on the reference's own call sites:
 *     - eam::chargeDensity(key, r^2)                 reference src/atom.cpp:181,183,218,219,235,237,256,275
 *     - eam::dEmbedEnergy(key, rho)                  reference src/atom.cpp:281,303
 *     - eam::toForce(key_i, key_j, r^2, df_i, df_j)  reference src/atom.cpp:341,383,409,436,461
 *     - SetflParser / eam::interpolateFile()         reference src/simulation.cpp:105-131
 *   What those call sites fix: the distance argument is SQUARED, species keys are atomic numbers
 *   (26/29/28, reference src/types/atom_types.h:61-73), toForce returns the scalar that multiplies the
 *   displacement vector x_i - x_j (reference src/atom.cpp:341-351), dEmbedEnergy returns F'(rho).
 *
 *   PARITY UNPINNED at this boundary: no reference test evaluates these functions.
 *   The GPU path consumes the coefficient tables built HERE-equivalent on the host, so CUDA-vs-oracle parity
 *   does not depend on the spline construction; only the evaluation order is compared.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use this.
 */
#ifndef MISA_ORACLE_POT_H
#define MISA_ORACLE_POT_H

#ifdef __cplusplus
extern "C" {
#endif

#define POT_MAX_ELE 3

/* One tabulated function on a uniform grid starting at 0: values[1..n] (1-based like LAMMPS) and the
 * 7-coefficient rows spline[m][0..6], m = 1..n; row m covers x in [(m-1)*dx, m*dx). */
typedef struct pot_table {
    int n;
    double dx;
    double inv_dx;
    double *values; /* n+1 doubles, index 0 unused */
    double *spline; /* (n+1)*7 doubles, row 0 unused */
} pot_table;

typedef struct pot_eam {
    int n_ele;
    int key[POT_MAX_ELE];      /* atomic numbers, file order */
    double mass[POT_MAX_ELE];
    double lat_const[POT_MAX_ELE];
    int n_rho, n_r;
    double d_rho, d_r, cutoff;
    pot_table embed[POT_MAX_ELE];           /* F(rho) */
    pot_table elec[POT_MAX_ELE];            /* rho(r) */
    pot_table phi[POT_MAX_ELE][POT_MAX_ELE]; /* r*phi(r); [i][j] and [j][i] alias the same storage */
} pot_eam;

/* Deterministic synthetic Fe-Cu-Ni setfl file in FeCuNi.eam.alloy format (the real file is not in the
 * reference tree: reference .gitignore:10, README.md:33-38). Zhou-Johnson-Wadley style analytic forms,
 * smoothly switched to 0 at `cutoff`. Returns 0 on success. */
int pot_write_synthetic_setfl(const char *path, int n_rho, double d_rho, int n_r, double d_r, double cutoff);

/* Parse a setfl file (LAMMPS "eam/alloy" format) and build the spline rows. NULL on error. */
pot_eam *pot_read_setfl(const char *path);

void pot_free(pot_eam *p);

/* index of species with atomic number `key`, or -1 */
int pot_index_of_key(const pot_eam *p, int key);

/* the three evaluators of the reference's `eam` class (see header comment) */
double pot_charge_density(const pot_eam *p, int key, double dist2);
double pot_d_embed_energy(const pot_eam *p, int key, double rho);
double pot_to_force(const pot_eam *p, int key_from, int key_to, double dist2, double df_from, double df_to);

/* energies (NOT in the reference, which never computes potential energy -- SURVEY.md section 5);
 * used only for the total-energy drift report. */
double pot_embed_energy(const pot_eam *p, int key, double rho);
double pot_pair_energy(const pot_eam *p, int key_from, int key_to, double dist2);

#ifdef __cplusplus
}
#endif
#endif
