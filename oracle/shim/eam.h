// oracle/shim/eam.h -- TEST INFRASTRUCTURE. Stand-in for libpot v0.1.0's `eam` class (reference pkg.yaml:16,
// source absent): the three evaluators the hot path calls, forwarded to oracle/pot.c ("parity unpinned").
#ifndef ORACLE_SHIM_EAM_H
#define ORACLE_SHIM_EAM_H
#include "pot.h"

class eam {
public:
    explicit eam(const pot_eam *p) : p(p) {}
    inline double chargeDensity(const unsigned short key, const double dist2) { return pot_charge_density(p, key, dist2); }
    inline double dEmbedEnergy(const unsigned short key, const double rho) { return pot_d_embed_energy(p, key, rho); }
    inline double toForce(const unsigned short key_from, const unsigned short key_to, const double dist2, const double df_from,
                          const double df_to) {
        return pot_to_force(p, key_from, key_to, dist2, df_from, df_to);
    }
    const pot_eam *p;
};
#endif
