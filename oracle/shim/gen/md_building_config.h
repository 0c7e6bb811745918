// oracle/shim/gen/md_building_config.h -- TEST INFRASTRUCTURE. What CMake generates from the reference's
// src/md_building_config.h.in for a release build with the default MD_RAND=MT (config.cmake:22,31-36).
#ifndef MISA_MD_PRE_CONFIG_H
#define MISA_MD_PRE_CONFIG_H
/* #undef MD_DEV_MODE */
#define RAND_MT
#define LAT_CUTOFF 0.5
namespace config {
    const double nei_lat_cutoff = LAT_CUTOFF;
}
#endif
