// oracle/shim/gen/arch_building_config.h -- TEST INFRASTRUCTURE. Generated-header stand-in
// (reference src/arch/arch_building_config.h.in). REF_WITH_CUDA_HOOKS selects the accelerated build in which
// atom::latRho/latDf/latForce call the cuda_* hooks (arch_cuda/cuda_hooks.cpp) instead of the CPU loops.
#ifndef MISA_MD_ARCH_BUILDING_CONFIG_H
#define MISA_MD_ARCH_BUILDING_CONFIG_H
#ifdef REF_WITH_CUDA_HOOKS
#define ACCELERATE_ENABLED
#define ARCH_NAME cuda
#define ARCH_CUDA
#endif
#endif
