/* Offline stand-in for google/cpu_features' cpu_features_macros.h, which the reference pulls
 * with CMake FetchContent (deps/VectorSimilarity/cmake/cpu_features.cmake:4-9) and which is not
 * vendored under /root/reference.  TEST INFRASTRUCTURE ONLY: used to compile the reference's
 * own VecSim sources into oracle/_ref/libvecsim_ref.so.  Never linked into the product. */
#pragma once
#if defined(__x86_64__)
#define CPU_FEATURES_ARCH_X86_64 1
#endif
