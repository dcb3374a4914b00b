// consume_inst.cu -- the specialised consume kernels of ONE k (see klist.h).
// Built once per listed k with -DOXG_INST_K=k; exports a single entry point that hands the
// kernels' addresses to capi.cu, which launches them with cudaLaunchKernel.
#include "consume.cuh"
#include "scatter.cuh"
#include "klist.h"

#ifndef OXG_INST_K
#define OXG_INST_K 31  // so that a bare `nvcc -c` of this file still compiles something
#endif

#define OXG_CAT2(a, b) a##b
#define OXG_CAT(a, b) OXG_CAT2(a, b)

extern "C" __attribute__((visibility("hidden"))) const void *OXG_CAT(oxg_consume_entry_, OXG_INST_K)(int mode) {
    using namespace oxg;
    switch (mode) {
    case kModeCount: return reinterpret_cast<const void *>(&consume_kernel<OXG_INST_K, kModeCount>);
    case kModeHash: return reinterpret_cast<const void *>(&consume_kernel<OXG_INST_K, kModeHash>);
    case kModeFirstBad: return reinterpret_cast<const void *>(&consume_kernel<OXG_INST_K, kModeFirstBad>);
    case kModePart: return reinterpret_cast<const void *>(&scatter_kernel<OXG_INST_K>);  // pass A of the partitioned pipeline
    default: return nullptr;
    }
}
