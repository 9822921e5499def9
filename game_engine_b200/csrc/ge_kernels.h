// ge_kernels.h — registry of the step kernels.
//
// The step kernels are templates over the record bucket and (for the shipped games) a build-time table; one nvcc
// translation unit holding all of them took 5.5 minutes to compile.  They are therefore instantiated in separate
// translation units — ge_k_generic.cu once per (family, bucket), ge_k_spec.cu once per shipped table (both compiled
// several times with different -D flags by build.py, in parallel) — and the host runtime (ge_capi.cu) reaches them
// through the plain function pointers below.  No relocatable device code: every kernel lives wholly in its own unit.
#pragma once
#include <cstddef>
#include "ge_common.cuh"

namespace ge {

typedef void (*step_fn)(const DevTable, const StepArgs);
typedef void (*ring_fn)(const DevTable, const StepArgs, const RingArgs);

// the kernels of one (family, record bucket) that interpret the run-time table; NULL = the variant does not exist
struct KernelSet {
    step_fn tps;        // thread per session (default mapping)
    step_fn coop;       // lane per player (the north-star mapping)
    step_fn tiled;      // tps with per-tile column needs (werewolf; batches with phase regrouping)
    step_fn human;      // tps with the human-seat path (SPEC D3h)
    ring_fn ring;       // one launch for a ring of batches
    step_fn tps_pk;     // tps over the PACKED session store (werewolf, up to 16 players)
    ring_fn ring_pk;
};
// build-time specialised twins for one shipped table, matched at run time by a byte-identical blob
struct SpecKernels {
    const unsigned char* blob;
    size_t len;
    step_fn tps, tiled, tps_pk;
    ring_fn ring, ring_pk;
};

}  // namespace ge

// one per (family, bucket): ge_k_generic.cu with -DGE_TU_FAM / -DGE_TU_BUCKET
#define GE_KERNEL_SETS(X) X(1, 8) X(1, 16) X(1, 24) X(1, 32) X(2, 4) X(2, 8) X(2, 16) X(2, 32)
#define GE_DECL_KSET(F, B) void ge_kernel_set_##F##_##B(ge::KernelSet* out);
GE_KERNEL_SETS(GE_DECL_KSET)
#undef GE_DECL_KSET
