// ge_k_spec.cu — the build-time specialised step kernels of ONE shipped table (ge_spec_gen.cuh): compiled once per
// table with -DGE_SPEC_INDEX=<position in specgen.SPECS> (build.py), see ge_kernels.h.
#include "ge_step_tps.cuh"
#include "ge_spec_gen.cuh"
#include "ge_kernels.h"

#ifndef GE_SPEC_INDEX
#error "compile with -DGE_SPEC_INDEX=.."
#endif
#define GE_CAT2_(a, b) a##b
#define GE_CAT2(a, b) GE_CAT2_(a, b)

using namespace ge;

template <int FAM, int BUCKET, class S> struct SpecKernelsOf;
template <int BUCKET, class S> struct SpecKernelsOf<FAM_WEREWOLF, BUCKET, S> {
    static void fill(SpecKernels* e) {
        e->tps = (step_fn)k_step_w_tps<BUCKET, S>;
        e->tiled = (step_fn)k_step_w_tps_tiled<BUCKET, S>;
        e->ring = (ring_fn)k_ring_w_tps<BUCKET, S>;
        if constexpr (BUCKET == 8 || BUCKET == 16) {
            e->tps_pk = (step_fn)k_step_w_tps<BUCKET, S, true>;
            e->ring_pk = (ring_fn)k_ring_w_tps<BUCKET, S, true>;
        }
    }
};
template <int BUCKET, class S> struct SpecKernelsOf<FAM_TTL, BUCKET, S> {
    static void fill(SpecKernels* e) {
        e->tps = (step_fn)k_step_t_tps<BUCKET, S>;
        e->ring = (ring_fn)k_ring_t_tps<BUCKET, S>;
    }
};

#define GE_DEFINE_SPEC(S, FAM, BUCKET)                                      \
    void ge_spec_kernels_##S(SpecKernels* e) {                              \
        *e = SpecKernels{};                                                 \
        e->blob = spec::S##_blob;                                           \
        e->len = sizeof(spec::S##_blob);                                    \
        SpecKernelsOf<FAM, BUCKET, spec::S>::fill(e);                       \
    }
GE_CAT2(GE_SPEC_AT_, GE_SPEC_INDEX)(GE_DEFINE_SPEC)
