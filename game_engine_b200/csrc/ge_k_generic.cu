// ge_k_generic.cu — the run-time-table step kernels of ONE (family, record bucket): compiled once per pair with
// -DGE_TU_FAM=<1 werewolf | 2 TTL> -DGE_TU_BUCKET=<players bucket> (build.py), see ge_kernels.h.
#include "ge_step_tps.cuh"
#include "ge_step_coop.cuh"
#include "ge_kernels.h"

#if !defined(GE_TU_FAM) || !defined(GE_TU_BUCKET)
#error "compile with -DGE_TU_FAM=.. -DGE_TU_BUCKET=.."
#endif
#define GE_CAT3_(a, b, c) a##b##_##c
#define GE_CAT3(a, b, c) GE_CAT3_(a, b, c)

using namespace ge;

void GE_CAT3(ge_kernel_set_, GE_TU_FAM, GE_TU_BUCKET)(KernelSet* out) {
    constexpr int B = GE_TU_BUCKET;
    KernelSet k{};
#if GE_TU_FAM == 1
    k.tps = (step_fn)k_step_w_tps<B>;
    k.coop = (step_fn)k_step_w_coop<B>;
    k.tiled = (step_fn)k_step_w_tps_tiled<B>;
    k.human = (step_fn)k_step_w_tps_h<B>;
    k.ring = (ring_fn)k_ring_w_tps<B>;
#if GE_TU_BUCKET == 8 || GE_TU_BUCKET == 16
    k.tps_pk = (step_fn)k_step_w_tps<B, void, true>;
    k.ring_pk = (ring_fn)k_ring_w_tps<B, void, true>;
#endif
#else
    k.tps = (step_fn)k_step_t_tps<B>;
    k.coop = (step_fn)k_step_t_coop<B>;
    k.human = (step_fn)k_step_t_tps_h<B>;
    k.ring = (ring_fn)k_ring_t_tps<B>;
#endif
    *out = k;
}
