// em_kernels.cu -- EM ("exact method") kernels.  Placeholder until the EM path lands: every entry point
// reports the missing feature loudly (no CPU fallback).
#include "engine_internal.cuh"

namespace nmchb {

int em_launch_points(nmch_engine *, cudaStream_t, const float *, const float *, const float *, int, double *,
                     float *, float *)
{
    return engine_fail(NMCH_ERR_ARG, "EM method not built yet");
}
int em_philox_compat_init(nmch_engine *) { return NMCH_OK; }
void em_release(nmch_engine *) {}

}  // namespace nmchb
