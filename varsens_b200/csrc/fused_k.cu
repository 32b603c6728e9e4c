// One translation unit per compile-time k of the fused pipeline (nvcc -DVS_FUSED_K=<k>, see Makefile).
#include "fused_impl.cuh"

#ifndef VS_FUSED_K
#error "compile with -DVS_FUSED_K=<k>"
#endif
#define VS_CAT2(a, b) a##b
#define VS_CAT(a, b) VS_CAT2(a, b)

namespace vs {
int VS_CAT(launch_fused_k, VS_FUSED_K)(vs_ctx *c, const SourceDev &src, const ScaleDev &s, const ObjectiveDev &o, uint64_t i_begin,
                                       uint64_t i_end, int flags, double *partials, const FusedReq *req, bool *finalized) {
    return dispatch_k<VS_FUSED_K>(c, src, s, o, i_begin, i_end, flags, partials, req, finalized);
}
bool VS_CAT(fused_tail_k, VS_FUSED_K)(const vs_ctx *c, int flags) { return tail_supported_k<VS_FUSED_K>(c, flags); }
}  // namespace vs
