// Dispatcher of the fused pipeline.  The kernels are templates on k (fused_impl.cuh); each supported k is
// compiled in its own translation unit (fused_k.cu with -DVS_FUSED_K=k, see Makefile) so the build parallelises.
#include "vs_internal.cuh"

namespace vs {

// VS_FUSED_EXP_K=<k>: experiment builds with a single k (make exp K=<k> TAG=<name> EXTRA=...)
#ifdef VS_FUSED_EXP_K
#define VS_EXPAND_X(X, KK) X(KK)
#define VS_FUSED_K_LIST(X) VS_EXPAND_X(X, VS_FUSED_EXP_K)
#else
#define VS_FUSED_K_LIST(X) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16) X(17) X(18) X(19) X(20)
#endif

#define VS_DECLARE(KK)                                                                                                         \
    int launch_fused_k##KK(vs_ctx *c, const SourceDev &src, const ScaleDev &s, const ObjectiveDev &o, uint64_t i_begin, \
                           uint64_t i_end, int flags, double *partials, const FusedReq *req, bool *finalized);          \
    bool fused_tail_k##KK(const vs_ctx *c, int flags);
VS_FUSED_K_LIST(VS_DECLARE)
#undef VS_DECLARE

bool fused_supported(int k, int objective, int flags) {
    (void)flags;
    if (objective == VS_OBJ_ISHIGAMI) return k == 3;
    if (objective != VS_OBJ_GFUNCTION) return false;
    switch (k) {
#define VS_CASE(KK) case KK:
        VS_FUSED_K_LIST(VS_CASE)
#undef VS_CASE
        return true;
    default:
        return false;
    }
}

bool fused_tail_supported(vs_ctx *c, int k, int objective, int flags) {
    if (!fused_supported(k, objective, flags)) return false;
    switch (k) {
#define VS_CASE(KK) case KK: return fused_tail_k##KK(c, flags);
        VS_FUSED_K_LIST(VS_CASE)
#undef VS_CASE
    }
    return false;
}

int launch_fused(vs_ctx *c, int k, const SourceDev &src, const ScaleDev &s, const ObjectiveDev &o, uint64_t i_begin,
                 uint64_t i_end, int flags, double *partials, const FusedReq *req, bool *finalized) {
    switch (k) {
#define VS_CASE(KK) case KK: return launch_fused_k##KK(c, src, s, o, i_begin, i_end, flags, partials, req, finalized);
        VS_FUSED_K_LIST(VS_CASE)
#undef VS_CASE
    }
    set_error("no fused kernel for k=%d", k);
    return VS_ERR_UNSUPPORTED;
}

}  // namespace vs
