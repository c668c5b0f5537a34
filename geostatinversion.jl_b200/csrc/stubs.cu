// Entry points declared in include/gsi_b200.h whose device implementation has not
// landed yet: they fail loudly (GSI_ERR_UNSUPPORTED) -- never a CPU fallback.
#include "common.cuh"
#define GSI_API extern "C" __attribute__((visibility("default")))

static int32_t unsupported(const char* name) {
    gsi::set_last_error(std::string(name) + ": not implemented yet in this build");
    return GSI_ERR_UNSUPPORTED;
}

GSI_API int32_t gsi_rangefinder_adaptive(gsi_op*, const gsi_buf*, const gsi_buf*, double, int64_t, gsi_buf*, int64_t*) {
    return unsupported("gsi_rangefinder_adaptive");
}
GSI_API int32_t gsi_eig_nystrom(gsi_op*, const gsi_buf*, gsi_buf*, double*) { return unsupported("gsi_eig_nystrom"); }
GSI_API int32_t gsi_pcga_lowrank_matvec(gsi_ctx*, int64_t, int64_t, const double*, int64_t, const double*, const double*,
                                        const double*, int64_t, const double*, double*) {
    return unsupported("gsi_pcga_lowrank_matvec");
}
GSI_API int32_t gsi_pcga_lsqr_solve(gsi_ctx*, int64_t, int64_t, const double*, int64_t, const double*, const double*,
                                    const double*, int64_t, const double*, double, double, double, int64_t, double*,
                                    int64_t*, int32_t*) {
    return unsupported("gsi_pcga_lsqr_solve");
}
GSI_API int32_t gsi_pcga_update(gsi_ctx*, const gsi_buf*, int64_t, const double*, const double*, int64_t, int64_t,
                                const double*, double*) {
    return unsupported("gsi_pcga_update");
}
GSI_API int32_t gsi_pcga_paramstorun(gsi_ctx*, const gsi_buf*, int64_t, const double*, const double*, double, gsi_buf*) {
    return unsupported("gsi_pcga_paramstorun");
}
GSI_API int32_t gsi_sketch_apply(gsi_ctx*, gsi_buf*, const gsi_buf*, gsi_buf*) { return unsupported("gsi_sketch_apply"); }
GSI_API int32_t gsi_sketch_cov(gsi_ctx*, gsi_buf*, const double*, double*, int64_t) { return unsupported("gsi_sketch_cov"); }
