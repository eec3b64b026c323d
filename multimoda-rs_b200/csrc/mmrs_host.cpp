// placeholder — replaced by the host-side orchestration (see next commit)
#include "mmrs_internal.hpp"
#include <cstdlib>
extern "C" void mmrs_free(void* p) { std::free(p); }
extern "C" int mmrs_geometry_from_dir(mmrs_ctx* ctx, const char*, const char*, int, double, double, double, uint32_t, double**, int64_t*) { return mmrs::set_err(ctx, MMRS_ERR_STATE, "not built yet"); }
extern "C" int mmrs_geometry_from_arrays(mmrs_ctx* ctx, const double*, int64_t, const double*, int64_t, const double*, int64_t, const double*, int64_t, const double*, int64_t, const double*, int, const char*, double, double, double, uint32_t, double**, int64_t*) { return mmrs::set_err(ctx, MMRS_ERR_STATE, "not built yet"); }
extern "C" int mmrs_process_cases(mmrs_ctx* ctx, int32_t, int64_t, const double* const*, const int64_t*, const mmrs_align_params*, double**, int64_t*, double**, int64_t*, int32_t*) { return mmrs::set_err(ctx, MMRS_ERR_STATE, "not built yet"); }
extern "C" int mmrs_process_stats(mmrs_ctx* ctx, int64_t s[5]) { for (int i = 0; i < 5; ++i) s[i] = ctx ? ctx->stats[i] : 0; return 0; }
