// temporary stubs (replaced as the kernels land)
#include "stocs_ctx.h"
extern "C" {
#define NI(ctx) do { if (ctx) (ctx)->err = "not implemented yet"; return STOCS_E_STATE; } while (0)
int stocs_b200_sample_bases(stocs_b200_ctx* ctx, uint64_t, uint32_t, int, int32_t*, float*, uint8_t*) { NI(ctx); }
int stocs_b200_find_congruent(stocs_b200_ctx* ctx, int, const int32_t*, const float*, int32_t*, int64_t, int64_t*) { NI(ctx); }
int stocs_b200_fit_transforms(stocs_b200_ctx* ctx, int64_t, const int32_t*, const int32_t*, float*, float*, uint8_t*) { NI(ctx); }
int stocs_b200_run_pipeline(stocs_b200_ctx* ctx, uint64_t, int, int, stocs_b200_pipeline_result*) { NI(ctx); }
}
