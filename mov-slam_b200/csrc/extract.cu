// extract.cu — placeholder until the propagation kernels land (next commit).
#include "common.cuh"
size_t movfe_extract_scratch_bytes(const movfe_ctx *) { return 0; }
int movfe_extract_launch(movfe_ctx *ctx, int64_t, int) { MOVFE_FAIL(ctx, MOVFE_E_STATE, "extract: not built yet"); }
#define NYI(ctx) do { if (!(ctx)) return MOVFE_E_INVALID; MOVFE_FAIL(ctx, MOVFE_E_STATE, "not built yet"); } while (0)
extern "C" int movfe_set_tracks(movfe_ctx *ctx, int, const movfe_track *, int, int32_t) { NYI(ctx); }
extern "C" int movfe_extract(movfe_ctx *ctx, int64_t, int) { NYI(ctx); }
extern "C" int movfe_track_count(movfe_ctx *ctx, int, int64_t, int32_t *, int32_t *) { NYI(ctx); }
extern "C" int movfe_download_tracks(movfe_ctx *ctx, int, int64_t, movfe_track *, int) { NYI(ctx); }
