// pose.cu — placeholder until the match/pose kernels land.
#include "common.cuh"
size_t movfe_pose_scratch_bytes(const movfe_ctx *) { return 0; }
int movfe_track_poses_launch(movfe_ctx *ctx, int64_t, int) { MOVFE_FAIL(ctx, MOVFE_E_STATE, "pose: not built yet"); }
#define NYI(ctx) do { if (!(ctx)) return MOVFE_E_INVALID; MOVFE_FAIL(ctx, MOVFE_E_STATE, "not built yet"); } while (0)
extern "C" int movfe_set_camera(movfe_ctx *ctx, const movfe_camera *, const movfe_pose_params *, float) { NYI(ctx); }
extern "C" int movfe_set_map_points(movfe_ctx *ctx, int, const movfe_map_point *, int, int) { NYI(ctx); }
extern "C" int movfe_set_pose(movfe_ctx *ctx, int, const movfe_pose *) { NYI(ctx); }
extern "C" int movfe_track_poses(movfe_ctx *ctx, int64_t, int) { NYI(ctx); }
extern "C" int movfe_download_poses(movfe_ctx *ctx, int64_t, int, movfe_pose *, int32_t *) { NYI(ctx); }
extern "C" int movfe_download_matches(movfe_ctx *ctx, int, int64_t, int32_t *, uint8_t *, int) { NYI(ctx); }
extern "C" int movfe_frustum(movfe_ctx *ctx, int, const movfe_pose *, const movfe_map_point *, const int32_t *, movfe_projection *) { NYI(ctx); }
extern "C" int movfe_join(movfe_ctx *ctx, int, const int32_t *, const int32_t *, const int32_t *, const uint8_t *, const int32_t *, int32_t *, int32_t *) { NYI(ctx); }
extern "C" int movfe_pose_optimize(movfe_ctx *ctx, int, const movfe_camera *, const movfe_pose_params *, const float *, const float *, const int32_t *, movfe_pose *, uint8_t *, int32_t *, int32_t *) { NYI(ctx); }
