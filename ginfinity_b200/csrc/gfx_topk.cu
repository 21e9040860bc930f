// K5 similarity search -- placeholder until the encoder path is verified.
#include "gfx_common.cuh"
extern "C" size_t gfx_topk_workspace_bytes(int64_t, int64_t, int) { return 0; }
extern "C" int gfx_topk(const void *, int64_t, const void *, int64_t, int, int, int, int64_t,
                        float *, int64_t *, void *, size_t, void *) {
  return gfx::fail(GFX_ERR_UNSUPPORTED, "gfx_topk: not built yet");
}
extern "C" int gfx_topk_merge(const float *, const int64_t *, int, int64_t, int, float *,
                              int64_t *, void *) {
  return gfx::fail(GFX_ERR_UNSUPPORTED, "gfx_topk_merge: not built yet");
}
