// HiFT vocoder entry points (placeholder while the kernels land; replaced by the real path).
#include "vt_common.cuh"
extern "C" {
int vt_hift_create(const vt_tensor*, int, int, vt_hift**) { vt::set_error("vt_hift: not built yet"); return VT_ERR_UNSUPPORTED; }
void vt_hift_destroy(vt_hift*) {}
int64_t vt_hift_workspace_bytes(const vt_hift*, int, int64_t, int64_t) { return VT_ERR_UNSUPPORTED; }
int vt_hift_forward(vt_hift*, const float*, const int32_t*, int, const float*, const float*, const float*, uint64_t,
                    float*, void*, int64_t, void*) { vt::set_error("vt_hift: not built yet"); return VT_ERR_UNSUPPORTED; }
int64_t vt_hift_read_tap(vt_hift*, const char*, int, float*, int64_t, void*, void*) { return VT_ERR_UNSUPPORTED; }
}
