// Temporary: entry points declared in the header whose kernels are not written yet.
#include "../../include/speechmix_sm100.h"
#include "host_common.h"
#define NOTIMPL(name) return smx::set_error(name ": not implemented yet")
extern "C" {
int smx_conv0_stats(const float*, const float*, float*, float*, int64_t, int64_t, int64_t, int, int, int, float, void*) { NOTIMPL("smx_conv0_stats"); }
int smx_conv0_gn_gelu_fwd(const float*, const float*, const float*, const float*, const float*, void*, int64_t, int64_t, int64_t, int, int, int, void*) { NOTIMPL("conv0 fwd"); }
int smx_conv0_gn_gelu_bwd(const float*, const float*, const float*, const float*, const float*, const void*, float*, int64_t, int64_t, int64_t, int, int, int, void*) { NOTIMPL("conv0 bwd"); }
int smx_posconv_fwd(const void*, const void*, const float*, void*, void*, int64_t, int64_t, int, int, int, int, void*) { NOTIMPL("posconv fwd"); }
int smx_posconv_dgrad(const void*, const void*, void*, int64_t, int64_t, int, int, int, void*) { NOTIMPL("posconv dgrad"); }
int smx_posconv_wgrad(const void*, const void*, float*, int64_t, int64_t, int, int, int, void*) { NOTIMPL("posconv wgrad"); }
size_t smx_lmhead_ws_bytes(int64_t, int64_t) { return 0; }
int smx_lmhead_ce_fwd(const void*, const void*, const float*, const int64_t*, float*, int64_t*, float*, float*, float*, void*, int64_t, int64_t, int64_t, float, int64_t, void*) { NOTIMPL("lmhead fwd"); }
int smx_lmhead_dlogits(const void*, const void*, const float*, const int64_t*, const float*, const float*, void*, int64_t, int64_t, int64_t, int64_t, int64_t, float, int64_t, void*) { NOTIMPL("lmhead dlogits"); }
}
