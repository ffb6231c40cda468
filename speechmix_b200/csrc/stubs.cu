// Temporary: entry points declared in the header whose kernels are not written yet.
#include "../../include/speechmix_sm100.h"
#include "host_common.h"
#define NOTIMPL(name) return smx::set_error(name ": not implemented yet")
extern "C" {
int smx_layernorm_fwd(const void*, const void*, const float*, const float*, void*, void*, float*, float*, int64_t, int64_t, float, int, void*) { NOTIMPL("smx_layernorm_fwd"); }
int smx_layernorm_bwd(const void*, const void*, const float*, const float*, const float*, const void*, void*, float*, float*, int64_t, int64_t, int, void*) { NOTIMPL("smx_layernorm_bwd"); }
int smx_colsum(const void*, float*, int64_t, int64_t, int64_t, void*) { NOTIMPL("smx_colsum"); }
int smx_cast_f32_to_bf16(const float*, void*, int64_t, void*) { NOTIMPL("smx_cast"); }
int smx_add_bf16(const void*, const void*, void*, int64_t, void*) { NOTIMPL("smx_add"); }
int smx_act_bf16(const void*, void*, int64_t, int, void*) { NOTIMPL("smx_act"); }
int smx_pack_conv_weight(const float*, void*, int64_t, int64_t, int64_t, void*) { NOTIMPL("smx_pack"); }
int smx_unpack_conv_wgrad(const float*, float*, int64_t, int64_t, int64_t, void*) { NOTIMPL("smx_unpack"); }
int smx_conv0_stats(const float*, const float*, float*, float*, int64_t, int64_t, int64_t, int, int, int, float, void*) { NOTIMPL("smx_conv0_stats"); }
int smx_conv0_gn_gelu_fwd(const float*, const float*, const float*, const float*, const float*, void*, int64_t, int64_t, int64_t, int, int, int, void*) { NOTIMPL("conv0 fwd"); }
int smx_conv0_gn_gelu_bwd(const float*, const float*, const float*, const float*, const float*, const void*, float*, int64_t, int64_t, int64_t, int, int, int, void*) { NOTIMPL("conv0 bwd"); }
int smx_posconv_fwd(const void*, const void*, const float*, void*, void*, int64_t, int64_t, int, int, int, int, void*) { NOTIMPL("posconv fwd"); }
int smx_posconv_dgrad(const void*, const void*, void*, int64_t, int64_t, int, int, int, void*) { NOTIMPL("posconv dgrad"); }
int smx_posconv_wgrad(const void*, const void*, float*, int64_t, int64_t, int, int, int, void*) { NOTIMPL("posconv wgrad"); }
int smx_attn_fwd(const SmxAttn*, void*) { NOTIMPL("attn fwd"); }
int smx_attn_bwd(const SmxAttn*, void*) { NOTIMPL("attn bwd"); }
int smx_embed_fwd(const int64_t*, const float*, const float*, void*, int64_t, int64_t, int64_t, float, int64_t, int64_t, void*) { NOTIMPL("embed fwd"); }
int smx_embed_bwd(const int64_t*, const void*, float*, float*, int64_t, int64_t, int64_t, float, int64_t, void*) { NOTIMPL("embed bwd"); }
size_t smx_lmhead_ws_bytes(int64_t, int64_t) { return 0; }
int smx_lmhead_ce_fwd(const void*, const void*, const float*, const int64_t*, float*, int64_t*, float*, float*, float*, void*, int64_t, int64_t, int64_t, float, int64_t, void*) { NOTIMPL("lmhead fwd"); }
int smx_lmhead_dlogits(const void*, const void*, const float*, const int64_t*, const float*, const float*, void*, int64_t, int64_t, int64_t, int64_t, int64_t, float, int64_t, void*) { NOTIMPL("lmhead dlogits"); }
int smx_weighted_sum_fwd(const void* const*, const float*, void*, int, int64_t, void*) { NOTIMPL("wsum fwd"); }
int smx_weighted_sum_bwd_w(const void* const*, const void*, float*, int, int64_t, void*) { NOTIMPL("wsum bwd"); }
}
