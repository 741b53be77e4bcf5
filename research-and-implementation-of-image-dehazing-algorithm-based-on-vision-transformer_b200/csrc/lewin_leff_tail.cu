// Translation unit of the fused LeFF-tail kernel (see leff_tail.cuh).
#define LEWIN_TU_LITE 1      // no GELU tables / non-template kernels in this unit (they live in lewin_abi.cu)
#include "leff_tail_api.h"
#include "leff_tail.cuh"

namespace lewin {

bool leff_tail_supported(const LewinLeffFwdArgs* a) {
    return a->fused && !a->save_for_backward && lt::supported(a->C, a->hidden, a->B, a->H, a->W);
}

int leff_tail_launch(const LewinLeffFwdArgs* a, const uint16_t* gelu_tab2, int num_sms, cudaStream_t stream) {
    lt::Args k{};
    k.h1 = static_cast<const __nv_bfloat16*>(a->h1);
    k.resid = static_cast<const __nv_bfloat16*>(a->y);
    k.out = static_cast<__nv_bfloat16*>(a->out);
    k.w_dw = a->w_dw; k.b_dw = a->b_dw; k.w2 = a->w2; k.b2 = a->b2;
    k.drop_scale = a->drop_scale;
    k.gelu_tab2 = gelu_tab2;
    k.B = a->B; k.H = a->H; k.W = a->W;
    k.ld_out = a->ld_out > 0 ? a->ld_out : a->C;
    return static_cast<int>(lt::launch(a->C, k, num_sms, stream));
}

}  // namespace lewin
