// api.cu -- error reporting, shape derivation and workspace sizing for libnnue_b200.
#include <stdio.h>
#include <string.h>

#include "common.cuh"
#include "plan.cuh"

namespace nnue {
static thread_local char g_cuda_err[256] = "";
void note_cuda_error(cudaError_t e, const char *what) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}
static unsigned long long g_launches = 0;
void note_launch() { __atomic_fetch_add(&g_launches, 1ULL, __ATOMIC_RELAXED); }
static int g_options[kNumOptions] = {/*ft_fwd_staging=*/1, /*ft_bwd_dw_owner=*/1, /*input_bwd_fused=*/1, /*input_bwd_variant=*/1, /*head_fused=*/1, /*ft_bwd_both=*/1, /*ft_mma=*/1, /*extract_tma=*/0, /*extract_fixed=*/1, /*ft_umma=*/1, /*input_bwd_swizzle=*/1, /*head_umma=*/1, /*q_tc_min_batch=*/2048, /*ft_form=*/0, /*ft_density_permille=*/400, /*ft_gather=*/1, /*ft_gather_slab=*/0, /*input_bwd_rows=*/1, /*ft_gather_units=*/1, /*input_bwd_onchip=*/0, /*q_cta_max_batch=*/592, /*q_conv_fixed=*/1, /*q_stack_fused=*/8192, /*head_mid=*/1, /*head_pair_epilogue=*/1, /*gemm_inline_a=*/0, /*conv_bwd_packed=*/1};
int get_option(int which) { return which >= 0 && which < kNumOptions ? g_options[which] : 0; }
}  // namespace nnue

extern "C" {

int nnue_set_option(const char *key, int value) {
    if (!key) return NNUE_ERR_INVALID_ARG;
    if (!strcmp(key, "ft_form")) { nnue::g_options[nnue::kOptFtForm] = value; return NNUE_OK; }
    if (!strcmp(key, "ft_density_permille")) { nnue::g_options[nnue::kOptFtDensity] = value < 1 ? 1 : (value > 1000 ? 1000 : value); return NNUE_OK; }
    if (!strcmp(key, "input_bwd_onchip")) { nnue::g_options[nnue::kOptInputFusedGbin] = value; return NNUE_OK; }
    if (!strcmp(key, "ft_gather_units")) { nnue::g_options[nnue::kOptGatherUnits] = value < 1 ? 1 : value; return NNUE_OK; }
    if (!strcmp(key, "input_bwd_rows")) { nnue::g_options[nnue::kOptInputRows] = value; return NNUE_OK; }
    if (!strcmp(key, "ft_gather")) { nnue::g_options[nnue::kOptFtGather] = value; return NNUE_OK; }
    if (!strcmp(key, "ft_gather_slab")) { nnue::g_options[nnue::kOptFtGatherVariant] = value; return NNUE_OK; }
    if (!strcmp(key, "ft_fwd_staging")) { nnue::g_options[nnue::kOptFtFwdStaging] = value; return NNUE_OK; }
    if (!strcmp(key, "ft_bwd_dw_owner")) { nnue::g_options[nnue::kOptDwOwner] = value; return NNUE_OK; }
    if (!strcmp(key, "input_bwd_variant")) { nnue::g_options[nnue::kOptInputVariant] = value; return NNUE_OK; }
    if (!strcmp(key, "q_cta_max_batch")) { nnue::g_options[nnue::kOptQCtaMaxBatch] = value; return NNUE_OK; }
    if (!strcmp(key, "q_stack_fused")) { nnue::g_options[nnue::kOptQStackFused] = value; return NNUE_OK; }
    if (!strcmp(key, "q_conv_fixed")) { nnue::g_options[nnue::kOptQConvFixed] = value; return NNUE_OK; }
    if (!strcmp(key, "q_tc_min_batch")) { nnue::g_options[nnue::kOptQTcMinBatch] = value; return NNUE_OK; }
    if (!strcmp(key, "head_mid")) { nnue::g_options[nnue::kOptHeadMid] = value; return NNUE_OK; }
    if (!strcmp(key, "head_pair_epilogue")) { nnue::g_options[nnue::kOptHeadPairEpi] = value; return NNUE_OK; }
    if (!strcmp(key, "gemm_inline_a")) { nnue::g_options[nnue::kOptInlineA] = value; return NNUE_OK; }
    if (!strcmp(key, "conv_bwd_packed")) { nnue::g_options[nnue::kOptConvBwdPacked] = value; return NNUE_OK; }
    if (!strcmp(key, "head_umma")) { nnue::g_options[nnue::kOptHeadUmma] = value; return NNUE_OK; }
    if (!strcmp(key, "input_bwd_swizzle")) { nnue::g_options[nnue::kOptInputSwizzle] = value; return NNUE_OK; }
    if (!strcmp(key, "ft_umma")) { nnue::g_options[nnue::kOptFtUmma] = value; return NNUE_OK; }
    if (!strcmp(key, "extract_fixed")) { nnue::g_options[nnue::kOptExtractFixed] = value; return NNUE_OK; }
    if (!strcmp(key, "extract_tma")) { nnue::g_options[nnue::kOptExtractTma] = value; return NNUE_OK; }
    if (!strcmp(key, "ft_mma")) { nnue::g_options[nnue::kOptFtMma] = value; return NNUE_OK; }
    if (!strcmp(key, "ft_bwd_both")) { nnue::g_options[nnue::kOptFtBwdBoth] = value; return NNUE_OK; }
    if (!strcmp(key, "head_fused")) { nnue::g_options[nnue::kOptHeadFused] = value; return NNUE_OK; }
    if (!strcmp(key, "input_bwd_fused")) { nnue::g_options[nnue::kOptInputFused] = value; return NNUE_OK; }
    return NNUE_ERR_INVALID_ARG;
}

unsigned long long nnue_launch_count(int reset) {
    const unsigned long long v = __atomic_load_n(&nnue::g_launches, __ATOMIC_RELAXED);
    if (reset) __atomic_store_n(&nnue::g_launches, 0ULL, __ATOMIC_RELAXED);
    return v;
}

unsigned long long nnue_launch_count_add(unsigned long long n) {
    return __atomic_add_fetch(&nnue::g_launches, n, __ATOMIC_RELAXED);
}

int nnue_b200_abi_version(void) { return NNUE_B200_ABI_VERSION; }

const char *nnue_error_string(int code) {
    switch (code) {
        case NNUE_OK: return "ok";
        case NNUE_ERR_INVALID_ARG: return "invalid argument";
        case NNUE_ERR_UNSUPPORTED: return "shape not supported by the sm_100a kernels";
        case NNUE_ERR_CUDA: return "CUDA error (see nnue_last_cuda_error)";
        case NNUE_ERR_IO: return "cannot read .nnue file";
        case NNUE_ERR_FORMAT: return "malformed .nnue file";
        case NNUE_ERR_WORKSPACE: return "workspace too small";
        case NNUE_ERR_RASTER: return "conv raster exceeds the feature buffer";
        default: return "unknown error";
    }
}

const char *nnue_last_cuda_error(void) { return nnue::g_cuda_err; }

int nnue_shape_init(nnue_shape *s, int B, int H, int W, int C, int G, int L1, int L2, int L3, int NC, int stride) {
    if (!s || B < 1 || H < 1 || W < 1 || C < 1 || G < 1 || L1 < 2 || L2 < 1 || L3 < 1 || NC < 1 || stride < 1)
        return NNUE_ERR_INVALID_ARG;
    if (L1 % 2) return NNUE_ERR_UNSUPPORTED;  // torch.split(l0, L1//2) must give exactly two halves (nnue.py:660)
    memset(s, 0, sizeof(*s));
    s->B = B; s->H = H; s->W = W; s->C = C; s->G = G;
    s->L1 = L1; s->L2 = L2; s->L3 = L3; s->NC = NC; s->stride = stride;
    const long long F = 1LL * G * G * C;
    s->Gh = (H - 1) / stride + 1;  // (H + 2*1 - 3)/stride + 1
    s->Gw = (W - 1) / stride + 1;
    const long long P = 1LL * C * s->Gh * s->Gw;
    s->CW = nnue::ceil_div(s->Gh * s->Gw, 32);
    const long long NW = 1LL * C * s->CW;
    if (F > (1LL << 30) || P > (1LL << 30) || NW * 32 > (1LL << 30)) return NNUE_ERR_UNSUPPORTED;
    s->F = (int)F; s->P = (int)P; s->NW = (int)NW; s->PP = (int)(NW * 32);
    s->BW = nnue::ceil_div(B, 32);
    return NNUE_OK;
}

size_t nnue_workspace_bytes(const nnue_shape *s) {
    if (!s) return 0;
    size_t m = nnue::ws_head_bwd(*s);
    size_t v = nnue::ws_ft_bwd_dw(*s); if (v > m) m = v;
    v = nnue::ws_extract_bwd(*s); if (v > m) m = v;
    v = nnue::ws_ft_bwd_dval(*s); if (v > m) m = v;
    v = nnue::ws_input_bwd(*s); if (v > m) m = v;
    v = nnue::ws_head_train(*s); if (v > m) m = v;
    v = nnue::ws_ft_bwd_both(*s); if (v > m) m = v;
    v = nnue::ws_ft_bwd_mma(*s); if (v > m) m = v;
    v = nnue::ws_ft_fwd(*s); if (v > m) m = v;
    v = nnue::ws_ft_gather_fwd(*s); if (v > m) m = v;
    v = nnue::ws_ft_bwd_umma(*s); if (v > m) m = v;
    v = nnue::plan_input_bwd_fused(*s).ws_bytes; if (v > m) m = v;
    v = nnue::ws_ce(s->B); if (v > m) m = v;
    return m + 256;
}

}  // extern "C"
