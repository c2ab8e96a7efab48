// Host-side launch helpers shared by the FFT translation units.
#pragma once
#include "fgb_internal.h"

template <typename K>
static cudaError_t set_smem(K kernel, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// choose the number of lanes per tile so that nbuf buffers of n*T complex fit the opt-in shared memory
static inline int pick_T(const fgb_ctx* ctx, int n, int nbuf, int pad, int want) {
    int T = want;
    while (T > 1 && (size_t)nbuf * n * (T + pad) * sizeof(double2) > ctx->smem_optin) T /= 2;
    if ((size_t)nbuf * n * (T + pad) * sizeof(double2) > ctx->smem_optin) return 0;
    return T;
}

// Every translation unit that instantiates the register FFTs has its own copy of the __constant__ table c_w32 (fft_pow2.cuh);
// fgb_fft_init fills all of them through these hooks.
#define FGB_DEFINE_W32_SETTER(name)                                                                    \
    cudaError_t name(const double2* w32) { return cudaMemcpyToSymbol(c_w32, w32, sizeof(double2) * 32); }

// fused x pass of one operator kind (fft_xg*.cu); signature of launch_x_green<NC, KIND>
struct GreenDev;
#define FGB_XG_ARGS fgb_ctx *ctx, double2 *base, const GreenDev &G, long estride, int ninner, int nouter, long ostride, long cstride, \
                    int jbase, const PencilMap &xo, const PeerTable &pt
int fgb_xg_staggered3(FGB_XG_ARGS);       // kind 1, 3 components
int fgb_xg_staggered1(FGB_XG_ARGS);       // kind 2, 1 component (heat)
int fgb_xg_poisson1(FGB_XG_ARGS);         // kind 10, 1 component: poisson_solve
int fgb_xg_colloc6(FGB_XG_ARGS);          // kind 3
int fgb_xg_colloc3(FGB_XG_ARGS);          // kind 4
int fgb_xg_colloc9(FGB_XG_ARGS);          // kind 5
int fgb_xg_g0div9(FGB_XG_ARGS);           // kind 6: G0DivOperatorFourierHyper
int fgb_xg_grad9(FGB_XG_ARGS);            // kind 7: GradOperatorFourierHyper
int fgb_xg_gradg0div9(FGB_XG_ARGS);       // kind 11: GradOperatorFourierHyper o G0DivOperatorFourierHyper
int fgb_xg_willot6(FGB_XG_ARGS);          // kind 8: GammaOperatorFourierWillotR
int fgb_xg_colloc6_zt(FGB_XG_ARGS);       // kind 9: collocated elasticity operator on the zero-trace representation (viscosity)
cudaError_t fgb_w32_set_xg1(const double2*);
cudaError_t fgb_w32_set_xg2(const double2*);
cudaError_t fgb_w32_set_xg3(const double2*);
cudaError_t fgb_w32_set_xg4(const double2*);
cudaError_t fgb_w32_set_xg5(const double2*);
