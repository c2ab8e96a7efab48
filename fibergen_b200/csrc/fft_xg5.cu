// fused x pass, G0DivOperatorFourierHyper (fg:20155) and Willot's rotated-scheme operator GammaOperatorFourierWillotR (fg:19083)
#include "fft_xgreen.cuh"
FGB_DEFINE_W32_SETTER(fgb_w32_set_xg5)
int fgb_xg_g0div9(FGB_XG_ARGS) { return launch_x_green<9, 6>(ctx, base, G, estride, ninner, nouter, ostride, cstride, jbase, xo, pt); }
int fgb_xg_willot6(FGB_XG_ARGS) { return launch_x_green<6, 8>(ctx, base, G, estride, ninner, nouter, ostride, cstride, jbase, xo, pt); }
