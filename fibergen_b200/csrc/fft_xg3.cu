// fused x pass, collocated heat operator GammaOperatorFourierCollocatedHeat (fg:19302) and GradOperatorFourierHyper (fg:22069)
#include "fft_xgreen.cuh"
FGB_DEFINE_W32_SETTER(fgb_w32_set_xg3)
int fgb_xg_colloc3(FGB_XG_ARGS) { return launch_x_green<3, 4>(ctx, base, G, estride, ninner, nouter, ostride, cstride, jbase, xo, pt); }
int fgb_xg_grad9(FGB_XG_ARGS) { return launch_x_green<9, 7>(ctx, base, G, estride, ninner, nouter, ostride, cstride, jbase, xo, pt); }
