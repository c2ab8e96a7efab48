// fused x pass, collocated hyperelasticity operator GammaOperatorFourierCollocatedHyper (fg:19619)
#include "fft_xgreen.cuh"
FGB_DEFINE_W32_SETTER(fgb_w32_set_xg4)
int fgb_xg_colloc9(FGB_XG_ARGS) { return launch_x_green<9, 5>(ctx, base, G, estride, ninner, nouter, ostride, cstride, jbase, xo, pt); }
int fgb_xg_gradg0div9(FGB_XG_ARGS) { return launch_x_green<9, 11>(ctx, base, G, estride, ninner, nouter, ostride, cstride, jbase, xo, pt); }
