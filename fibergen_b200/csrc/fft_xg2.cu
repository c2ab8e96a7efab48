// fused x pass, collocated elasticity operator GammaOperatorFourierCollocated (fg:19381) and its zero-trace form for the
// viscosity Delta operator (DeltaOperatorCollocated fg:20464, fftTensor(zero_trace) fg:18557)
#include "fft_xgreen.cuh"
FGB_DEFINE_W32_SETTER(fgb_w32_set_xg2)
int fgb_xg_colloc6(FGB_XG_ARGS) { return launch_x_green<6, 3>(ctx, base, G, estride, ninner, nouter, ostride, cstride, jbase, xo, pt); }
int fgb_xg_colloc6_zt(FGB_XG_ARGS) { return launch_x_green<6, 9>(ctx, base, G, estride, ninner, nouter, ostride, cstride, jbase, xo, pt); }
