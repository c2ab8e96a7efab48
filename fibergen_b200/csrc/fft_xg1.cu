// fused x pass, staggered-grid operators: G0OperatorFourierStaggeredGeneral (3 components, fg:19834) and ...Heat (1 component, fg:19778)
#include "fft_xgreen.cuh"
FGB_DEFINE_W32_SETTER(fgb_w32_set_xg1)
int fgb_xg_staggered3(FGB_XG_ARGS) { return launch_x_green<3, 1>(ctx, base, G, estride, ninner, nouter, ostride, cstride, jbase, xo, pt); }
int fgb_xg_staggered1(FGB_XG_ARGS) { return launch_x_green<1, 2>(ctx, base, G, estride, ninner, nouter, ostride, cstride, jbase, xo, pt); }
int fgb_xg_poisson1(FGB_XG_ARGS) { return launch_x_green<1, 10>(ctx, base, G, estride, ninner, nouter, ostride, cstride, jbase, xo, pt); }
