// Multi-GPU plumbing: x-slab partition, all-to-all transposes around the fused x pass, halo planes for the
// staggered stencils and rank-ordered reductions.  One process per GPU; NCCL over NVLink.
#include "fgb_internal.h"

int fgb_comm_free(fgb_ctx* ctx) { (void)ctx; return FGB_OK; }

int fgb_comm_fft_x(fgb_ctx* ctx, double*, int, const FftLayout&, const GreenArgs*) {
    return fgb_fail(ctx, FGB_EUNSUPPORTED, "slab-partitioned x pass not built yet");
}
int fgb_comm_halo_tau(fgb_ctx* ctx, const double*) { return fgb_fail(ctx, FGB_EUNSUPPORTED, "halo exchange not built yet"); }
int fgb_comm_halo_u(fgb_ctx* ctx) { return fgb_fail(ctx, FGB_EUNSUPPORTED, "halo exchange not built yet"); }
int fgb_allreduce_host(fgb_ctx* ctx, double*, int, int) { return fgb_fail(ctx, FGB_EUNSUPPORTED, "multi-rank reductions not built yet"); }

extern "C" int fgb_comm_unique_id(void*) { return FGB_EUNSUPPORTED; }
extern "C" int fgb_comm_init(fgb_ctx* ctx, const void*) { return fgb_fail(ctx, FGB_EUNSUPPORTED, "multi-GPU communicator not built yet"); }
