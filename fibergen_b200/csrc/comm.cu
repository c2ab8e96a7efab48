// Multi-GPU plumbing: one process per GPU, x-slab partition, NCCL over NVLink 5 / NVSwitch.
//
// The reference has no distributed code at all; this is the new exchange step of the path:
//   * 3-D FFT of a slab-partitioned buffer: z and y passes are local; the y pass writes its output straight into the
//     all-to-all staging layout S[c][q][il][jl][k] (PencilMap), one grouped ncclSend/ncclRecv per (component, peer)
//     moves it to the y-slab layout R[c][ii][jl][k] on which the fused x pass (forward x, Green operator, inverse x)
//     runs in place; the way back is symmetric and the inverse y pass reads the staging layout directly.  No separate
//     pack / unpack sweeps touch HBM.
//   * staggered stencils: one x plane of the needed components from each slab neighbour (periodic).
//   * reductions: per-rank partial results are all-gathered and summed in rank order (deterministic).
// NCCL is resolved at run time (dlopen "libnccl.so.2": torch's bundled copy if the host process already loaded it, the
// system one otherwise), so libfgb200 has no link-time dependency on it.
#include "fgb_internal.h"
#include <dlfcn.h>
#include <cstdlib>
#include <nccl.h>

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int load_nccl(fgb_ctx* ctx) {
    if (g_nccl.lib) return FGB_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* n : names) {
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) return fgb_fail(ctx, FGB_ECOMM, "cannot load libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                                     \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(lib, name));               \
    if (!g_nccl.field) return fgb_fail(ctx, FGB_ECOMM, "libnccl lacks %s", name);
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(AllGather, "ncclAllGather")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.lib = lib;
    return FGB_OK;
}

#define FGB_NCCL(ctx, call)                                                                       \
    do {                                                                                          \
        ncclResult_t r__ = (call);                                                                \
        if (r__ != ncclSuccess)                                                                   \
            return fgb_fail(ctx, FGB_ECOMM, "%s failed: %s", #call, g_nccl.GetErrorString(r__));  \
    } while (0)

static int need_comm(fgb_ctx* ctx) {
    if (!ctx->nccl_comm) return fgb_fail(ctx, FGB_ECOMM, "context has %d ranks but fgb_comm_init was not called", ctx->nranks);
    return FGB_OK;
}

// The transposition buffers sbuf / xbuf hold `cap` doubles each.  Before the peers have mapped them they grow on demand; afterwards
// their addresses are fixed (exported with CUDA IPC), and a larger request -- only the 9-component Fourier operators of the
// hyperelastic post-processing on a staggered context -- gets a temporary pair and takes the NCCL all-to-all path.
static int ensure_xbuf(fgb_ctx* ctx, size_t need) {
    if (ctx->xbuf && ctx->xbuf_cap >= need) return FGB_OK;
    if (ctx->p2p) return fgb_fail(ctx, FGB_EINVAL, "transposition buffers are mapped by the peers and cannot grow");
    if (ctx->sbuf) cudaFree(ctx->sbuf);
    if (ctx->xbuf) cudaFree(ctx->xbuf);
    ctx->sbuf = ctx->xbuf = nullptr;
    ctx->xbuf_cap = 0;
    cudaError_t e = cudaMalloc(&ctx->sbuf, sizeof(double) * need);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->xbuf, sizeof(double) * need);
    if (e != cudaSuccess) return fgb_fail(ctx, FGB_ENOMEM, "cannot allocate the transpose buffers (%zu bytes each): %s", sizeof(double) * need, cudaGetErrorString(e));
    ctx->xbuf_cap = need;
    return FGB_OK;
}


// ---- synchronisation over peer memory --------------------------------------------------------------------------------------
// Every rank owns a small block (sync_base, mapped by all peers with CUDA IPC):
//   u64 bar[8]        bar[q]  = number of the last barrier rank q has entered
//   u64 exf[8]        exf[q]  = number of the last scalar exchange rank q has published
//   f64 exv[2][8][8]  exv[parity][q][i] = value i of rank q in the exchange with that parity
// A barrier / exchange is ONE single-warp kernel: lane q publishes to rank q (values, system fence, flag) and then polls its own
// block until rank q's flag has arrived.  It is stream ordered, so the stores of the preceding FFT kernel into peer memory have
// completed before the flag goes out, and the following kernel starts only after every peer has signalled.  The poll gives up
// after ~4 s and raises the context's error flag instead of hanging the device.
#define FGB_SYNC_BYTES 4096
struct SyncPeers {
    unsigned long long* p[8];
};
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__global__ void k_peer_barrier(SyncPeers P, int me, int n, unsigned long long seq, int* flag) {
    const int q = threadIdx.x;
    __threadfence_system();
    if (q < n) {
        st_release_sys(P.p[q] + me, seq);                    // bar[me] on rank q
        const unsigned long long* mine = P.p[me] + q;        // bar[q] on this rank
        const unsigned long long t0 = global_ns();
        while (ld_acquire_sys(mine) < seq) {
            if (global_ns() - t0 > 4000000000ull) { atomicOr(flag, 4); break; }
        }
    }
    __syncwarp();
    __threadfence_system();
}
// all-gather of nv <= 8 doubles per rank: out[q*nv + i] = value i of rank q
__global__ void k_peer_allgather(SyncPeers P, int me, int n, unsigned long long seq, const double* __restrict__ vals, int nv,
                                 double* __restrict__ out, int* flag) {
    const int q = threadIdx.x;
    const int par = (int)(seq & 1ull);
    if (q < n) {
        double* dst = reinterpret_cast<double*>(P.p[q] + 16) + ((size_t)par * 8 + me) * 8;      // exv[par][me][.] on rank q
        for (int i = 0; i < nv; i++) reinterpret_cast<volatile double*>(dst)[i] = vals[i];
        __threadfence_system();
        st_release_sys(P.p[q] + 8 + me, seq);                // exf[me] on rank q
        const unsigned long long* mine = P.p[me] + 8 + q;
        const unsigned long long t0 = global_ns();
        bool ok = true;
        while (ld_acquire_sys(mine) < seq) {
            if (global_ns() - t0 > 4000000000ull) { atomicOr(flag, 4); ok = false; break; }
        }
        const volatile double* src = reinterpret_cast<const volatile double*>(reinterpret_cast<double*>(P.p[me] + 16) + ((size_t)par * 8 + q) * 8);
        for (int i = 0; i < nv; i++) out[q * nv + i] = ok ? src[i] : NAN;
    }
}

static bool use_peer_sync(const fgb_ctx* ctx) {
    static const bool off = getenv("FGB_NCCL_BARRIER") != nullptr;
    return ctx->p2p && !off;
}

// stream-ordered barrier over all ranks: everything the peers enqueued before it has completed when it completes
static int comm_barrier(fgb_ctx* ctx) {
    if (use_peer_sync(ctx)) {
        SyncPeers P;
        for (int q = 0; q < 8; q++) P.p[q] = ctx->peer_sync[q];
        k_peer_barrier<<<1, 32, 0, ctx->stream>>>(P, ctx->rank, ctx->nranks, ++ctx->bar_seq, ctx->d_flag);
        FGB_CHECK_LAUNCH(ctx, "k_peer_barrier");
        return FGB_OK;
    }
    FGB_NCCL(ctx, g_nccl.AllGather(ctx->d_result + 60, ctx->d_gather, 1, ncclDouble, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    return FGB_OK;
}

// Map the transposition buffers of every peer (CUDA IPC over NVLink/NVSwitch) so that the FFT kernels can write the
// transposed data straight into the destination GPU: compute and transfer become one kernel (no staging, no NCCL copy).
// Falls back to the grouped ncclSend/ncclRecv all-to-all if any rank cannot map its peers (or FGB_NO_P2P is set).
static int map_peers(fgb_ctx* ctx) {
    ctx->p2p = false;
    const int P = ctx->nranks, me = ctx->rank;
    for (int q = 0; q < 8; q++) { ctx->peer_xbuf[q] = ctx->peer_sbuf[q] = ctx->peer_halo[q] = nullptr; ctx->peer_sync[q] = nullptr; }
    if (P > 8) return FGB_OK;
    struct Handles { cudaIpcMemHandle_t x, s, h, y; double ok; };
    static_assert(sizeof(Handles) % 8 == 0, "handle record must be a multiple of 8 bytes");
    Handles mine;
    memset(&mine, 0, sizeof(mine));
    bool ok = getenv("FGB_NO_P2P") == nullptr;
    if (ok) ok = cudaIpcGetMemHandle(&mine.x, ctx->xbuf) == cudaSuccess && cudaIpcGetMemHandle(&mine.s, ctx->sbuf) == cudaSuccess &&
                 cudaIpcGetMemHandle(&mine.h, ctx->halo_base) == cudaSuccess && cudaIpcGetMemHandle(&mine.y, ctx->sync_base) == cudaSuccess;
    cudaGetLastError();
    mine.ok = ok ? 1.0 : 0.0;
    Handles* d_all = nullptr;
    Handles* d_mine = nullptr;
    FGB_CUDA(ctx, cudaMalloc(&d_all, sizeof(Handles) * P));
    FGB_CUDA(ctx, cudaMalloc(&d_mine, sizeof(Handles)));
    std::vector<Handles> all(P);
    FGB_CUDA(ctx, cudaMemcpyAsync(d_mine, &mine, sizeof(Handles), cudaMemcpyHostToDevice, ctx->stream));
    FGB_NCCL(ctx, g_nccl.AllGather(d_mine, d_all, sizeof(Handles), ncclChar, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    FGB_CUDA(ctx, cudaMemcpyAsync(all.data(), d_all, sizeof(Handles) * P, cudaMemcpyDeviceToHost, ctx->stream));
    FGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int q = 0; q < P; q++) ok = ok && all[q].ok == 1.0;
    if (ok) {
        for (int q = 0; q < P && ok; q++) {
            if (q == me) {
                ctx->peer_xbuf[q] = ctx->xbuf; ctx->peer_sbuf[q] = ctx->sbuf; ctx->peer_halo[q] = ctx->halo_base; ctx->peer_sync[q] = ctx->sync_base;
                continue;
            }
            void *px = nullptr, *psb = nullptr, *ph = nullptr, *py = nullptr;
            ok = cudaIpcOpenMemHandle(&px, all[q].x, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess &&
                 cudaIpcOpenMemHandle(&psb, all[q].s, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess &&
                 cudaIpcOpenMemHandle(&ph, all[q].h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess &&
                 cudaIpcOpenMemHandle(&py, all[q].y, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
            ctx->peer_xbuf[q] = (double*)px;
            ctx->peer_sbuf[q] = (double*)psb;
            ctx->peer_halo[q] = (double*)ph;
            ctx->peer_sync[q] = (unsigned long long*)py;
        }
        cudaGetLastError();
    }
    // consensus: every rank must have mapped every peer
    mine.ok = ok ? 1.0 : 0.0;
    FGB_CUDA(ctx, cudaMemcpyAsync(d_mine, &mine, sizeof(Handles), cudaMemcpyHostToDevice, ctx->stream));
    FGB_NCCL(ctx, g_nccl.AllGather(d_mine, d_all, sizeof(Handles), ncclChar, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    FGB_CUDA(ctx, cudaMemcpyAsync(all.data(), d_all, sizeof(Handles) * P, cudaMemcpyDeviceToHost, ctx->stream));
    FGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int q = 0; q < P; q++) ok = ok && all[q].ok == 1.0;
    cudaFree(d_all);
    cudaFree(d_mine);
    ctx->p2p = ok;
    return FGB_OK;
}

extern "C" int fgb_comm_unique_id(void* id128) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    int rc = load_nccl(nullptr);
    if (rc) return rc;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return fgb_fail(nullptr, FGB_ECOMM, "ncclGetUniqueId failed");
    memcpy(id128, &id, 128);
    return FGB_OK;
}

extern "C" int fgb_comm_init(fgb_ctx* ctx, const void* id128) {
    if (!ctx) return FGB_EINVAL;
    cudaSetDevice(ctx->device);
    if (ctx->nranks == 1) return FGB_OK;
    int rc = load_nccl(ctx);
    if (rc) return rc;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t comm;
    FGB_NCCL(ctx, g_nccl.CommInitRank(&comm, ctx->nranks, id, ctx->rank));
    ctx->nccl_comm = comm;
    const GridDev& g = ctx->g;
    // halo slots: a full x plane of one component in either layout
    size_t a = (size_t)g.ny * g.nzp, b = (size_t)g.ny * 2 * g.unzcs;
    ctx->halo_slot = a > b ? a : b;
    // two alternating sets of each halo buffer (an exchange may start while a neighbour still reads the previous one)
    ctx->iso_set = a * (10 + 2 * FGB_MAX_PHASES);
    FGB_CUDA(ctx, cudaMalloc(&ctx->halo_base, sizeof(double) * 2 * (6 * ctx->halo_slot + ctx->iso_set)));
    ctx->halo = ctx->halo_base;
    ctx->iso_halo = ctx->halo_base + 12 * ctx->halo_slot;
    ctx->halo_seq = ctx->iso_seq = 0;
    ctx->phi_halo_valid = false;
    FGB_CUDA(ctx, cudaMalloc(&ctx->d_gather, sizeof(double) * 64 * ctx->nranks));
    FGB_CUDA(ctx, cudaMalloc(&ctx->sync_base, FGB_SYNC_BYTES));
    FGB_CUDA(ctx, cudaMemset(ctx->sync_base, 0, FGB_SYNC_BYTES));
    ctx->bar_seq = ctx->ex_seq = 0;
    // transposition buffers are allocated once (their addresses are exported to the peers)
    // sized for the scheme's own operator and for the staggered-grid operators that get_raw_field("u") applies on every context
    // (fg:15517-15557): max(dim*nzc, udim*unzcs) complex numbers per row of the slab
    const bool stag = ctx->scheme == FGB_GAMMA_STAGGERED;
    size_t per_row = (size_t)ctx->udim * g.unzcs;
    if (!stag && (size_t)ctx->dim * g.nzc > per_row) per_row = (size_t)ctx->dim * g.nzc;
    if ((rc = ensure_xbuf(ctx, 2 * per_row * g.lnx * g.ny))) return rc;
    return map_peers(ctx);
}

int fgb_comm_free(fgb_ctx* ctx) {
    if (ctx->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    if (ctx->p2p)
        for (int q = 0; q < ctx->nranks && q < 8; q++)
            if (q != ctx->rank) {
                if (ctx->peer_xbuf[q]) cudaIpcCloseMemHandle(ctx->peer_xbuf[q]);
                if (ctx->peer_sbuf[q]) cudaIpcCloseMemHandle(ctx->peer_sbuf[q]);
                if (ctx->peer_halo[q]) cudaIpcCloseMemHandle(ctx->peer_halo[q]);
                if (ctx->peer_sync[q]) cudaIpcCloseMemHandle(ctx->peer_sync[q]);
            }
    ctx->p2p = false;
    if (ctx->sbuf) cudaFree(ctx->sbuf);
    if (ctx->xbuf) cudaFree(ctx->xbuf);
    if (ctx->halo_base) cudaFree(ctx->halo_base);
    if (ctx->sync_base) cudaFree(ctx->sync_base);
    ctx->sync_base = nullptr;
    ctx->halo_base = ctx->iso_halo = nullptr;
    if (ctx->d_gather) cudaFree(ctx->d_gather);
    ctx->sbuf = ctx->xbuf = ctx->halo = ctx->d_gather = nullptr;
    return FGB_OK;
}

// all-to-all of `ncomp` components: chunk (c, q) of `chunk` complex numbers at src + (c*P + q)*chunk goes to rank q and lands
// at dst + (c*P + me)*chunk there
static int alltoall(fgb_ctx* ctx, const double* src, double* dst, int ncomp, size_t chunk_complex) {
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    const int P = ctx->nranks, me = ctx->rank;
    const size_t cnt = 2 * chunk_complex;          // doubles
    FGB_NCCL(ctx, g_nccl.GroupStart());
    for (int c = 0; c < ncomp; c++) {
        for (int dq = 1; dq < P; dq++) {
            const int to = (me + dq) % P, from = (me - dq + P) % P;
            FGB_NCCL(ctx, g_nccl.Send(src + ((size_t)c * P + to) * cnt, cnt, ncclDouble, to, comm, ctx->stream));
            FGB_NCCL(ctx, g_nccl.Recv(dst + ((size_t)c * P + from) * cnt, cnt, ncclDouble, from, comm, ctx->stream));
        }
    }
    FGB_NCCL(ctx, g_nccl.GroupEnd());
    for (int c = 0; c < ncomp; c++)
        FGB_CUDA(ctx, cudaMemcpyAsync(dst + ((size_t)c * P + me) * cnt, src + ((size_t)c * P + me) * cnt, sizeof(double) * cnt,
                                      cudaMemcpyDeviceToDevice, ctx->stream));
    return FGB_OK;
}

// forward y -> transpose -> fused x pass -> transpose -> inverse y; `base` holds the z-transformed slab (rows of lay.nzcs complex)
int fgb_comm_fft_x(fgb_ctx* ctx, double* base, int ncomp, const FftLayout& lay, const GreenArgs* ga) {
    int rc = need_comm(ctx);
    if (rc) return rc;
    const GridDev& g = ctx->g;
    const int P = ctx->nranks, lny = g.ny / P, nzcs = lay.nzcs;
    const size_t need = 2 * (size_t)ncomp * g.lnx * g.ny * nzcs;
    struct TmpPair {
        double *s = nullptr, *x = nullptr;
        ~TmpPair() { if (s) cudaFree(s); if (x) cudaFree(x); }
    } tmp;
    double *sbuf = ctx->sbuf, *xbuf = ctx->xbuf;
    bool p2p = ctx->p2p;
    if (ctx->p2p && need > ctx->xbuf_cap) {
        if (cudaMalloc(&tmp.s, sizeof(double) * need) != cudaSuccess || cudaMalloc(&tmp.x, sizeof(double) * need) != cudaSuccess)
            return fgb_fail(ctx, FGB_ENOMEM, "cannot allocate temporary transpose buffers (%zu bytes each)", sizeof(double) * need);
        sbuf = tmp.s; xbuf = tmp.x; p2p = false;
    } else {
        if ((rc = ensure_xbuf(ctx, need))) return rc;
        sbuf = ctx->sbuf; xbuf = ctx->xbuf;
    }
    const size_t chunk = (size_t)g.lnx * lny * nzcs;                 // complex numbers per (component, peer)
    const PencilMap nat = {nzcs, g.ny, 0, (long)g.ny * nzcs, (long)g.lnx * g.ny * nzcs};
    const PencilMap stg = {nzcs, lny, (long)chunk, (long)lny * nzcs, (long)P * (long)chunk};
    if (p2p) {
        // fused compute + transfer: the forward y pass stores segment q of every pencil into R of rank q over NVLink,
        // the fused x pass stores segment q of its output into the staging buffer of rank q; two stream-ordered barriers.
        const int me = ctx->rank;
        PeerTable pr, psb;
        pr.n = psb.n = P;
        for (int q = 0; q < P; q++) {
            pr.p[q] = (double2*)ctx->peer_xbuf[q] + (long)me * g.lnx * lny * nzcs;     // R_q[c][me*lnx + il][jl][k]
            psb.p[q] = (double2*)ctx->peer_sbuf[q] + (long)me * (long)chunk;            // S_q[c][me][il][jl][k]
        }
        const PencilMap rmap = {nzcs, lny, 0, (long)lny * nzcs, (long)g.nx * lny * nzcs};
        const PencilMap smap = {(long)lny * nzcs, g.lnx, 0, (long)nzcs, (long)P * (long)chunk};
        {
            ProfScope ps(ctx, "fft_y_fwd_p2p");
            if ((rc = fgb_fft_strided(ctx, 1, base, ctx->xbuf, nat, rmap, g.nzc, g.lnx, ncomp, -1, &pr))) return rc;
        }
        if ((rc = comm_barrier(ctx))) return rc;
        if (!getenv("FGB_P2P_MEMCPY")) {
            // default: the fused x pass stores its output straight into the peers' staging buffers (measured faster than the
            // copy-engine variant below: 5.18 vs 5.40 ms per iteration at 2 GPUs)
            if ((rc = fgb_fft_x_green_layout(ctx, ctx->xbuf, ga, (long)lny * nzcs, g.nzc, lny, nzcs, (long)g.nx * lny * nzcs, me * lny, &smap, &psb)))
                return rc;
        } else {
            // variant: x pass in place on R, then the copy engines push chunk
            // (c, q) = R[c][q*lnx .. (q+1)*lnx) to S_q[c][me] over NVLink (contiguous on both sides)
            if ((rc = fgb_fft_x_green_layout(ctx, ctx->xbuf, ga, (long)lny * nzcs, g.nzc, lny, nzcs, (long)g.nx * lny * nzcs, me * lny))) return rc;
            ProfScope ps(ctx, "p2p_push_bwd");
            const size_t cb = sizeof(double) * 2 * chunk;
            for (int c = 0; c < ncomp; c++)
                for (int dq = 0; dq < P; dq++) {
                    const int q = (me + dq) % P;
                    FGB_CUDA(ctx, cudaMemcpyAsync(ctx->peer_sbuf[q] + 2 * ((size_t)c * P + me) * chunk, ctx->xbuf + 2 * ((size_t)c * P + q) * chunk, cb,
                                                  cudaMemcpyDeviceToDevice, ctx->stream));
                }
        }
        if ((rc = comm_barrier(ctx))) return rc;
        ProfScope ps(ctx, "fft_y_bwd");
        return fgb_fft_strided(ctx, 1, ctx->sbuf, base, stg, nat, g.nzc, g.lnx, ncomp, +1);
    }
    {
        ProfScope ps(ctx, "fft_y_fwd");
        if ((rc = fgb_fft_strided(ctx, 1, base, sbuf, nat, stg, g.nzc, g.lnx, ncomp, -1))) return rc;
    }
    {
        ProfScope ps(ctx, "alltoall_fwd");
        if ((rc = alltoall(ctx, sbuf, xbuf, ncomp, chunk))) return rc;
    }
    // y-slab layout R[c][ii][jl][k]: x pencils have stride lny*nzcs, the outer index is jl, jj = rank*lny + jl
    if ((rc = fgb_fft_x_green_layout(ctx, xbuf, ga, (long)lny * nzcs, g.nzc, lny, nzcs, (long)g.nx * lny * nzcs, ctx->rank * lny))) return rc;
    {
        ProfScope ps(ctx, "alltoall_bwd");
        if ((rc = alltoall(ctx, xbuf, sbuf, ncomp, chunk))) return rc;
    }
    {
        ProfScope ps(ctx, "fft_y_bwd");
        if ((rc = fgb_fft_strided(ctx, 1, sbuf, base, stg, nat, g.nzc, g.lnx, ncomp, +1))) return rc;
    }
    if (tmp.s) FGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // the temporary pair is freed on return
    return FGB_OK;
}

// Direct halo push over peer memory: one kernel copies this rank's boundary planes straight into the neighbours' halo buffers
// (16-byte accesses), followed by one stream-ordered barrier -- instead of ~20 grouped ncclSend/ncclRecv operations.
struct HaloJobs {
    int n;
    const double* src[20];
    double* dst[20];
};
__global__ void __launch_bounds__(256) k_halo_push(HaloJobs J, size_t n2) {
    const double2* s = reinterpret_cast<const double2*>(J.src[blockIdx.y]);
    double2* d = reinterpret_cast<double2*>(J.dst[blockIdx.y]);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) d[i] = s[i];
}
static int halo_push(fgb_ctx* ctx, const HaloJobs& J, size_t plane_elems) {
    if (J.n == 0) return comm_barrier(ctx);
    const size_t n2 = plane_elems / 2;
    unsigned gx = (unsigned)((n2 + 255) / 256);
    if (gx > 64) gx = 64;
    dim3 grid(gx, J.n);
    k_halo_push<<<grid, 256, 0, ctx->stream>>>(J, n2);
    FGB_CHECK_LAUNCH(ctx, "k_halo_push");
    return comm_barrier(ctx);
}

// neighbour planes: lo slot s <- plane lnx-1 of lo_src[s] on the left rank, hi slot s <- plane 0 of hi_src[s] on the right rank
static int halo_exchange(fgb_ctx* ctx, const double* const* lo_src, int nlo, const double* const* hi_src, int nhi, size_t plane_elems,
                         size_t comp_stride_unused) {
    (void)comp_stride_unused;
    int rc = need_comm(ctx);
    if (rc) return rc;
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    const int P = ctx->nranks, me = ctx->rank;
    const int left = (me - 1 + P) % P, right = (me + 1) % P;
    const GridDev& g = ctx->g;
    // alternate between the two halo sets
    const bool push = ctx->p2p && (plane_elems % 2) == 0 && !getenv("FGB_HALO_NCCL");
    const size_t set_off = push ? (size_t)(ctx->halo_seq++ & 1u) * 6 * ctx->halo_slot : 0;
    ctx->halo = ctx->halo_base + set_off;
    double* lo = ctx->halo;
    double* hi = ctx->halo + 3 * ctx->halo_slot;
    ProfScope ps(ctx, "halo_exchange");
    if (push) {
        HaloJobs J;
        J.n = 0;
        for (int s = 0; s < nhi; s++) { J.src[J.n] = hi_src[s]; J.dst[J.n++] = ctx->peer_halo[left] + set_off + (3 + s) * ctx->halo_slot; }
        for (int s = 0; s < nlo; s++) {
            J.src[J.n] = lo_src[s] + (size_t)(g.lnx - 1) * plane_elems;
            J.dst[J.n++] = ctx->peer_halo[right] + set_off + s * ctx->halo_slot;
        }
        return halo_push(ctx, J, plane_elems);
    }
    FGB_NCCL(ctx, g_nccl.GroupStart());
    // my first planes go to the left rank (its hi halo); my last planes go to the right rank (its lo halo)
    for (int s = 0; s < nhi; s++) FGB_NCCL(ctx, g_nccl.Send(hi_src[s], plane_elems, ncclDouble, left, comm, ctx->stream));
    for (int s = 0; s < nlo; s++)
        FGB_NCCL(ctx, g_nccl.Send(lo_src[s] + (size_t)(g.lnx - 1) * plane_elems, plane_elems, ncclDouble, right, comm, ctx->stream));
    for (int s = 0; s < nhi; s++) FGB_NCCL(ctx, g_nccl.Recv(hi + s * ctx->halo_slot, plane_elems, ncclDouble, right, comm, ctx->stream));
    for (int s = 0; s < nlo; s++) FGB_NCCL(ctx, g_nccl.Recv(lo + s * ctx->halo_slot, plane_elems, ncclDouble, left, comm, ctx->stream));
    FGB_NCCL(ctx, g_nccl.GroupEnd());
    return FGB_OK;
}

// k_div needs tau_0 at i-1 and the two shear components at i+1 (fg:18868-18901, hyper fg:19026-19064)
int fgb_comm_halo_tau(fgb_ctx* ctx, const double* tau) {
    const GridDev& g = ctx->g;
    const size_t pe = (size_t)g.ny * g.nzp;
    const double* lo[1] = {tau};
    if (ctx->dim == 3) return halo_exchange(ctx, lo, 1, nullptr, 0, pe, 0);
    const int c1x = (ctx->dim == 6) ? 5 : 8, c2x = (ctx->dim == 6) ? 4 : 7;
    const double* hi[2] = {tau + (size_t)c1x * g.plane, tau + (size_t)c2x * g.plane};
    return halo_exchange(ctx, lo, 1, hi, 2, pe, 0);
}

// k_heat_march needs tau_0 at i-1 (fg:18924-18962): component 0 of r, p_old and of the effective conductivity on the left rank's last plane
int fgb_comm_halo_heat(fgb_ctx* ctx, const double* r, const double* p_old) {
    const GridDev& g = ctx->g;
    const size_t pe = (size_t)g.ny * g.nzp;
    const double* lo[3] = {r ? r : p_old, p_old, ctx->heatK};
    return halo_exchange(ctx, lo, 3, nullptr, 0, pe, 0);
}

// k_eps needs u_0 at i+1 and u_0..2 at i-1 (fg:18632-18654)
int fgb_comm_halo_u(fgb_ctx* ctx) {
    const GridDev& g = ctx->g;
    const size_t pe = (size_t)g.ny * 2 * g.unzcs;
    const double* u = ctx->ubuf;
    const double* hi[1] = {u};
    if (ctx->dim == 3) return halo_exchange(ctx, nullptr, 0, hi, 1, pe, 0);
    const double* lo[3] = {u, u + g.uplane, u + 2 * g.uplane};
    return halo_exchange(ctx, lo, 3, hi, 1, pe, 0);
}

// combine per-rank partial results in rank order (op 0 sum, 1 min, 2 max); vals holds this rank's values on entry
int fgb_allreduce_host(fgb_ctx* ctx, double* vals, int n, int op) {
    int rc = need_comm(ctx);
    if (rc) return rc;
    if (n > 64) return fgb_fail(ctx, FGB_EINVAL, "too many reduction values");
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    // d_result still holds this rank's n values (fgb_reduce_finish wrote them); gather all ranks' vectors
    if (use_peer_sync(ctx) && n <= 8) {
        if ((rc = fgb_allgather_dev(ctx, n))) return rc;
    } else FGB_NCCL(ctx, g_nccl.AllGather(ctx->d_result, ctx->d_gather, (size_t)n, ncclDouble, comm, ctx->stream));
    std::vector<double> all((size_t)n * ctx->nranks);
    FGB_CUDA(ctx, cudaMemcpyAsync(all.data(), ctx->d_gather, sizeof(double) * all.size(), cudaMemcpyDeviceToHost, ctx->stream));
    FGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n; i++) {
        double acc = all[i];
        for (int r = 1; r < ctx->nranks; r++) {
            const double x = all[(size_t)r * n + i];
            acc = (op == 0) ? acc + x : (op == 1 ? (x < acc ? x : acc) : (x > acc ? x : acc));
        }
        vals[i] = acc;
    }
    return FGB_OK;
}

int fgb_allgather_dev(fgb_ctx* ctx, int n) {
    int rc = need_comm(ctx);
    if (rc) return rc;
    if (use_peer_sync(ctx) && n <= 8) {
        SyncPeers P;
        for (int q = 0; q < 8; q++) P.p[q] = ctx->peer_sync[q];
        k_peer_allgather<<<1, 32, 0, ctx->stream>>>(P, ctx->rank, ctx->nranks, ++ctx->ex_seq, ctx->d_result, n, ctx->d_gather, ctx->d_flag);
        FGB_CHECK_LAUNCH(ctx, "k_peer_allgather");
        return FGB_OK;
    }
    FGB_NCCL(ctx, g_nccl.AllGather(ctx->d_result, ctx->d_gather, (size_t)n, ncclDouble, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    return FGB_OK;
}

// planes for the fused isotropic sweep (fused.cu): see fgb_internal.h for the slot order
int fgb_comm_halo_iso(fgb_ctx* ctx, const double* r, const double* p_old) {
    int rc = need_comm(ctx);
    if (rc) return rc;
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    const GridDev& g = ctx->g;
    const int P = ctx->nranks, me = ctx->rank, NPH = ctx->nphases;
    const int left = (me - 1 + P) % P, right = (me + 1) % P;
    const size_t pe = (size_t)g.ny * g.nzp;
    const bool push = ctx->p2p && !getenv("FGB_HALO_NCCL");
    const size_t iso_off = 12 * ctx->halo_slot + (push ? (size_t)(ctx->iso_seq++ & 1u) * ctx->iso_set : 0);
    ctx->iso_halo = ctx->halo_base + iso_off;
    double* H = ctx->iso_halo;
    double* r_lo = H;            double* p_lo = H + 3 * pe;
    double* r_hi = H + 6 * pe;   double* p_hi = H + 8 * pe;
    double* phi_lo = H + 10 * pe; double* phi_hi = H + (10 + FGB_MAX_PHASES) * pe;
    const size_t last = (size_t)(g.lnx - 1) * pe;
    static const int lo_c[3] = {0, 1, 2}, hi_c[2] = {5, 4};
    ProfScope ps(ctx, "halo_exchange");
    // the phase fractions do not change between exchanges: they are sent once, into both sets
    const bool send_phi = !ctx->phi_halo_valid;
    if (push) {
        HaloJobs J;
        J.n = 0;
        double* L = ctx->peer_halo[left] + iso_off;       // the left rank's current set: my first planes are its hi halo
        double* R = ctx->peer_halo[right] + iso_off;      // the right rank's current set: my last planes are its lo halo
        for (int s = 0; s < 2; s++) {
            if (r) { J.src[J.n] = r + (size_t)hi_c[s] * g.plane; J.dst[J.n++] = L + (6 + s) * pe; }
            J.src[J.n] = p_old + (size_t)hi_c[s] * g.plane; J.dst[J.n++] = L + (8 + s) * pe;
        }
        for (int s = 0; s < 3; s++) {
            if (r) { J.src[J.n] = r + (size_t)lo_c[s] * g.plane + last; J.dst[J.n++] = R + s * pe; }
            J.src[J.n] = p_old + (size_t)lo_c[s] * g.plane + last; J.dst[J.n++] = R + (3 + s) * pe;
        }
        if (send_phi) {
            // separate launch (job table size): both sets of both neighbours
            HaloJobs Jp;
            Jp.n = 0;
            for (int set = 0; set < 2; set++) {
                double* Ls = ctx->peer_halo[left] + 12 * ctx->halo_slot + (size_t)set * ctx->iso_set;
                double* Rs = ctx->peer_halo[right] + 12 * ctx->halo_slot + (size_t)set * ctx->iso_set;
                for (int q = 0; q < NPH; q++) {
                    Jp.src[Jp.n] = ctx->phi[q]; Jp.dst[Jp.n++] = Ls + (10 + FGB_MAX_PHASES + q) * pe;
                    Jp.src[Jp.n] = ctx->phi[q] + last; Jp.dst[Jp.n++] = Rs + (10 + q) * pe;
                }
            }
            if (Jp.n > 20) return fgb_fail(ctx, FGB_EUNSUPPORTED, "too many phases for the halo push");
            k_halo_push<<<dim3(64, Jp.n), 256, 0, ctx->stream>>>(Jp, pe / 2);
            FGB_CHECK_LAUNCH(ctx, "k_halo_push");
            ctx->phi_halo_valid = true;
        }
        return halo_push(ctx, J, pe);
    }
    FGB_NCCL(ctx, g_nccl.GroupStart());
    // sends: first planes (components 5,4 [+phi]) to the left rank, last planes (components 0,1,2 [+phi]) to the right rank
    for (int s = 0; s < 2; s++) {
        if (r) FGB_NCCL(ctx, g_nccl.Send(r + (size_t)hi_c[s] * g.plane, pe, ncclDouble, left, comm, ctx->stream));
        FGB_NCCL(ctx, g_nccl.Send(p_old + (size_t)hi_c[s] * g.plane, pe, ncclDouble, left, comm, ctx->stream));
    }
    if (!ctx->phi_halo_valid)
        for (int q = 0; q < NPH; q++) FGB_NCCL(ctx, g_nccl.Send(ctx->phi[q], pe, ncclDouble, left, comm, ctx->stream));
    for (int s = 0; s < 3; s++) {
        if (r) FGB_NCCL(ctx, g_nccl.Send(r + (size_t)lo_c[s] * g.plane + last, pe, ncclDouble, right, comm, ctx->stream));
        FGB_NCCL(ctx, g_nccl.Send(p_old + (size_t)lo_c[s] * g.plane + last, pe, ncclDouble, right, comm, ctx->stream));
    }
    if (!ctx->phi_halo_valid)
        for (int q = 0; q < NPH; q++) FGB_NCCL(ctx, g_nccl.Send(ctx->phi[q] + last, pe, ncclDouble, right, comm, ctx->stream));
    // receives in the matching order: from the right rank its first planes (my hi halo), from the left rank its last planes (my lo halo)
    for (int s = 0; s < 2; s++) {
        if (r) FGB_NCCL(ctx, g_nccl.Recv(r_hi + s * pe, pe, ncclDouble, right, comm, ctx->stream));
        FGB_NCCL(ctx, g_nccl.Recv(p_hi + s * pe, pe, ncclDouble, right, comm, ctx->stream));
    }
    if (!ctx->phi_halo_valid)
        for (int q = 0; q < NPH; q++) FGB_NCCL(ctx, g_nccl.Recv(phi_hi + q * pe, pe, ncclDouble, right, comm, ctx->stream));
    for (int s = 0; s < 3; s++) {
        if (r) FGB_NCCL(ctx, g_nccl.Recv(r_lo + s * pe, pe, ncclDouble, left, comm, ctx->stream));
        FGB_NCCL(ctx, g_nccl.Recv(p_lo + s * pe, pe, ncclDouble, left, comm, ctx->stream));
    }
    if (!ctx->phi_halo_valid)
        for (int q = 0; q < NPH; q++) FGB_NCCL(ctx, g_nccl.Recv(phi_lo + q * pe, pe, ncclDouble, left, comm, ctx->stream));
    FGB_NCCL(ctx, g_nccl.GroupEnd());
    ctx->phi_halo_valid = true;
    return FGB_OK;
}
