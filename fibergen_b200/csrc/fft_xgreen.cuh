// Fused x pass (forward x, Green operator, inverse x: one HBM round trip) for power-of-two nx, and the launcher that picks a
// kernel for (components, operator kind, nx).  Included by the per-operator translation units fft_xg*.cu so that the
// instantiations compile in parallel.
#pragma once
#include "fft_generic.cuh"
#include "fft_pow2.cuh"
#include "fft_pow2_3.cuh"
#include "fft_launch.cuh"
#include <cstdlib>
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

using p2::Max;

// x pass fused with the Green operator (power-of-two nx).  One shared-memory exchange per direction:
//   forward : threads n2: R1-point FFT over n1, twiddle W_N^(n2 k1) | exchange | threads k1: R2-point FFT over n2 -> X[k1 + R1 k2]
//   operator: every thread owns all NC components at its R2 frequencies (registers)
//   inverse : threads k1: R2-point inverse FFT over k2, twiddle conj W_N^(k1 na) | exchange | threads na: R1-point inverse FFT over k1
//             -> x[na + R2 nb], the same distribution the forward pass loaded, stored straight back to HBM.
template <int N, int R1, int R2, int NC, int KIND, int T, int ASYNC>
__global__ void __launch_bounds__(Max<R1, R2>::v* T) k_fftx_green_p2(double2* __restrict__ base, const double2* __restrict__ tw, GreenDev G,
                                                                     long estride, int ninner, long ostride, long cstride, int jbase,
                                                                     PencilMap xo, PeerTable pt) {
    constexpr int TPP = Max<R1, R2>::v;
    constexpr int NT = TPP * T;
    extern __shared__ double2 smem[];
    double2* tw_s = smem;                 // N
    double2* S = smem + N;                // NC * N * T exchange space, component c at S + c*N*T
    const int tid = threadIdx.x;
    for (int i = tid; i < N; i += NT) tw_s[i] = tw[i];
    const int t = tid % T, s = tid / T;
    const int inner = blockIdx.x * T + t;
    const bool valid = inner < ninner;
    double2* g = base + (long)blockIdx.y * ostride + inner;
    __syncthreads();

    // ---- forward pass 1 (per component; one component in registers at a time).  All NC*R1 loads of a thread are issued up front
    // as 16-byte asynchronous copies into the exchange buffer (element x of component c at Sc[x*T + t], exactly where the
    // thread stores its pass-1 result for k1 = x / R2), one commit group per component, so the whole tile is in flight while
    // the first component is being transformed.
    if (s < R2) {
        if (ASYNC) {
#pragma unroll
            for (int c = 0; c < NC; c++) {
                double2* Sc = S + (size_t)c * N * T;
#pragma unroll
                for (int n1 = 0; n1 < R1; n1++) {
                    double2* d = Sc + ((R2 * n1 + s) * T + t);
                    if (valid) {
                        const unsigned sa = (unsigned)__cvta_generic_to_shared(d);
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(g + c * cstride + (long)(R2 * n1 + s) * estride) : "memory");
                    } else {
                        *d = make_double2(0, 0);
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
        }
#pragma unroll(ASYNC == 2 ? 1 : NC)
        for (int c = 0; c < NC; c++) {
            double2* Sc = S + (size_t)c * N * T;
            double2 v[R1];
            if (ASYNC) {
                if (c == 0) asm volatile("cp.async.wait_group %0;" ::"n"(NC - 1) : "memory");
                else if (c == 1) asm volatile("cp.async.wait_group %0;" ::"n"(NC > 2 ? NC - 2 : 0) : "memory");
                else asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
                for (int n1 = 0; n1 < R1; n1++) v[n1] = Sc[(R2 * n1 + s) * T + t];
            } else {
#pragma unroll
                for (int n1 = 0; n1 < R1; n1++) v[n1] = valid ? g[c * cstride + (long)(R2 * n1 + s) * estride] : make_double2(0, 0);
            }
            p2::pass1<R1, R2, -1>(v, s, tw_s);
#pragma unroll
            for (int k1 = 0; k1 < R1; k1++) Sc[(k1 * R2 + s) * T + t] = v[k1];
        }
    }
    __syncthreads();
    // ---- forward pass 2, Green operator and inverse pass 1 on registers
    double2 w[NC][R2];
    if (s < R1) {
#pragma unroll
        for (int c = 0; c < NC; c++) {
            const double2* Sc = S + (size_t)c * N * T;
#pragma unroll
            for (int n2 = 0; n2 < R2; n2++) w[c][n2] = Sc[(s * R2 + n2) * T + t];
            p2::RegFFT<R2, -1>::run(w[c]);
        }
        const int jj = jbase + blockIdx.y;
        const int kk = inner;
        if (valid) {
#pragma unroll
            for (int k2 = 0; k2 < R2; k2++) {
                const int ii = s + R1 * k2;
                double2 f[NC];
#pragma unroll
                for (int c = 0; c < NC; c++) f[c] = w[c][k2];
                if (ii == 0 && jj == 0 && kk == 0) {
#pragma unroll
                    for (int c = 0; c < NC; c++) f[c] = make_double2(G.dc[c], 0.0);
                } else {
                    green_apply<KIND>(G, ii, jj, kk, f);
                }
#pragma unroll
                for (int c = 0; c < NC; c++) w[c][k2] = f[c];
            }
        }
#pragma unroll
        for (int c = 0; c < NC; c++) {
            p2::RegFFT<R2, +1>::run(w[c]);          // inverse over k2 -> Y[k1][na], na = 0..R2-1
#pragma unroll
            for (int na = 1; na < R2; na++) {
                double2 tws = tw_s[s * na];
                tws.y = -tws.y;
                w[c][na] = p2::pmul(w[c][na], tws);
            }
        }
    }
    __syncthreads();          // all exchange reads done before S is overwritten
    if (s < R1) {
#pragma unroll
        for (int c = 0; c < NC; c++) {
            double2* Sc = S + (size_t)c * N * T;
#pragma unroll
            for (int na = 0; na < R2; na++) Sc[(na * R1 + s) * T + t] = w[c][na];
        }
    }
    __syncthreads();
    // ---- inverse pass 2: thread na gathers Y[k1][na] over k1, R1-point inverse FFT, output x[na + R2*nb]
    if (s < R2) {
#pragma unroll(ASYNC == 2 ? 1 : NC)
        for (int c = 0; c < NC; c++) {
            const double2* Sc = S + (size_t)c * N * T;
            double2 v[R1];
#pragma unroll
            for (int k1 = 0; k1 < R1; k1++) v[k1] = Sc[(s * R1 + k1) * T + t];
            p2::RegFFT<R1, +1>::run(v);
            if (valid) {
                if (!pt.n && xo.seglen >= N) {
                    // unsegmented destination (single GPU): no per-element division
                    double2* b = base + c * xo.cstride + (long)blockIdx.y * xo.ostride + inner;
#pragma unroll
                    for (int nb = 0; nb < R1; nb++) b[(long)(s + R2 * nb) * xo.estride] = v[nb];
                } else {
#pragma unroll
                    for (int nb = 0; nb < R1; nb++) {
                        const int e = s + R2 * nb;
                        double2* b = pt.n ? pt.p[e / xo.seglen] : base;
                        b[c * xo.cstride + (long)blockIdx.y * xo.ostride + inner + xo.at(e)] = v[nb];
                    }
                }
            }
        }
    }
}

// fused x pass: forward (three passes, one component at a time through the two exchange buffers), Green operator on the
// NC*R3 register values of a thread, mirrored inverse back to the distribution that was loaded.
template <int R1, int R2, int R3, int NC, int KIND, int T, int MINB>
__global__ void __launch_bounds__(p3::Plan<R1, R2, R3, T>::TPP* T, MINB)
    k_fftx_green_p3(double2* __restrict__ base, const double2* __restrict__ tw, GreenDev G, long estride, int ninner, long ostride,
                    long cstride, int jbase, PencilMap xo, PeerTable pt) {
    using P = p3::Plan<R1, R2, R3, T>;
    constexpr int N = P::N, M = P::M;
    extern __shared__ double2 smem_x3[];
    double2* tw_s = smem_x3;
    double2* B1 = smem_x3 + N;
    double2* B2 = B1 + P::BUF1;
    const int tid = threadIdx.x;
    for (int i = tid; i < N; i += P::TPP * T) tw_s[i] = tw[i];
    const int t = tid % T, s = tid / T;
    const int inner = blockIdx.x * T + t;
    const bool valid = inner < ninner;
    double2* g = base + (long)blockIdx.y * ostride + inner;
    __syncthreads();

    double2 w[NC][R3];
#pragma unroll
    for (int c = 0; c < NC; c++) {
        double2 v[R1];
        if (s < M) {
#pragma unroll
            for (int n1 = 0; n1 < R1; n1++) v[n1] = valid ? g[c * cstride + (long)(M * n1 + s) * estride] : make_double2(0, 0);
        }
        P::template forward<-1>(v, w[c], s, t, B1, B2, tw_s);
    }
    if (s < R1 * R2 && valid) {
        const int jj = jbase + blockIdx.y;
        const int kk = inner;
#pragma unroll
        for (int k3 = 0; k3 < R3; k3++) {
            const int ii = s + R1 * R2 * k3;
            double2 f[NC];
#pragma unroll
            for (int c = 0; c < NC; c++) f[c] = w[c][k3];
            if (ii == 0 && jj == 0 && kk == 0) {
#pragma unroll
                for (int c = 0; c < NC; c++) f[c] = make_double2(G.dc[c], 0.0);
            } else {
                green_apply<KIND>(G, ii, jj, kk, f);
            }
#pragma unroll
            for (int c = 0; c < NC; c++) w[c][k3] = f[c];
        }
    }
    __syncthreads();          // forward pass-3 reads of B2 are done before the inverse overwrites it
#pragma unroll
    for (int c = 0; c < NC; c++) {
        double2 o[R1];
        P::inverse(w[c], o, s, t, B1, B2, tw_s);
        if (s < M && valid) {
#pragma unroll
            for (int na = 0; na < R1; na++) {
                const int e = s + M * na;
                double2* b = pt.n ? pt.p[e / xo.seglen] : base;
                b[c * xo.cstride + (long)blockIdx.y * xo.ostride + inner + xo.at(e)] = o[na];
            }
        }
    }
}

// The same pass for a CTA PAIR (thread-block cluster of 2 along the tile index) whose output goes to peer GPUs: with T lanes per
// CTA a row segment is only T*16 bytes (64 at nx = 1024, where a wider tile does not fit), and 64-byte stores reach ~430 GB/s
// over NVLink against ~690 GB/s for 128-byte ones.  The two CTAs exchange one component's output through shared memory (half of
// it with asynchronous stores into the partner's: distributed shared memory + mbarrier), and each stores HALF of the x range with the
// 2T lanes of the pair -- so every segment that crosses NVLink is 2T*16 = 128 bytes.
template <int R1, int R2, int R3, int NC, int KIND, int T>
__global__ void __launch_bounds__(p3::Plan<R1, R2, R3, T>::TPP* T, 1)
    k_fftx_green_p3c(double2* __restrict__ base, const double2* __restrict__ tw, GreenDev G, long estride, int ninner, long ostride,
                     long cstride, int jbase, PencilMap xo, PeerTable pt) {
    using P = p3::Plan<R1, R2, R3, T>;
    constexpr int N = P::N, M = P::M, NT = P::TPP * T;
    static_assert((N / 2) % M == 0, "the pass-1 distribution must not straddle the two halves of the x range");
    extern __shared__ double2 smem_x3[];
    double2* tw_s = smem_x3;
    double2* B1 = smem_x3 + N;
    double2* B2 = B1 + P::BUF1;
    double2* stage = B2 + P::BUF2;          // [N][T]
    __shared__ unsigned long long xbar;          // counts the bytes the partner CTA sends into this CTA's stage
    cg::cluster_group cl = cg::this_cluster();
    const unsigned cr = cl.block_rank();
    const int tid = threadIdx.x;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&xbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // both barriers of the pair exist before either CTA sends (no stores are in flight yet, so this release is cheap)
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    for (int i = tid; i < N; i += NT) tw_s[i] = tw[i];
    const int t = tid % T, s = tid / T;
    const int inner = blockIdx.x * T + t;
    const bool valid = inner < ninner;
    double2* g = base + (long)blockIdx.y * ostride + inner;
    __syncthreads();

    double2 w[NC][R3];
#pragma unroll
    for (int c = 0; c < NC; c++) {
        double2 v[R1];
        if (s < M) {
#pragma unroll
            for (int n1 = 0; n1 < R1; n1++) v[n1] = valid ? g[c * cstride + (long)(M * n1 + s) * estride] : make_double2(0, 0);
        }
        P::template forward<-1>(v, w[c], s, t, B1, B2, tw_s);
    }
    if (s < R1 * R2 && valid) {
        const int jj = jbase + blockIdx.y;
        const int kk = inner;
#pragma unroll
        for (int k3 = 0; k3 < R3; k3++) {
            const int ii = s + R1 * R2 * k3;
            double2 f[NC];
#pragma unroll
            for (int c = 0; c < NC; c++) f[c] = w[c][k3];
            if (ii == 0 && jj == 0 && kk == 0) {
#pragma unroll
                for (int c = 0; c < NC; c++) f[c] = make_double2(G.dc[c], 0.0);
            } else {
                green_apply<KIND>(G, ii, jj, kk, f);
            }
#pragma unroll
            for (int c = 0; c < NC; c++) w[c][k3] = f[c];
        }
    }
    __syncthreads();
    // the pair's stage: CTA h holds x indices [h N/2, (h+1) N/2) with all 2T lanes of the pair, [N/2][2T].  A thread hands its
    // R1 outputs of a component to the CTA that will store them: its own half with ordinary shared-memory stores, the partner's
    // half with st.async into the partner's shared memory, which counts the bytes on the partner's mbarrier.  Nothing on this
    // path is a release fence: a fence would also order the peer stores in flight, i.e. wait for their NVLink round trip (measured:
    // release/acquire cluster barriers around the exchange made the pass 40 % slower than the 64-byte version it replaces).
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&xbar);
    const unsigned partner = cr ^ 1u;
    unsigned rstage, rbar;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rstage) : "r"((unsigned)__cvta_generic_to_shared(stage)), "r"(partner));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(bar), "r"(partner));
    const int lane2 = tid % (2 * T), eo = tid / (2 * T);
    const int inner2 = (int)(blockIdx.x - cr) * T + lane2;
    constexpr int EPI = NT / (2 * T);          // x indices per sweep of the CTA
    constexpr int NIT = (N / 2) / EPI;
    constexpr unsigned TX_BYTES = (N / 2) * T * sizeof(double2);          // what the partner sends per component
#define CL_ARRIVE_RELAXED() asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory")
#define CL_WAIT() asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory")
#pragma unroll
    for (int c = 0; c < NC; c++) {
        double2 o[R1];
        P::inverse(w[c], o, s, t, B1, B2, tw_s);
        if (c) CL_WAIT();          // both CTAs have read the previous component out of their stages
        if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(TX_BYTES) : "memory");
        if (s < M) {
#pragma unroll
            for (int na = 0; na < R1; na++) {
                // x index e = s + M na with s < M and M | N/2: the owner CTA and the slot offset are compile-time per na
                const unsigned slot = (unsigned)((s + (M * na) % (N / 2)) * (2 * T) + (int)cr * T + t);
                if ((unsigned)((M * na) / (N / 2)) == cr) {
                    stage[slot] = o[na];
                } else {
                    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];" ::"r"(
                                     rstage + slot * (unsigned)sizeof(double2)),
                                 "l"(__double_as_longlong(o[na].x)), "l"(__double_as_longlong(o[na].y)), "r"(rbar)
                                 : "memory");
                }
            }
        }
        __syncthreads();          // this CTA's own half
        {                         // the partner's half has landed (phase parity = component parity)
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "XBAR_WAIT:\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                "@p bra XBAR_DONE;\n"
                "bra XBAR_WAIT;\n"
                "XBAR_DONE:\n"
                "}\n" ::"r"(bar),
                "r"((unsigned)(c & 1))
                : "memory");
        }
        if (inner2 < ninner) {
#pragma unroll
            for (int h = 0; h < NIT; h += 4) {
                double2 q[4];
#pragma unroll
                for (int u = 0; u < 4; u++) q[u] = stage[(size_t)(eo + (h + u) * EPI) * (2 * T) + lane2];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int e = (int)cr * (N / 2) + eo + (h + u) * EPI;
                    double2* b = pt.n ? pt.p[e / xo.seglen] : base;
                    b[c * xo.cstride + (long)blockIdx.y * xo.ostride + inner2 + xo.at(e)] = q[u];
                }
            }
        }
        CL_ARRIVE_RELAXED();          // "done reading my stage": carries no data
    }
    CL_WAIT();          // no CTA leaves while its partner may still write into its stage
#undef CL_ARRIVE_RELAXED
#undef CL_WAIT
}

// ---- x with Green operator ---------------------------------------------------------------------------------
template <int N, int R1, int R2, int NC, int KIND, int T>
static int launch_xg_p2(fgb_ctx* ctx, double2* base, const GreenDev& G, long estride, int ninner, int nouter, long ostride, long cstride,
                        int jbase, const PencilMap& xo, const PeerTable& pt) {
    constexpr int NT = Max<R1, R2>::v * T;
    const size_t smem = (size_t)(N + (size_t)NC * N * T) * sizeof(double2);
    if (smem > ctx->smem_optin) return -1;
    dim3 grid((ninner + T - 1) / T, nouter, 1);
    // default: asynchronous tile prefetch, component loops of the first and last pass not unrolled (smaller code, fewer
    // instruction-cache misses); FGB_XG_UNROLLED / FGB_XG_NOASYNC select the older variants for comparison
    static const bool no_async = getenv("FGB_XG_NOASYNC") != nullptr;
    static const bool rolled = getenv("FGB_XG_UNROLLED") == nullptr && !no_async;
    if (rolled) {
        FGB_CUDA(ctx, set_smem(k_fftx_green_p2<N, R1, R2, NC, KIND, T, 2>, smem));
        k_fftx_green_p2<N, R1, R2, NC, KIND, T, 2><<<grid, NT, smem, ctx->stream>>>(base, ctx->plan[0].tw, G, estride, ninner, ostride, cstride, jbase, xo, pt);
    } else if (no_async) {
        FGB_CUDA(ctx, set_smem(k_fftx_green_p2<N, R1, R2, NC, KIND, T, 0>, smem));
        k_fftx_green_p2<N, R1, R2, NC, KIND, T, 0><<<grid, NT, smem, ctx->stream>>>(base, ctx->plan[0].tw, G, estride, ninner, ostride, cstride, jbase, xo, pt);
    } else {
        FGB_CUDA(ctx, set_smem(k_fftx_green_p2<N, R1, R2, NC, KIND, T, 1>, smem));
        k_fftx_green_p2<N, R1, R2, NC, KIND, T, 1><<<grid, NT, smem, ctx->stream>>>(base, ctx->plan[0].tw, G, estride, ninner, ostride, cstride, jbase, xo, pt);
    }
    FGB_CHECK_LAUNCH(ctx, "k_fftx_green_p2");
    return FGB_OK;
}

template <int R1, int R2, int R3, int NC, int KIND, int T, int MINB>
static int launch_xg_p3(fgb_ctx* ctx, double2* base, const GreenDev& G, long estride, int ninner, int nouter, long ostride, long cstride,
                        int jbase, const PencilMap& xo, const PeerTable& pt) {
    using P = p3::Plan<R1, R2, R3, T>;
    constexpr int NT = P::TPP * T;
    const size_t smem = (size_t)(P::N + P::SMEM_ELEMS) * sizeof(double2);
    if (smem > ctx->smem_optin) return -1;
    dim3 grid((ninner + T - 1) / T, nouter, 1);
    FGB_CUDA(ctx, set_smem(k_fftx_green_p3<R1, R2, R3, NC, KIND, T, MINB>, smem));
    k_fftx_green_p3<R1, R2, R3, NC, KIND, T, MINB><<<grid, NT, smem, ctx->stream>>>(base, ctx->plan[0].tw, G, estride, ninner, ostride, cstride,
                                                                                   jbase, xo, pt);
    FGB_CHECK_LAUNCH(ctx, "k_fftx_green_p3");
    return FGB_OK;
}

template <int R1, int R2, int R3, int NC, int KIND, int T>
static int launch_xg_p3c(fgb_ctx* ctx, double2* base, const GreenDev& G, long estride, int ninner, int nouter, long ostride, long cstride,
                         int jbase, const PencilMap& xo, const PeerTable& pt) {
    using P = p3::Plan<R1, R2, R3, T>;
    constexpr int NT = P::TPP * T;
    const size_t smem = (size_t)(P::N + P::SMEM_ELEMS + (size_t)P::N * T) * sizeof(double2);
    if (smem > ctx->smem_optin) return -1;
    auto kern = k_fftx_green_p3c<R1, R2, R3, NC, KIND, T>;
    FGB_CUDA(ctx, set_smem(kern, smem));
    const unsigned tiles = (unsigned)((ninner + T - 1) / T);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((tiles + 1) / 2 * 2, nouter, 1);          // whole pairs; a CTA past the last tile only serves the barriers
    cfg.blockDim = dim3(NT, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    FGB_CUDA(ctx, cudaLaunchKernelEx(&cfg, kern, base, (const double2*)ctx->plan[0].tw, G, estride, ninner, ostride, cstride, jbase, xo, pt));
    FGB_CHECK_LAUNCH(ctx, "k_fftx_green_p3c");
    return FGB_OK;
}

template <int NC, int KIND>
static int launch_x_green(fgb_ctx* ctx, double2* base, const GreenDev& G, long estride, int ninner, int nouter, long ostride, long cstride,
                          int jbase, const PencilMap& xo, const PeerTable& pt) {
    const int nx = ctx->g.nx;
    int rc = -1;
    // register budget: NC*R2 complex per thread -> the fast path covers NC <= 3 (staggered / heat); larger tensors use the generic kernel
    // three-pass kernels: NC*R3 complex values per thread, so every tensor rank stays in registers.  They carry nx = 512 / 1024 for
    // all operators and the 6- and 9-component (collocated) operators at every power of two; the 1- and 3-component operators at
    // nx <= 256 are faster with the two-pass kernel below.
    static const bool no_p3 = getenv("FGB_NO_P3") != nullptr;
    static const bool xg_p3 = getenv("FGB_XG_P3") != nullptr;
    constexpr int MB = (NC <= 3 ? 2 : 1);
    if (!no_p3) {
#define XG3(R1, R2, R3, T) rc = launch_xg_p3<R1, R2, R3, NC, KIND, T, MB>(ctx, base, G, estride, ninner, nouter, ostride, cstride, jbase, xo, pt)
        // stores into peer memory (slab partition): wider tiles so that a row segment is 128 bytes (nx = 512) or 64 bytes (nx = 1024)
        // instead of 64 / 32; over NVLink the segment size decides the achieved bandwidth (4-lane tiles reached ~430 GB/s at 8 GPUs).
        // FGB_XG_NARROW_PEER selects the single-GPU tile shapes for comparison.
        static const bool narrow = getenv("FGB_XG_NARROW_PEER") != nullptr;
        if constexpr (NC <= 3) {
            static const bool force_wide = getenv("FGB_XG_FORCE_PEER_TILES") != nullptr;          // profiling on one GPU
            if ((pt.n > 0 || force_wide) && !narrow) {
                if (nx == 512) rc = launch_xg_p3<8, 8, 8, NC, KIND, 8, 1>(ctx, base, G, estride, ninner, nouter, ostride, cstride, jbase, xo, pt);
                else if (nx == 1024) {
                    // CTA pairs exchanging through distributed shared memory, 128-byte segments over NVLink.  The exchange costs
                    // ~40 % more SM time per tile, so it pays only when nearly everything leaves the GPU: 1024^3 on 8 GPUs 5.83 vs
                    // 6.68 ms (default from 8 ranks on), on 2 GPUs 21.6 vs 15.5 ms.  FGB_XG_CLUSTER / FGB_XG_NO_CLUSTER force it.
                    static const bool use_cl = getenv("FGB_XG_CLUSTER") != nullptr, no_cl = getenv("FGB_XG_NO_CLUSTER") != nullptr;
                    if ((use_cl || pt.n >= 8) && !no_cl)
                        rc = launch_xg_p3c<16, 8, 8, NC, KIND, 4>(ctx, base, G, estride, ninner, nouter, ostride, cstride, jbase, xo, pt);
                    if (rc == -1) rc = launch_xg_p3<16, 8, 8, NC, KIND, 4, 1>(ctx, base, G, estride, ninner, nouter, ostride, cstride, jbase, xo, pt);
                }
            }
        }
        if (rc != -1) return rc;
        if (nx == 512) XG3(8, 8, 8, 4);
        else if (nx == 1024) {
            // 4 lanes, one CTA per SM: 0.900 ms against 0.947 ms for 2 lanes / 2 CTAs (1024x128x256, one GPU)
            if constexpr (NC <= 3) rc = launch_xg_p3<16, 8, 8, NC, KIND, 4, 1>(ctx, base, G, estride, ninner, nouter, ostride, cstride, jbase, xo, pt);
            else XG3(16, 8, 8, 2);
        }
        else if (NC > 3 || xg_p3) {
            static const bool p3_t8 = getenv("FGB_XG_P3_T8") != nullptr;          // A/B: 8-lane three-pass tile at nx = 256
            if (nx == 64) XG3(4, 4, 4, 8);
            else if (nx == 128) XG3(8, 4, 4, 8);
            else if (nx == 256 && p3_t8 && NC <= 3) XG3(8, 8, 4, 8);
            else if (nx == 256) XG3(8, 8, 4, 4);
        }
#undef XG3
        if (rc != -1) return rc;
    }
    if constexpr (NC <= 3) {
        static const bool t4 = getenv("FGB_XG_T4") != nullptr;                    // A/B: 4-lane two-pass tile (4 CTAs per SM)
        if (t4 && nx == 256) rc = launch_xg_p2<256, 16, 16, NC, KIND, 4>(ctx, base, G, estride, ninner, nouter, ostride, cstride, jbase, xo, pt);
        if (rc != -1) return rc;
        switch (nx) {
            case 64: rc = launch_xg_p2<64, 8, 8, NC, KIND, 8>(ctx, base, G, estride, ninner, nouter, ostride, cstride, jbase, xo, pt); break;
            case 128: rc = launch_xg_p2<128, 16, 8, NC, KIND, 8>(ctx, base, G, estride, ninner, nouter, ostride, cstride, jbase, xo, pt); break;
            case 256: rc = launch_xg_p2<256, 16, 16, NC, KIND, 8>(ctx, base, G, estride, ninner, nouter, ostride, cstride, jbase, xo, pt); break;
            case 512: rc = launch_xg_p2<512, 32, 16, NC, KIND, 4>(ctx, base, G, estride, ninner, nouter, ostride, cstride, jbase, xo, pt); break;
            case 1024: if constexpr (NC == 1) rc = launch_xg_p2<1024, 32, 32, NC, KIND, 2>(ctx, base, G, estride, ninner, nouter, ostride, cstride, jbase, xo, pt); break;
        }
        if (rc != -1) return rc;
    }
    int T = pick_T(ctx, nx, NC + 1, 0, 8);
    if (!T) return fgb_fail(ctx, FGB_EUNSUPPORTED, "nx=%d with %d components does not fit shared memory", nx, NC);
    const size_t smem = (size_t)(NC + 1) * nx * T * sizeof(double2);
    dim3 grid((ninner + T - 1) / T, nouter, 1);
    FGB_CUDA(ctx, set_smem(k_fft_x_green<NC, KIND>, smem));
    k_fft_x_green<NC, KIND><<<grid, 256, smem, ctx->stream>>>(base, ctx->plan[0], G, estride, ninner, ostride, cstride, T, jbase, xo, pt);
    FGB_CHECK_LAUNCH(ctx, "k_fft_x_green");
    return FGB_OK;
}

