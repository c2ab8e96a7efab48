// 3-D real<->complex FFT passes and the fused x pass with the Green operator: plans, tables, launches.
// Device code: fft_generic.cuh (any axis length: Stockham radix 4/2 + generic odd-prime stages in shared memory)
// and the power-of-two fast path below (fft_pow2.cuh: register-resident radix-16/32 passes, warp shuffles).
// Reference functions replaced: FFT3<double>::forward/backward fg:7204-7245, fftVector/fftTensor scaling
// fg:18501-18506/18548-18553, G0OperatorFourierStaggered* fg:19749-19927, GammaOperatorFourierCollocated* fg:19302-19745.
#include "fft_generic.cuh"
#include "fft_pow2.cuh"
#include "fft_pow2_3.cuh"
#include "fft_launch.cuh"
#include <cmath>
#include <complex>
#include <cstdlib>

using p2::Max;


// =================================================================================================
// power-of-two kernels
// =================================================================================================

// z pass: two real rows -> one complex pencil of length N in registers; TPP threads per pencil pair,
// exchange buffer of N(+pad) complex per pair, __syncwarp only (a pair never leaves its warp).
#define FFTZ_THREADS 128
template <int N, int R1, int R2, int FWD>
__global__ void __launch_bounds__(FFTZ_THREADS) k_fftz_p2(double* __restrict__ base, long rows, long rowstride,
                                                          const double2* __restrict__ tw, double scale) {
    constexpr int TPP = Max<R1, R2>::v;
    constexpr int PP = FFTZ_THREADS / TPP;         // pencil pairs per CTA
    constexpr int XS = R2 + 1;                     // padded exchange stride
    constexpr int NZC = N / 2 + 1;
    extern __shared__ double2 smem_z[];
    double2* tw_s = smem_z;                        // N
    const int tid = threadIdx.x;
    for (int i = tid; i < N; i += FFTZ_THREADS) tw_s[i] = tw[i];
    __syncthreads();
    const int s = tid % TPP, pp = tid / TPP;
    const long pair = (long)blockIdx.x * PP + pp;
    const long r0 = 2 * pair, r1 = r0 + 1;
    const bool v0 = r0 < rows, v1 = r1 < rows;
    double* rowA = base + r0 * rowstride;
    double* rowB = base + r1 * rowstride;
    double2* x = smem_z + N + (size_t)pp * (R1 * XS);

    if (FWD) {
        if (s < R2) {
            double2 v[R1];
#pragma unroll
            for (int n1 = 0; n1 < R1; n1++) {
                const int z = R2 * n1 + s;
                v[n1] = make_double2(v0 ? rowA[z] : 0.0, v1 ? rowB[z] : 0.0);
            }
            p2::pass1<R1, R2, -1>(v, s, tw_s);
#pragma unroll
            for (int k1 = 0; k1 < R1; k1++) x[k1 * XS + s] = v[k1];
        }
        __syncwarp();
        double2 w[R2];
        if (s < R1) {
#pragma unroll
            for (int n2 = 0; n2 < R2; n2++) w[n2] = x[s * XS + n2];
            p2::RegFFT<R2, -1>::run(w);
        } else {
#pragma unroll
            for (int n2 = 0; n2 < R2; n2++) w[n2] = make_double2(0, 0);
        }
        // split: A[k] = (Z[k] + conj Z[N-k])/2, B[k] = (Z[k] - conj Z[N-k])/(2i); Z[N-k] lives in thread (R1-k1)%R1
        const int lane = threadIdx.x & 31;
        const int partner = (lane - s) + ((R1 - s) % R1);
        const double h = 0.5 * scale;
#pragma unroll
        for (int k2 = 0; k2 < R2; k2++) {
            double2 zn;
            zn.x = __shfl_sync(0xffffffffu, w[R2 - 1 - k2].x, partner);
            zn.y = __shfl_sync(0xffffffffu, w[R2 - 1 - k2].y, partner);
            if (s == 0) zn = w[(R2 - k2) % R2];
            const int k = s + R1 * k2;
            if (s < R1 && k < NZC && v0) {
                const double2 zk = w[k2];
                reinterpret_cast<double2*>(rowA)[k] = make_double2(h * (zk.x + zn.x), h * (zk.y - zn.y));
                if (v1) reinterpret_cast<double2*>(rowB)[k] = make_double2(h * (zk.y + zn.y), -h * (zk.x - zn.x));
            }
        }
    } else {
        if (s < R2) {
            double2 v[R1];
#pragma unroll
            for (int n1 = 0; n1 < R1; n1++) {
                const int idx = R2 * n1 + s;
                const bool mirror = idx > N / 2;
                const int k = mirror ? N - idx : idx;
                double2 A = make_double2(0, 0), B = make_double2(0, 0);
                if (v0) A = reinterpret_cast<const double2*>(rowA)[k];
                if (v1) B = reinterpret_cast<const double2*>(rowB)[k];
                if (k == 0 || 2 * k == N) v[n1] = make_double2(A.x, B.x);          // c2r ignores these imaginary parts
                else if (!mirror) v[n1] = make_double2(A.x - B.y, A.y + B.x);        // A + iB
                else v[n1] = make_double2(A.x + B.y, B.x - A.y);                     // conj(A) + i conj(B)
            }
            p2::pass1<R1, R2, +1>(v, s, tw_s);
#pragma unroll
            for (int k1 = 0; k1 < R1; k1++) x[k1 * XS + s] = v[k1];
        }
        __syncwarp();
        if (s < R1) {
            double2 w[R2];
#pragma unroll
            for (int n2 = 0; n2 < R2; n2++) w[n2] = x[s * XS + n2];
            p2::RegFFT<R2, +1>::run(w);
#pragma unroll
            for (int k2 = 0; k2 < R2; k2++) {
                const int z = s + R1 * k2;
                if (v0) rowA[z] = w[k2].x;
                if (v1) rowB[z] = w[k2].y;
            }
        }
    }
}

// z pass, half-length form: ONE real row of length N = 2*N2 is the complex sequence z[m] = x[2m] + i x[2m+1] of length N2
// (16-byte loads), transformed with the two-pass register FFT and split with the twiddle W_N^k:
//   forward : X[k] = E[k] + W_N^k O[k],  E = (Z[k] + conj Z[N2-k])/2,  O = (Z[k] - conj Z[N2-k])/(2i),  X[N2] = Re Z[0] - Im Z[0]
//   backward: Z[k] = (X[k] + conj X[N2-k]) + i conj(W_N^k) (X[k] - conj X[N2-k])   (c2r: Im X[0], Im X[N2] ignored)
// Against the row-pair form above this halves the transform length (fewer flops, fewer registers) and needs no second row.
// The partner Z[N2-k] of thread s lives in thread (R1-s)%R1 of the same warp: shuffles, __syncwarp only.
template <int N2, int R1, int R2, int FWD>
__global__ void __launch_bounds__(FFTZ_THREADS) k_fftz_h(double* __restrict__ base, long rows, long rowstride,
                                                         const double2* __restrict__ tw, double scale) {
    constexpr int TPP = Max<R1, R2>::v;
    constexpr int PP = FFTZ_THREADS / TPP;         // rows per CTA
    constexpr int XS = R2 + 1;
    extern __shared__ double2 smem_zh[];
    double2* tw2_s = smem_zh;                      // W_N2^j = tw[2j]
    double2* twh_s = smem_zh + N2;                 // W_N^k, k < N2
    const int tid = threadIdx.x;
    for (int i = tid; i < N2; i += FFTZ_THREADS) { tw2_s[i] = tw[2 * i]; twh_s[i] = tw[i]; }
    __syncthreads();
    const int s = tid % TPP, pp = tid / TPP;
    const long row = (long)blockIdx.x * PP + pp;
    const bool valid = row < rows;
    double2* r2 = reinterpret_cast<double2*>(base + (valid ? row : 0) * rowstride);
    double2* x = smem_zh + 2 * N2 + (size_t)pp * (R1 * XS);
    const int lane = threadIdx.x & 31;
    const int partner = (lane - s) + ((R1 - s) % R1);

    if (FWD) {
        if (s < R2) {
            double2 v[R1];
#pragma unroll
            for (int n1 = 0; n1 < R1; n1++) v[n1] = valid ? r2[R2 * n1 + s] : make_double2(0, 0);
            p2::pass1<R1, R2, -1>(v, s, tw2_s);
#pragma unroll
            for (int k1 = 0; k1 < R1; k1++) x[k1 * XS + s] = v[k1];
        }
        __syncwarp();
        double2 w[R2];
        if (s < R1) {
#pragma unroll
            for (int n2 = 0; n2 < R2; n2++) w[n2] = x[s * XS + n2];
            p2::RegFFT<R2, -1>::run(w);
        } else {
#pragma unroll
            for (int n2 = 0; n2 < R2; n2++) w[n2] = make_double2(0, 0);
        }
        const double h = 0.5 * scale;
#pragma unroll
        for (int k2 = 0; k2 < R2; k2++) {
            double2 zn;
            zn.x = __shfl_sync(0xffffffffu, w[R2 - 1 - k2].x, partner);
            zn.y = __shfl_sync(0xffffffffu, w[R2 - 1 - k2].y, partner);
            if (s == 0) zn = w[(R2 - k2) % R2];
            const int k = s + R1 * k2;
            if (s < R1 && valid) {
                const double2 a = w[k2];
                const double ex = a.x + zn.x, ey = a.y - zn.y;          // 2E
                const double ox = a.y + zn.y, oy = -(a.x - zn.x);       // 2O = -i (a - conj zn)
                const double2 wk = twh_s[k];
                r2[k] = make_double2(h * (ex + wk.x * ox - wk.y * oy), h * (ey + wk.x * oy + wk.y * ox));
                if (k == 0) r2[N2] = make_double2(scale * (a.x - a.y), 0.0);
            }
        }
    } else {
        if (s < R2) {
            double2 v[R1];
#pragma unroll
            for (int n1 = 0; n1 < R1; n1++) {
                const int k = R2 * n1 + s;
                double2 a = make_double2(0, 0), xn = make_double2(0, 0);
                if (valid) { a = r2[k]; xn = r2[N2 - k]; }
                if (k == 0) { a.y = 0.0; xn.y = 0.0; }                   // c2r ignores the imaginary parts of X[0] and X[N/2]
                const double sx = a.x + xn.x, sy = a.y - xn.y;
                const double dx = a.x - xn.x, dy = a.y + xn.y;
                const double2 wk = twh_s[k];
                const double tx = wk.x * dx + wk.y * dy, ty = wk.x * dy - wk.y * dx;     // conj(W_N^k) * D
                v[n1] = make_double2(sx - ty, sy + tx);
            }
            p2::pass1<R1, R2, +1>(v, s, tw2_s);
#pragma unroll
            for (int k1 = 0; k1 < R1; k1++) x[k1 * XS + s] = v[k1];
        }
        __syncwarp();
        if (s < R1) {
            double2 w[R2];
#pragma unroll
            for (int n2 = 0; n2 < R2; n2++) w[n2] = x[s * XS + n2];
            p2::RegFFT<R2, +1>::run(w);
            if (valid) {
#pragma unroll
                for (int k2 = 0; k2 < R2; k2++) r2[s + R1 * k2] = w[k2];
            }
        }
    }
}

// strided pass (y; x without Green): tile of T lanes, thread (t, s); exchange buffer X[(k1*R2+n2)*T + t].
// Source and destination may differ and are addressed through PencilMaps, so the y pass of a slab-partitioned run
// writes straight into (reads straight from) the all-to-all staging layout.
template <int N, int R1, int R2, int DIR, int T>
__global__ void __launch_bounds__(Max<R1, R2>::v* T) k_ffts_p2(const double2* __restrict__ src, double2* __restrict__ dst,
                                                               const double2* __restrict__ tw, PencilMap mi, PencilMap mo, int ninner,
                                                               PeerTable pt) {
    constexpr int TPP = Max<R1, R2>::v;
    extern __shared__ double2 smem_s[];
    double2* tw_s = smem_s;          // N
    double2* X = smem_s + N;         // N * T
    const int tid = threadIdx.x;
    for (int i = tid; i < N; i += TPP * T) tw_s[i] = tw[i];
    const int t = tid % T, s = tid / T;
    const int inner = blockIdx.x * T + t;
    const bool valid = inner < ninner;
    const double2* gi = src + (long)blockIdx.z * mi.cstride + (long)blockIdx.y * mi.ostride + inner;
    const long coff = (long)blockIdx.z * mo.cstride + (long)blockIdx.y * mo.ostride + inner;
    double2 v[R1];
    if (s < R2) {
        if (mi.seglen >= N) {
#pragma unroll
            for (int n1 = 0; n1 < R1; n1++) v[n1] = valid ? gi[(long)(R2 * n1 + s) * mi.estride] : make_double2(0, 0);
        } else {
#pragma unroll
            for (int n1 = 0; n1 < R1; n1++) v[n1] = valid ? gi[mi.at(R2 * n1 + s)] : make_double2(0, 0);
        }
    }
    __syncthreads();
    if (s < R2) {
        p2::pass1<R1, R2, DIR>(v, s, tw_s);
#pragma unroll
        for (int k1 = 0; k1 < R1; k1++) X[(k1 * R2 + s) * T + t] = v[k1];
    }
    __syncthreads();
    if (s < R1) {
        double2 w[R2];
#pragma unroll
        for (int n2 = 0; n2 < R2; n2++) w[n2] = X[(s * R2 + n2) * T + t];
        p2::RegFFT<R2, DIR>::run(w);
        if (valid) {
            if (mo.seglen >= N) {
#pragma unroll
                for (int k2 = 0; k2 < R2; k2++) dst[coff + (long)(s + R1 * k2) * mo.estride] = w[k2];
            } else {
#pragma unroll
                for (int k2 = 0; k2 < R2; k2++) {
                    const int e = s + R1 * k2;
                    double2* b = pt.n ? pt.p[e / mo.seglen] : dst;
                    b[coff + mo.at(e)] = w[k2];
                }
            }
        }
    }
}

// =================================================================================================
// three-pass power-of-two kernels (N = 512, 1024): fft_pow2_3.cuh
// =================================================================================================
template <int R1, int R2, int R3, int DIR, int T>
__global__ void __launch_bounds__(p3::Plan<R1, R2, R3, T>::TPP* T) k_ffts_p3(const double2* __restrict__ src, double2* __restrict__ dst,
                                                                            const double2* __restrict__ tw, PencilMap mi, PencilMap mo,
                                                                            int ninner, PeerTable pt) {
    using P = p3::Plan<R1, R2, R3, T>;
    constexpr int N = P::N, M = P::M;
    extern __shared__ double2 smem_s3[];
    double2* tw_s = smem_s3;             // N
    double2* B = smem_s3 + N;            // P::INPLACE_ELEMS
    const int tid = threadIdx.x;
    for (int i = tid; i < N; i += P::TPP * T) tw_s[i] = tw[i];
    const int t = tid % T, s = tid / T;
    const int inner = blockIdx.x * T + t;
    const bool valid = inner < ninner;
    const double2* gi = src + (long)blockIdx.z * mi.cstride + (long)blockIdx.y * mi.ostride + inner;
    const long coff = (long)blockIdx.z * mo.cstride + (long)blockIdx.y * mo.ostride + inner;
    double2 v[R1];
    if (s < M) {
        if (mi.seglen >= N) {
#pragma unroll
            for (int n1 = 0; n1 < R1; n1++) v[n1] = valid ? gi[(long)(M * n1 + s) * mi.estride] : make_double2(0, 0);
        } else {
#pragma unroll
            for (int n1 = 0; n1 < R1; n1++) v[n1] = valid ? gi[mi.at(M * n1 + s)] : make_double2(0, 0);
        }
    }
    __syncthreads();
    double2 w[R3];
    P::template forward_inplace<DIR>(v, w, s, t, B, tw_s);
    if (s < R1 * R2 && valid) {
        if (mo.seglen >= N) {
#pragma unroll
            for (int k3 = 0; k3 < R3; k3++) dst[coff + (long)(s + R1 * R2 * k3) * mo.estride] = w[k3];
        } else {
#pragma unroll
            for (int k3 = 0; k3 < R3; k3++) {
                const int e = s + R1 * R2 * k3;
                double2* b = pt.n ? pt.p[e / mo.seglen] : dst;
                b[coff + mo.at(e)] = w[k3];
            }
        }
    }
}

// =================================================================================================
// host side: plans, tables, launches
// =================================================================================================
static void factorize(int n, FftPlanDev& P) {
    P.n = n;
    P.nstages = 0;
    int m = n;
    while (m % 4 == 0) { P.radix[P.nstages++] = 4; m /= 4; }
    while (m % 2 == 0) { P.radix[P.nstages++] = 2; m /= 2; }
    for (int p = 3; m > 1; p += 2)
        while (m % p == 0) { P.radix[P.nstages++] = p; m /= p; }
}

static double freq_index(int i, int n) {
    // fg:19393-19406: Nyquist of an even axis is treated as -n/2
    const int half = (n % 2 == 0) ? (n / 2 - 1) : (n / 2);
    return (i <= half) ? (double)i : ((double)i - (double)n);
}

static bool is_fast_pow2(int n) { return n == 64 || n == 128 || n == 256 || n == 512 || n == 1024; }

int fgb_fft_init(fgb_ctx* ctx) {
    const int dims[3] = {ctx->g.nx, ctx->g.ny, ctx->g.nz};
    const long double PI = 3.14159265358979323846264338327950288L;
    {
        double2 w32[32];
        for (int k = 0; k < 32; k++) {
            const long double ang = -2.0L * PI * (long double)k / 32.0L;
            w32[k] = make_double2((double)cosl(ang), (double)sinl(ang));
        }
        FGB_CUDA(ctx, cudaMemcpyToSymbol(c_w32, w32, sizeof(w32)));
        FGB_CUDA(ctx, fgb_w32_set_xg1(w32));
        FGB_CUDA(ctx, fgb_w32_set_xg2(w32));
        FGB_CUDA(ctx, fgb_w32_set_xg3(w32));
        FGB_CUDA(ctx, fgb_w32_set_xg4(w32));
        FGB_CUDA(ctx, fgb_w32_set_xg5(w32));
    }
    for (int a = 0; a < 3; a++) {
        const int n = dims[a];
        if (n > 16384) return fgb_fail(ctx, FGB_EUNSUPPORTED, "axis length %d too large for the shared-memory FFT", n);
        factorize(n, ctx->plan[a]);
        std::vector<double2> tw(n);
        for (int k = 0; k < n; k++) {
            const long double ang = -2.0L * PI * (long double)k / (long double)n;
            tw[k] = make_double2((double)cosl(ang), (double)sinl(ang));
        }
        FGB_CUDA(ctx, cudaMalloc(&ctx->tw_dev[a], sizeof(double2) * n));
        FGB_CUDA(ctx, cudaMemcpy(ctx->tw_dev[a], tw.data(), sizeof(double2) * n, cudaMemcpyHostToDevice));
        ctx->plan[a].tw = ctx->tw_dev[a];

        // frequency tables, computed exactly as the reference does on the host (fg:19838-19876, fg:19386)
        const double L = ctx->L[a];
        const double h = L / (2 * (double)n);
        const double xi_0 = 2 * M_PI * h / L;
        std::vector<double> kpm(n), xi(n), xi2pi(n), wtan(n), pois(n);
        std::vector<double2> kp(n), wex(n);
        const double xi2_0 = 2 * M_PI / L, wvox = L / n;
        for (int i = 0; i < n; i++) {
            const double m = freq_index(i, n);
            const double x = xi_0 * m;
            kpm[i] = std::sin(x) / h;
            const std::complex<double> z = kpm[i] * std::exp(std::complex<double>(0, x));
            kp[i] = make_double2(z.real(), z.imag());
            xi[i] = (1 / L) * m;
            // G0DivOperatorFourierHyper / GradOperatorFourierHyper fg:20159, fg:22073; Willot-R fg:19130-19153
            xi2pi[i] = xi2_0 * m;
            const double q = xi2pi[i] * wvox;
            const std::complex<double> e = 1.0 + std::polar<double>(1, q);
            wex[i] = make_double2(e.real(), e.imag());
            wtan[i] = 0.25 * std::tan(0.5 * q);
            pois[i] = ((double)n * n / (L * L)) * (std::cos((2.0 * M_PI / n) * i) - 1.0);      // poisson_solve fg:23462-23486
        }
        FGB_CUDA(ctx, cudaMalloc(&ctx->kpm_dev[a], sizeof(double) * n));
        FGB_CUDA(ctx, cudaMalloc(&ctx->kp_dev[a], sizeof(double2) * n));
        FGB_CUDA(ctx, cudaMalloc(&ctx->xi_dev[a], sizeof(double) * n));
        FGB_CUDA(ctx, cudaMemcpy(ctx->kpm_dev[a], kpm.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
        FGB_CUDA(ctx, cudaMemcpy(ctx->kp_dev[a], kp.data(), sizeof(double2) * n, cudaMemcpyHostToDevice));
        FGB_CUDA(ctx, cudaMemcpy(ctx->xi_dev[a], xi.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
        FGB_CUDA(ctx, cudaMalloc(&ctx->xi2pi_dev[a], sizeof(double) * n));
        FGB_CUDA(ctx, cudaMalloc(&ctx->wtan_dev[a], sizeof(double) * n));
        FGB_CUDA(ctx, cudaMalloc(&ctx->wex_dev[a], sizeof(double2) * n));
        FGB_CUDA(ctx, cudaMalloc(&ctx->pois_dev[a], sizeof(double) * n));
        FGB_CUDA(ctx, cudaMemcpy(ctx->pois_dev[a], pois.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
        FGB_CUDA(ctx, cudaMemcpy(ctx->xi2pi_dev[a], xi2pi.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
        FGB_CUDA(ctx, cudaMemcpy(ctx->wtan_dev[a], wtan.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
        FGB_CUDA(ctx, cudaMemcpy(ctx->wex_dev[a], wex.data(), sizeof(double2) * n, cudaMemcpyHostToDevice));
    }
    return FGB_OK;
}

void fgb_fft_free(fgb_ctx* ctx) {
    for (int a = 0; a < 3; a++) {
        if (ctx->tw_dev[a]) cudaFree(ctx->tw_dev[a]);
        if (ctx->kpm_dev[a]) cudaFree(ctx->kpm_dev[a]);
        if (ctx->kp_dev[a]) cudaFree(ctx->kp_dev[a]);
        if (ctx->xi_dev[a]) cudaFree(ctx->xi_dev[a]);
        if (ctx->xi2pi_dev[a]) cudaFree(ctx->xi2pi_dev[a]);
        if (ctx->wtan_dev[a]) cudaFree(ctx->wtan_dev[a]);
        if (ctx->wex_dev[a]) cudaFree(ctx->wex_dev[a]);
        if (ctx->pois_dev[a]) cudaFree(ctx->pois_dev[a]);
    }
}


// ---- z ---------------------------------------------------------------------------------------------
template <int N, int R1, int R2>
static void launch_z_p2(fgb_ctx* ctx, double* base, long rows, long rowstride, bool fwd, double scale) {
    constexpr int PP = FFTZ_THREADS / Max<R1, R2>::v;
    const long pairs = (rows + 1) / 2;
    const unsigned grid = (unsigned)((pairs + PP - 1) / PP);
    const size_t smem = sizeof(double2) * (N + (size_t)PP * R1 * (R2 + 1));
    if (fwd) {
        set_smem(k_fftz_p2<N, R1, R2, 1>, smem);
        k_fftz_p2<N, R1, R2, 1><<<grid, FFTZ_THREADS, smem, ctx->stream>>>(base, rows, rowstride, ctx->plan[2].tw, scale);
    } else {
        set_smem(k_fftz_p2<N, R1, R2, 0>, smem);
        k_fftz_p2<N, R1, R2, 0><<<grid, FFTZ_THREADS, smem, ctx->stream>>>(base, rows, rowstride, ctx->plan[2].tw, 1.0);
    }
}

template <int N2, int R1, int R2>
static void launch_z_h(fgb_ctx* ctx, double* base, long rows, long rowstride, bool fwd, double scale) {
    constexpr int PP = FFTZ_THREADS / Max<R1, R2>::v;
    const unsigned grid = (unsigned)((rows + PP - 1) / PP);
    const size_t smem = sizeof(double2) * (2 * N2 + (size_t)PP * R1 * (R2 + 1));
    if (fwd) {
        set_smem(k_fftz_h<N2, R1, R2, 1>, smem);
        k_fftz_h<N2, R1, R2, 1><<<grid, FFTZ_THREADS, smem, ctx->stream>>>(base, rows, rowstride, ctx->plan[2].tw, scale);
    } else {
        set_smem(k_fftz_h<N2, R1, R2, 0>, smem);
        k_fftz_h<N2, R1, R2, 0><<<grid, FFTZ_THREADS, smem, ctx->stream>>>(base, rows, rowstride, ctx->plan[2].tw, 1.0);
    }
}

static int fft_z(fgb_ctx* ctx, double* base, int ncomp, const FftLayout& lay, bool fwd) {
    const GridDev& g = ctx->g;
    const long rows = (long)ncomp * g.lnx * g.ny;
    if (rows == 0) return FGB_OK;
    const long rowstride = 2L * lay.nzcs;
    const double scale = 1.0 / ((double)g.nx * g.ny * g.nz);
    ProfScope ps(ctx, fwd ? "fft_z_r2c" : "fft_z_c2r");
    // nz = 512: the half-length form (256-point complex transform per row) beats the 512-point row-pair kernel on the c2r side
    // (0.106 vs 0.134 ms on 128x128x512) and ties on r2c; at 256 and 1024 the row-pair kernel is faster (measured), FGB_ZH forces it
    static const bool zh = getenv("FGB_ZH") != nullptr, no_zh = getenv("FGB_NO_ZH") != nullptr;
    if (!no_zh && (g.nz == 512 || (zh && (g.nz == 128 || g.nz == 256 || g.nz == 1024)))) {
        switch (g.nz) {
            case 128: launch_z_h<64, 8, 8>(ctx, base, rows, rowstride, fwd, scale); break;
            case 256: launch_z_h<128, 16, 8>(ctx, base, rows, rowstride, fwd, scale); break;
            case 512: launch_z_h<256, 16, 16>(ctx, base, rows, rowstride, fwd, scale); break;
            case 1024: launch_z_h<512, 32, 16>(ctx, base, rows, rowstride, fwd, scale); break;
        }
        FGB_CHECK_LAUNCH(ctx, "k_fftz_h");
        return FGB_OK;
    }
    if (is_fast_pow2(g.nz)) {
        switch (g.nz) {
            case 64: launch_z_p2<64, 8, 8>(ctx, base, rows, rowstride, fwd, scale); break;
            case 128: launch_z_p2<128, 16, 8>(ctx, base, rows, rowstride, fwd, scale); break;
            case 256: launch_z_p2<256, 16, 16>(ctx, base, rows, rowstride, fwd, scale); break;
            case 512: launch_z_p2<512, 32, 16>(ctx, base, rows, rowstride, fwd, scale); break;
            case 1024: launch_z_p2<1024, 32, 32>(ctx, base, rows, rowstride, fwd, scale); break;
        }
        FGB_CHECK_LAUNCH(ctx, "k_fftz_p2");
        return FGB_OK;
    }
    int T = pick_T(ctx, g.nz, 2, 1, 8);
    if (!T) return fgb_fail(ctx, FGB_EUNSUPPORTED, "nz=%d does not fit shared memory", g.nz);
    const int TS = T + 1;
    const size_t smem = (size_t)2 * g.nz * TS * sizeof(double2);
    const long pairs = (rows + 1) / 2;
    const unsigned grid = (unsigned)((pairs + T - 1) / T);
    if (fwd) {
        FGB_CUDA(ctx, set_smem(k_fft_z<1>, smem));
        k_fft_z<1><<<grid, 256, smem, ctx->stream>>>(base, rows, g.nz, g.nzc, rowstride, ctx->plan[2], T, TS, scale);
    } else {
        FGB_CUDA(ctx, set_smem(k_fft_z<0>, smem));
        k_fft_z<0><<<grid, 256, smem, ctx->stream>>>(base, rows, g.nz, g.nzc, rowstride, ctx->plan[2], T, TS, 1.0);
    }
    FGB_CHECK_LAUNCH(ctx, "k_fft_z");
    return FGB_OK;
}

int fgb_fft_z_forward(fgb_ctx* ctx, double* base, int ncomp, const FftLayout& lay) { return fft_z(ctx, base, ncomp, lay, true); }
int fgb_fft_z_backward(fgb_ctx* ctx, double* base, int ncomp, const FftLayout& lay) { return fft_z(ctx, base, ncomp, lay, false); }

// ---- strided (y, plain x) -----------------------------------------------------------------------------
template <int N, int R1, int R2, int T>
static void launch_s_p2(fgb_ctx* ctx, const double2* src, double2* dst, const double2* tw, const PencilMap& mi, const PencilMap& mo,
                        int ninner, int nouter, int ncomp, int dir, const PeerTable& pt) {
    dim3 grid((ninner + T - 1) / T, nouter, ncomp);
    constexpr int NT = Max<R1, R2>::v * T;
    const size_t smem = sizeof(double2) * (size_t)(N + N * T);
    if (dir < 0) {
        set_smem(k_ffts_p2<N, R1, R2, -1, T>, smem);
        k_ffts_p2<N, R1, R2, -1, T><<<grid, NT, smem, ctx->stream>>>(src, dst, tw, mi, mo, ninner, pt);
    } else {
        set_smem(k_ffts_p2<N, R1, R2, +1, T>, smem);
        k_ffts_p2<N, R1, R2, +1, T><<<grid, NT, smem, ctx->stream>>>(src, dst, tw, mi, mo, ninner, pt);
    }
}

template <int R1, int R2, int R3, int T>
static void launch_s_p3(fgb_ctx* ctx, const double2* src, double2* dst, const double2* tw, const PencilMap& mi, const PencilMap& mo,
                        int ninner, int nouter, int ncomp, int dir, const PeerTable& pt) {
    using P = p3::Plan<R1, R2, R3, T>;
    dim3 grid((ninner + T - 1) / T, nouter, ncomp);
    constexpr int NT = P::TPP * T;
    const size_t smem = sizeof(double2) * (size_t)(P::N + P::INPLACE_ELEMS);
    if (dir < 0) {
        set_smem(k_ffts_p3<R1, R2, R3, -1, T>, smem);
        k_ffts_p3<R1, R2, R3, -1, T><<<grid, NT, smem, ctx->stream>>>(src, dst, tw, mi, mo, ninner, pt);
    } else {
        set_smem(k_ffts_p3<R1, R2, R3, +1, T>, smem);
        k_ffts_p3<R1, R2, R3, +1, T><<<grid, NT, smem, ctx->stream>>>(src, dst, tw, mi, mo, ninner, pt);
    }
}

int fgb_fft_strided(fgb_ctx* ctx, int axis, const double* src_, double* dst_, const PencilMap& mi, const PencilMap& mo, int ninner,
                    int nouter, int ncomp, int dir, const PeerTable* peers) {
    const int n = ctx->plan[axis].n;
    PeerTable pt;
    pt.n = 0;
    if (peers) pt = *peers;
    const double2* src = (const double2*)src_;
    double2* dst = (double2*)dst_;
    if (ninner == 0 || nouter == 0) return FGB_OK;
    if (n == 1 && src == dst) return FGB_OK;
    const double2* tw = ctx->plan[axis].tw;
    // 512: three passes of radix 8 (64 registers, 3 CTAs/SM) beat the 32x16 two-pass kernel; at 1024 the two-pass kernel wins
    static const bool no_p3 = getenv("FGB_NO_P3") != nullptr;
    if (n == 512 && !no_p3) {
        launch_s_p3<8, 8, 8, 8>(ctx, src, dst, tw, mi, mo, ninner, nouter, ncomp, dir, pt);
        FGB_CHECK_LAUNCH(ctx, "k_ffts_p3");
        return FGB_OK;
    }
    // 1024 with stores into peer memory: 8-lane tiles (128-byte segments over NVLink); the two-pass kernel only fits 4 lanes
    static const bool s1024_p3 = getenv("FGB_S1024_P3") != nullptr;          // A/B: the three-pass 8-lane kernel for local 1024-point passes too
    if (n == 1024 && (pt.n > 0 || s1024_p3) && !no_p3) {
        launch_s_p3<16, 8, 8, 8>(ctx, src, dst, tw, mi, mo, ninner, nouter, ncomp, dir, pt);
        FGB_CHECK_LAUNCH(ctx, "k_ffts_p3");
        return FGB_OK;
    }
    if (is_fast_pow2(n)) {
        switch (n) {
            case 64: launch_s_p2<64, 8, 8, 8>(ctx, src, dst, tw, mi, mo, ninner, nouter, ncomp, dir, pt); break;
            case 128: launch_s_p2<128, 16, 8, 8>(ctx, src, dst, tw, mi, mo, ninner, nouter, ncomp, dir, pt); break;
            case 256: launch_s_p2<256, 16, 16, 8>(ctx, src, dst, tw, mi, mo, ninner, nouter, ncomp, dir, pt); break;
            case 512: launch_s_p2<512, 32, 16, 8>(ctx, src, dst, tw, mi, mo, ninner, nouter, ncomp, dir, pt); break;
            case 1024: launch_s_p2<1024, 32, 32, 4>(ctx, src, dst, tw, mi, mo, ninner, nouter, ncomp, dir, pt); break;
        }
        FGB_CHECK_LAUNCH(ctx, "k_ffts_p2");
        return FGB_OK;
    }
    int T = pick_T(ctx, n, 2, 0, 8);
    if (!T) return fgb_fail(ctx, FGB_EUNSUPPORTED, "axis length %d does not fit shared memory", n);
    const size_t smem = (size_t)2 * n * T * sizeof(double2);
    dim3 grid((ninner + T - 1) / T, nouter, ncomp);
    if (dir < 0) {
        FGB_CUDA(ctx, set_smem(k_fft_strided<-1>, smem));
        k_fft_strided<-1><<<grid, 256, smem, ctx->stream>>>(src, dst, ctx->plan[axis], mi, mo, ninner, T, pt);
    } else {
        FGB_CUDA(ctx, set_smem(k_fft_strided<1>, smem));
        k_fft_strided<1><<<grid, 256, smem, ctx->stream>>>(src, dst, ctx->plan[axis], mi, mo, ninner, T, pt);
    }
    FGB_CHECK_LAUNCH(ctx, "k_fft_strided");
    return FGB_OK;
}

int fgb_fft_y(fgb_ctx* ctx, double* base, int ncomp, const FftLayout& lay, int dir) {
    const GridDev& g = ctx->g;
    ProfScope ps(ctx, dir < 0 ? "fft_y_fwd" : "fft_y_bwd");
    const PencilMap m = {lay.nzcs, g.ny, 0, (long)g.ny * lay.nzcs, (long)g.lnx * g.ny * lay.nzcs};
    return fgb_fft_strided(ctx, 1, base, base, m, m, g.nzc, g.lnx, ncomp, dir);
}

static void fill_green(const fgb_ctx* ctx, const GreenArgs* ga, GreenDev& G) {
    const GridDev& g = ctx->g;
    G.kind = ga->kind;
    G.c10 = ga->c10;
    G.c20 = ga->c20;
    G.beta = ga->beta;
    G.alpha = ga->alpha;
    for (int i = 0; i < 9; i++) G.dc[i] = ga->dc[i];
    G.freq_hack = ga->freq_hack;
    G.nx = g.nx; G.ny = g.ny; G.nz = g.nz;
    for (int a = 0; a < 3; a++) {
        G.kpm[a] = ctx->kpm_dev[a]; G.kp[a] = ctx->kp_dev[a]; G.xi[a] = ctx->xi_dev[a];
        G.xi2pi[a] = ctx->xi2pi_dev[a]; G.wex[a] = ctx->wex_dev[a]; G.wtan[a] = ctx->wtan_dev[a];
        G.wvox[a] = ctx->L[a] / (a == 0 ? g.nx : a == 1 ? g.ny : g.nz);
        G.pois[a] = ctx->pois_dev[a];
    }
}

// x pass on a buffer whose x extent is complete: element (ii, jj, kk) at base[c*cstride + (jj-jbase)*ostride + ii*estride + kk]
int fgb_fft_x_green_layout(fgb_ctx* ctx, double* base, const GreenArgs* ga, long estride, int nzc_valid, int nouter, long ostride,
                           long cstride, int jbase, const PencilMap* out_map, const PeerTable* peers) {
    GreenDev G;
    fill_green(ctx, ga, G);
    PencilMap xo = {estride, ctx->g.nx, 0, ostride, cstride};       // default: store back in place
    if (out_map) xo = *out_map;
    PeerTable pt;
    pt.n = 0;
    if (peers) pt = *peers;
    double2* b = (double2*)base;
    ProfScope ps(ctx, "fft_x_green");
    switch (ga->kind) {
        case 1: return fgb_xg_staggered3(ctx, b, G, estride, nzc_valid, nouter, ostride, cstride, jbase, xo, pt);
        case 2: return fgb_xg_staggered1(ctx, b, G, estride, nzc_valid, nouter, ostride, cstride, jbase, xo, pt);
        case 3: return fgb_xg_colloc6(ctx, b, G, estride, nzc_valid, nouter, ostride, cstride, jbase, xo, pt);
        case 4: return fgb_xg_colloc3(ctx, b, G, estride, nzc_valid, nouter, ostride, cstride, jbase, xo, pt);
        case 5: return fgb_xg_colloc9(ctx, b, G, estride, nzc_valid, nouter, ostride, cstride, jbase, xo, pt);
        case 6: return fgb_xg_g0div9(ctx, b, G, estride, nzc_valid, nouter, ostride, cstride, jbase, xo, pt);
        case 7: return fgb_xg_grad9(ctx, b, G, estride, nzc_valid, nouter, ostride, cstride, jbase, xo, pt);
        case 8: return fgb_xg_willot6(ctx, b, G, estride, nzc_valid, nouter, ostride, cstride, jbase, xo, pt);
        case 11: return fgb_xg_gradg0div9(ctx, b, G, estride, nzc_valid, nouter, ostride, cstride, jbase, xo, pt);
        case 10: return fgb_xg_poisson1(ctx, b, G, estride, nzc_valid, nouter, ostride, cstride, jbase, xo, pt);
        case 9: return fgb_xg_colloc6_zt(ctx, b, G, estride, nzc_valid, nouter, ostride, cstride, jbase, xo, pt);
    }
    return fgb_fail(ctx, FGB_EINVAL, "unknown Green operator kind %d", ga->kind);
}

int fgb_fft_x(fgb_ctx* ctx, double* base, int ncomp, const FftLayout& lay, int dir, const GreenArgs* ga) {
    const GridDev& g = ctx->g;
    if (ctx->nranks > 1) return fgb_fail(ctx, FGB_EUNSUPPORTED, "fgb_fft_x: slab-partitioned x pass goes through comm.cu");
    const long estride = (long)g.ny * lay.nzcs;
    const long cstride = (long)g.lnx * g.ny * lay.nzcs;
    if (!ga || ga->kind == 0) {
        ProfScope ps(ctx, dir < 0 ? "fft_x_fwd" : "fft_x_bwd");
        const PencilMap m = {estride, g.nx, 0, (long)lay.nzcs, cstride};
        return fgb_fft_strided(ctx, 0, base, base, m, m, g.nzc, g.ny, ncomp, dir);
    }
    return fgb_fft_x_green_layout(ctx, base, ga, estride, g.nzc, g.ny, lay.nzcs, cstride, 0);
}
