// Per-voxel constitutive sweeps:
//   calcStress fg:18134-18184, calcStressDeriv fg:18425-18478, calcPolarizationDim fg:18044-18118,
//   meanPK1 fg:12312, meanW fg:12239, calcMinDetF fg:17871, getRefMaterial/eig fg:12153-12236, fg:12472-12559.
// One thread per voxel, SoA component planes (coalesced along z), no virtual dispatch: the law set is a
// small by-value table and the mixing rule a template-free switch that is uniform across a warp except
// at interface voxels.
#include "material.cuh"
#include "reduce.cuh"

static unsigned grid_for(const fgb_ctx* ctx, size_t n, int block) {
    size_t b = (n + block - 1) / block;
    size_t cap = (size_t)ctx->red_blocks;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

MaterialDev fgb_material_dev(fgb_ctx* ctx) {
    MaterialDev M;
    memset(&M, 0, sizeof(M));
    M.nphases = ctx->nphases;
    M.mix = ctx->mix;
    // select_dfg (fg:18146): with the doubly fine grid the phase data of the fine grid is used
    const bool f = ctx->dfg != 0;
    const double* normals = f ? ctx->normals_f : ctx->normals;
    const double* orient = f ? ctx->orient_f : ctx->orient;
    const size_t plane = f ? ctx->gf.plane : ctx->g.plane;
    for (int p = 0; p < ctx->nphases; p++) {
        M.phi[p] = f ? ctx->phi_f[p] : ctx->phi[p];
        M.law[p] = ctx->laws[p];
    }
    for (int a = 0; a < 3; a++) {
        M.normals[a] = normals ? normals + (size_t)a * plane : nullptr;
        M.orient[a] = orient ? orient + (size_t)a * plane : nullptr;
    }
    M.lam = ctx->lam;
    return M;
}

// The grid a constitutive sweep runs on.  With the doubly fine grid (half_staggered / full_staggered) the operands are prolongated
// into _temp_dfg_1 / _temp_dfg_2, the sweep runs in place on the fine grid and the result is restricted (fg:18143-18149, 18343-18347).
struct SweepGrid {
    GridDev g;
    const double* a;      // first operand
    const double* b;      // second operand (tangent sweeps) or null
    double* out;          // where the kernel writes
    double* final_dst;    // coarse destination to restrict into (dfg) or null
};
static int sweep_begin(fgb_ctx* ctx, const double* a, const double* b, double* dst, SweepGrid& S) {
    S.g = ctx->g; S.a = a; S.b = b; S.out = dst; S.final_dst = nullptr;
    if (!ctx->dfg) return FGB_OK;
    int rc;
    if ((rc = fgb_k_prolongate(ctx, a, ctx->dfg1))) return rc;
    S.a = ctx->dfg1;
    if (b) {
        if (!ctx->dfg2) {
            cudaError_t e = cudaMalloc(&ctx->dfg2, sizeof(double) * ctx->gf.plane * ctx->dim);
            if (e != cudaSuccess) { ctx->dfg2 = nullptr; return fgb_fail(ctx, FGB_ENOMEM, "cannot allocate the second doubly-fine-grid field"); }
        }
        if ((rc = fgb_k_prolongate(ctx, b, ctx->dfg2))) return rc;
        S.b = ctx->dfg2;
    }
    S.g = ctx->gf;
    S.out = ctx->dfg1;
    S.final_dst = dst;
    return FGB_OK;
}
static int sweep_end(fgb_ctx* ctx, const SweepGrid& S) {
    if (S.final_dst) return fgb_k_restrict(ctx, S.out, S.final_dst);
    return FGB_OK;
}

// 32-bit voxel index arithmetic (a slab never exceeds 2^32 voxels); 64-bit divisions were the bottleneck of v1
#define VOXEL_LOOP(g)                                                                                              \
    const unsigned nvox = (unsigned)g.lnx * (unsigned)g.ny * (unsigned)g.nz;                                       \
    for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += gridDim.x * blockDim.x)
#define VOXEL_OFFSET(g)                         \
    const unsigned row_ = v / (unsigned)g.nz;   \
    const size_t o = (size_t)row_ * g.nzp + (v - row_ * (unsigned)g.nz);

// sigma = alpha*P_mix(eps) + beta*eps + gamma*tr(eps)*I ; DERIV: with dP_mix/dF(F):W and W in the correction terms
template <int D, int DERIV>
__global__ void __launch_bounds__(256) k_calc_stress(const double* __restrict__ Fsrc, const double* __restrict__ Wsrc,
                                                     double* __restrict__ dst, GridDev g, MaterialDev M, double alpha, double beta,
                                                     double gamma, int* flag) {
    VOXEL_LOOP(g) {
        VOXEL_OFFSET(g)
        double F[9], P[9];
#pragma unroll
        for (int d = 0; d < D; d++) F[d] = Fsrc[(size_t)d * g.plane + o];
        if (DERIV) {
            double W[9];
#pragma unroll
            for (int d = 0; d < D; d++) W[d] = Wsrc[(size_t)d * g.plane + o];
            Mixed<D>::dPK1(M, o, F, alpha, W, P, flag);
            if (beta != 0) {
#pragma unroll
                for (int d = 0; d < D; d++) P[d] += beta * W[d];
            }
            if (gamma != 0) {
                const double tr = W[0] + W[1] + W[2];
                P[0] += gamma * tr; P[1] += gamma * tr; P[2] += gamma * tr;
            }
        } else {
            Mixed<D>::PK1(M, o, F, alpha, P, flag);
            if (beta != 0) {
#pragma unroll
                for (int d = 0; d < D; d++) P[d] += beta * F[d];
            }
            if (gamma != 0) {
                const double tr = F[0] + F[1] + F[2];
                P[0] += gamma * tr; P[1] += gamma * tr; P[2] += gamma * tr;
            }
        }
#pragma unroll
        for (int d = 0; d < D; d++) dst[(size_t)d * g.plane + o] = P[d];
    }
}

// dense solve A x = b (partial pivoting), A is D x D row-major and is destroyed
template <int D>
__device__ void solve_dense(double* A, double* b) {
    for (int c = 0; c < D; c++) {
        int piv = c;
        double best = fabs(A[c * D + c]);
        for (int r = c + 1; r < D; r++) {
            const double v = fabs(A[r * D + c]);
            if (v > best) { best = v; piv = r; }
        }
        if (piv != c) {
            for (int k = 0; k < D; k++) { const double t = A[c * D + k]; A[c * D + k] = A[piv * D + k]; A[piv * D + k] = t; }
            const double t = b[c]; b[c] = b[piv]; b[piv] = t;
        }
        const double id = 1.0 / A[c * D + c];
        for (int r = c + 1; r < D; r++) {
            const double f = A[r * D + c] * id;
            for (int k = c; k < D; k++) A[r * D + k] -= f * A[c * D + k];
            b[r] -= f * b[c];
        }
    }
    for (int r = D - 1; r >= 0; r--) {
        double s = b[r];
        for (int k = r + 1; k < D; k++) s -= A[r * D + k] * b[k];
        b[r] = s / A[r * D + r];
    }
}

// tangent "rows" matrix C[m][:] = dP_mix/dF(F):e_m  (fg:10426, fg:12540)
template <int D>
__device__ void tangent_rows(const MaterialDev& M, size_t o, const double* F, double* C, int* flag) {
    for (int m = 0; m < D; m++) {
        double e[9], r[9];
#pragma unroll
        for (int k = 0; k < 9; k++) e[k] = (k == m) ? 1.0 : 0.0;
        Mixed<D>::dPK1(M, o, F, 1.0, e, r, flag);
        for (int k = 0; k < D; k++) C[m * D + k] = r[k];
    }
}

// Eyre-Milton polarization map, fg:18044-18118 with the law hooks fg:11427-11467 / fg:10414-10445
template <int D>
__global__ void __launch_bounds__(128) k_calc_polarization(const double* __restrict__ src, double* __restrict__ dst, GridDev g,
                                                           MaterialDev M, double mu_0, int inv, int* flag) {
    VOXEL_LOOP(g) {
        VOXEL_OFFSET(g)
        double F[9], P[9];
#pragma unroll
        for (int d = 0; d < D; d++) F[d] = src[(size_t)d * g.plane + o];
        int pure = -1;
        for (int p = 0; p < M.nphases; p++)
            if (M.phi[p][o] == 1) { pure = p; break; }
        if (D == 6 && pure >= 0 && M.law[pure].id == FGB_LAW_ISO) {
            const double mu = M.law[pure].p[0], lambda = M.law[pure].p[1];
            double m = 2.0 * (mu + mu_0);
            const double a = 1.0 / m;
            const double b = lambda / (m * (3.0 * lambda + m));
            const double trF = F[0] + F[1] + F[2];
            P[0] = a * F[0] - b * trF; P[1] = a * F[1] - b * trF; P[2] = a * F[2] - b * trF;
            P[3] = a * F[3]; P[4] = a * F[4]; P[5] = a * F[5];
            if (!inv) {
                m = 2.0 * (mu - mu_0);
                const double trP = P[0] + P[1] + P[2];
                P[0] = m * P[0] + lambda * trP; P[1] = m * P[1] + lambda * trP; P[2] = m * P[2] + lambda * trP;
                P[3] = m * P[3]; P[4] = m * P[4]; P[5] = m * P[5];
            }
        } else {
            // generic: C2 = C + 2 mu_0 I ; solve C2^T Q = F (gesv on the row-major storage) ; P = C Q - 2 mu_0 Q
            double C[D * D], A[D * D], Q[D];
            tangent_rows<D>(M, o, F, C, flag);
            for (int r = 0; r < D; r++)
                for (int c = 0; c < D; c++) A[r * D + c] = C[c * D + r] + ((r == c) ? 2.0 * mu_0 : 0.0);
            for (int d = 0; d < D; d++) Q[d] = F[d];
            solve_dense<D>(A, Q);
            if (inv) {
                for (int d = 0; d < D; d++) P[d] = Q[d];
            } else {
                for (int r = 0; r < D; r++) {
                    double s = 0;
                    for (int c = 0; c < D; c++) s += C[r * D + c] * Q[c];
                    P[r] = s - 2.0 * mu_0 * Q[r];
                }
            }
        }
#pragma unroll
        for (int d = 0; d < D; d++) dst[(size_t)d * g.plane + o] = P[d];
    }
}

template <int D>
__global__ void __launch_bounds__(256) k_mean_pk1(const double* __restrict__ src, GridDev g, MaterialDev M, double alpha,
                                                  double* __restrict__ partials, int* flag) {
    double s[D];
#pragma unroll
    for (int d = 0; d < D; d++) s[d] = 0;
    VOXEL_LOOP(g) {
        VOXEL_OFFSET(g)
        double F[9], P[9];
#pragma unroll
        for (int d = 0; d < D; d++) F[d] = src[(size_t)d * g.plane + o];
        Mixed<D>::PK1(M, o, F, alpha, P, flag);
#pragma unroll
        for (int d = 0; d < D; d++) s[d] += P[d];
    }
    block_reduce_store<D, 0>(s, partials);
}

// mean Cauchy stress: sigma = P(F) F^T / det F per voxel (Cauchy fg:10326-10346, meanCauchy fg:12268-12308); 9 components
__global__ void __launch_bounds__(256) k_mean_cauchy(const double* __restrict__ src, GridDev g, MaterialDev M, double alpha,
                                                     double* __restrict__ partials, int* flag) {
    // stored component order as a matrix (fg:9262-9276)
    const int IDX[3][3] = {{0, 5, 4}, {8, 1, 3}, {7, 6, 2}};
    double s[9];
#pragma unroll
    for (int d = 0; d < 9; d++) s[d] = 0;
    VOXEL_LOOP(g) {
        VOXEL_OFFSET(g)
        double F[9], P[9];
#pragma unroll
        for (int d = 0; d < 9; d++) F[d] = src[(size_t)d * g.plane + o];
        const double c = 1.0 / det9(F);
        Mixed<9>::PK1(M, o, F, c * alpha, P, flag);
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) s[IDX[i][j]] += P[IDX[i][0]] * F[IDX[j][0]] + P[IDX[i][1]] * F[IDX[j][1]] + P[IDX[i][2]] * F[IDX[j][2]];
    }
    block_reduce_store<9, 0>(s, partials);
}

template <int D>
__global__ void __launch_bounds__(256) k_mean_energy(const double* __restrict__ src, GridDev g, MaterialDev M,
                                                     double* __restrict__ partials, int* flag) {
    double s[1] = {0};
    VOXEL_LOOP(g) {
        VOXEL_OFFSET(g)
        double F[9];
#pragma unroll
        for (int d = 0; d < D; d++) F[d] = src[(size_t)d * g.plane + o];
        s[0] += Mixed<D>::W(M, o, F, flag);
    }
    block_reduce_store<1, 0>(s, partials);
}

__global__ void __launch_bounds__(256) k_min_detF(const double* __restrict__ src, GridDev g, double* __restrict__ partials) {
    double s[1] = {INFINITY};
    VOXEL_LOOP(g) {
        VOXEL_OFFSET(g)
        double F[9];
#pragma unroll
        for (int d = 0; d < 9; d++) F[d] = src[(size_t)d * g.plane + o];
        s[0] = fmin(s[0], det9(F));
    }
    block_reduce_store<1, 1>(s, partials);
}

// cyclic Jacobi eigenvalues of a symmetric N x N matrix (only min/max are returned)
template <int N>
__device__ void jacobi_minmax(double* A, double& lmin, double& lmax) {
    for (int sweep = 0; sweep < 30; sweep++) {
        double off = 0, diag = 0;
        for (int p = 0; p < N; p++) {
            diag += A[p * N + p] * A[p * N + p];
            for (int q = p + 1; q < N; q++) off += A[p * N + q] * A[p * N + q];
        }
        if (off <= 1e-30 * diag || off == 0) break;
        for (int p = 0; p < N - 1; p++) {
            for (int q = p + 1; q < N; q++) {
                const double apq = A[p * N + q];
                if (apq == 0) continue;
                const double theta = (A[q * N + q] - A[p * N + p]) / (2 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
                const double c = 1 / sqrt(t * t + 1), s = t * c;
                for (int k = 0; k < N; k++) {
                    const double akp = A[k * N + p], akq = A[k * N + q];
                    A[k * N + p] = c * akp - s * akq;
                    A[k * N + q] = s * akp + c * akq;
                }
                for (int k = 0; k < N; k++) {
                    const double apk = A[p * N + k], aqk = A[q * N + k];
                    A[p * N + k] = c * apk - s * aqk;
                    A[q * N + k] = s * apk + c * aqk;
                }
            }
        }
    }
    lmin = INFINITY;
    lmax = -INFINITY;
    for (int p = 0; p < N; p++) {
        lmin = fmin(lmin, A[p * N + p]);
        lmax = fmax(lmax, A[p * N + p]);
    }
}

// getRefMaterial: eigenvalues of the symmetric matrix defined by the LOWER triangle of the row-major tangent rows
// (lapack::syev('N','U') on ublas row-major storage reads exactly those entries, fg:12509/12522)
template <int D, int ZT>
__global__ void __launch_bounds__(128) k_ref_material(const double* __restrict__ src, GridDev g, MaterialDev M,
                                                      double* __restrict__ partials, int linear_laws, int* flag) {
    double mm[2] = {INFINITY, INFINITY};   // min(lambda), min(-lambda)
    VOXEL_LOOP(g) {
        VOXEL_OFFSET(g)
        double F[9], C[D * D];
#pragma unroll
        for (int d = 0; d < D; d++) F[d] = src[(size_t)d * g.plane + o];
        tangent_rows<D>(M, o, F, C, flag);
        constexpr int N = ZT ? D - 1 : D;
        double A[N * N];
        for (int r = 0; r < N; r++)
            for (int c = 0; c <= r; c++) {
                const double v2 = C[(r + ZT) * D + (c + ZT)];
                A[r * N + c] = v2;
                A[c * N + r] = v2;
            }
        double lo, hi;
        jacobi_minmax<N>(A, lo, hi);
        mm[0] = fmin(mm[0], lo);
        mm[1] = fmin(mm[1], -hi);
    }
    block_reduce_store<2, 1>(mm, partials);
}

// ------------------------------------------------------------------------------------------------
#define DISPATCH_D(ctx, K3, K6, K9)        \
    do {                                   \
        if ((ctx)->dim == 3) { K3; }       \
        else if ((ctx)->dim == 6) { K6; }  \
        else { K9; }                       \
    } while (0)

static int check_material(fgb_ctx* ctx) {
    if (ctx->nphases < 1) return fgb_fail(ctx, FGB_EINVAL, "no materials specified");     // fg:15306
    for (int p = 0; p < ctx->nphases; p++)
        if (!(ctx->dfg ? ctx->phi_f[p] : ctx->phi[p])) return fgb_fail(ctx, FGB_EINVAL, "phase %d has no volume fraction field", p);
    if (ctx->mix == FGB_MIX_LAMINATE && !(ctx->dfg ? ctx->normals_f : ctx->normals)) return fgb_fail(ctx, FGB_EINVAL, "laminate mixing needs normals");
    for (int p = 0; p < ctx->nphases; p++)
        if (ctx->laws[p].id == FGB_LAW_TISO && !(ctx->dfg ? ctx->orient_f : ctx->orient) && ctx->laws[p].p[5] == 0 && ctx->laws[p].p[6] == 0 && ctx->laws[p].p[7] == 0)
            return fgb_fail(ctx, FGB_EINVAL, "tiso law needs the orientation field or a constant axis");
    return FGB_OK;
}

static int check_flag(fgb_ctx* ctx) {
    // polled lazily by the scheme-level entry points together with the scalars they return
    return FGB_OK;
}

int fgb_k_calc_stress(fgb_ctx* ctx, const double* src, double* dst, double mu0, double lambda0, double alpha) {
    int rc = check_material(ctx);
    if (rc) return rc;
    const double beta = -alpha * 2 * mu0, gamma = -alpha * lambda0;
    const MaterialDev M = fgb_material_dev(ctx);
    SweepGrid S;
    if ((rc = sweep_begin(ctx, src, nullptr, dst, S))) return rc;
    const size_t nvox = (size_t)S.g.lnx * S.g.ny * S.g.nz;
    const unsigned grid = grid_for(ctx, nvox, 256);
    ProfScope ps(ctx, "calc_stress");
    DISPATCH_D(ctx, (k_calc_stress<3, 0><<<grid, 256, 0, ctx->stream>>>(S.a, nullptr, S.out, S.g, M, alpha, beta, gamma, ctx->d_flag)),
               (k_calc_stress<6, 0><<<grid, 256, 0, ctx->stream>>>(S.a, nullptr, S.out, S.g, M, alpha, beta, gamma, ctx->d_flag)),
               (k_calc_stress<9, 0><<<grid, 256, 0, ctx->stream>>>(S.a, nullptr, S.out, S.g, M, alpha, beta, gamma, ctx->d_flag)));
    FGB_CHECK_LAUNCH(ctx, "k_calc_stress");
    return sweep_end(ctx, S);
}

int fgb_k_calc_stress_deriv(fgb_ctx* ctx, const double* F, const double* W, double* dst, double mu0, double lambda0, double alpha) {
    int rc = check_material(ctx);
    if (rc) return rc;
    const double beta = -alpha * 2 * mu0, gamma = -alpha * lambda0;
    const MaterialDev M = fgb_material_dev(ctx);
    SweepGrid S;
    if ((rc = sweep_begin(ctx, F, W, dst, S))) return rc;
    const size_t nvox = (size_t)S.g.lnx * S.g.ny * S.g.nz;
    const unsigned grid = grid_for(ctx, nvox, 256);
    ProfScope ps(ctx, "calc_stress_deriv");
    DISPATCH_D(ctx, (k_calc_stress<3, 1><<<grid, 256, 0, ctx->stream>>>(S.a, S.b, S.out, S.g, M, alpha, beta, gamma, ctx->d_flag)),
               (k_calc_stress<6, 1><<<grid, 256, 0, ctx->stream>>>(S.a, S.b, S.out, S.g, M, alpha, beta, gamma, ctx->d_flag)),
               (k_calc_stress<9, 1><<<grid, 256, 0, ctx->stream>>>(S.a, S.b, S.out, S.g, M, alpha, beta, gamma, ctx->d_flag)));
    FGB_CHECK_LAUNCH(ctx, "k_calc_stress_deriv");
    return sweep_end(ctx, S);
}

int fgb_k_calc_polarization(fgb_ctx* ctx, const double* src, double* dst, double mu0, int inv) {
    int rc = check_material(ctx);
    if (rc) return rc;
    const MaterialDev M = fgb_material_dev(ctx);
    SweepGrid S;
    if ((rc = sweep_begin(ctx, src, nullptr, dst, S))) return rc;
    const size_t nvox = (size_t)S.g.lnx * S.g.ny * S.g.nz;
    const unsigned grid = grid_for(ctx, nvox, 128);
    ProfScope ps(ctx, "calc_polarization");
    DISPATCH_D(ctx, (k_calc_polarization<3><<<grid, 128, 0, ctx->stream>>>(S.a, S.out, S.g, M, mu0, inv, ctx->d_flag)),
               (k_calc_polarization<6><<<grid, 128, 0, ctx->stream>>>(S.a, S.out, S.g, M, mu0, inv, ctx->d_flag)),
               (k_calc_polarization<9><<<grid, 128, 0, ctx->stream>>>(S.a, S.out, S.g, M, mu0, inv, ctx->d_flag)));
    FGB_CHECK_LAUNCH(ctx, "k_calc_polarization");
    return sweep_end(ctx, S);
}

int fgb_k_mean_pk1(fgb_ctx* ctx, const double* src, double alpha, double* out) {
    int rc = check_material(ctx);
    if (rc) return rc;
    const MaterialDev M = fgb_material_dev(ctx);
    SweepGrid S;
    if ((rc = sweep_begin(ctx, src, nullptr, nullptr, S))) return rc;
    const size_t nvox = (size_t)S.g.lnx * S.g.ny * S.g.nz;
    const double nxyz = (double)S.g.nx * S.g.ny * S.g.nz;
    const double a = alpha / nxyz;                                     // fg:12318
    const unsigned grid = grid_for(ctx, nvox, 256);
    {
        ProfScope ps(ctx, "mean_pk1");
        DISPATCH_D(ctx, (k_mean_pk1<3><<<grid, 256, 0, ctx->stream>>>(S.a, S.g, M, a, ctx->d_partials, ctx->d_flag)),
                   (k_mean_pk1<6><<<grid, 256, 0, ctx->stream>>>(S.a, S.g, M, a, ctx->d_partials, ctx->d_flag)),
                   (k_mean_pk1<9><<<grid, 256, 0, ctx->stream>>>(S.a, S.g, M, a, ctx->d_partials, ctx->d_flag)));
        FGB_CHECK_LAUNCH(ctx, "k_mean_pk1");
    }
    return fgb_reduce_finish(ctx, grid, ctx->dim, 0, out);
}

int fgb_k_mean_cauchy(fgb_ctx* ctx, const double* src, double alpha, double* out) {
    int rc = check_material(ctx);
    if (rc) return rc;
    if (ctx->dim != 9) return fgb_fail(ctx, FGB_EINVAL, "the Cauchy stress needs the 9-component deformation gradient (hyperelasticity)");
    const MaterialDev M = fgb_material_dev(ctx);
    SweepGrid S;
    if ((rc = sweep_begin(ctx, src, nullptr, nullptr, S))) return rc;
    const size_t nvox = (size_t)S.g.lnx * S.g.ny * S.g.nz;
    const double a = alpha / ((double)S.g.nx * S.g.ny * S.g.nz);               // fg:12274
    const unsigned grid = grid_for(ctx, nvox, 256);
    {
        ProfScope ps(ctx, "mean_cauchy");
        k_mean_cauchy<<<grid, 256, 0, ctx->stream>>>(S.a, S.g, M, a, ctx->d_partials, ctx->d_flag);
        FGB_CHECK_LAUNCH(ctx, "k_mean_cauchy");
    }
    return fgb_reduce_finish(ctx, grid, 9, 0, out);
}

int fgb_k_mean_energy(fgb_ctx* ctx, const double* src, double* out) {
    int rc = check_material(ctx);
    if (rc) return rc;
    const MaterialDev M = fgb_material_dev(ctx);
    SweepGrid S;
    if ((rc = sweep_begin(ctx, src, nullptr, nullptr, S))) return rc;
    const size_t nvox = (size_t)S.g.lnx * S.g.ny * S.g.nz;
    const unsigned grid = grid_for(ctx, nvox, 256);
    {
        ProfScope ps(ctx, "mean_energy");
        DISPATCH_D(ctx, (k_mean_energy<3><<<grid, 256, 0, ctx->stream>>>(S.a, S.g, M, ctx->d_partials, ctx->d_flag)),
                   (k_mean_energy<6><<<grid, 256, 0, ctx->stream>>>(S.a, S.g, M, ctx->d_partials, ctx->d_flag)),
                   (k_mean_energy<9><<<grid, 256, 0, ctx->stream>>>(S.a, S.g, M, ctx->d_partials, ctx->d_flag)));
        FGB_CHECK_LAUNCH(ctx, "k_mean_energy");
    }
    rc = fgb_reduce_finish(ctx, grid, 1, 0, out);
    if (rc) return rc;
    out[0] /= (double)S.g.nx * S.g.ny * S.g.nz;
    return FGB_OK;
}

int fgb_k_min_detF(fgb_ctx* ctx, const double* src, double* out) {
    if (ctx->dim != 9) return fgb_fail(ctx, FGB_EINVAL, "min det(F) needs a 9-component field");
    SweepGrid S;
    int rc = sweep_begin(ctx, src, nullptr, nullptr, S);
    if (rc) return rc;
    const size_t nvox = (size_t)S.g.lnx * S.g.ny * S.g.nz;
    const unsigned grid = grid_for(ctx, nvox, 256);
    k_min_detF<<<grid, 256, 0, ctx->stream>>>(S.a, S.g, ctx->d_partials);
    FGB_CHECK_LAUNCH(ctx, "k_min_detF");
    return fgb_reduce_finish(ctx, grid, 1, 1, out);
}

int fgb_k_ref_material(fgb_ctx* ctx, const double* src, int zero_trace, double* lmin, double* lmax) {
    int rc = check_material(ctx);
    if (rc) return rc;
    const MaterialDev M = fgb_material_dev(ctx);
    SweepGrid S;
    if ((rc = sweep_begin(ctx, src, nullptr, nullptr, S))) return rc;
    const size_t nvox = (size_t)S.g.lnx * S.g.ny * S.g.nz;
    const unsigned grid = grid_for(ctx, nvox, 128);
    int linear = 0;   // the per-phase shortcut is disabled: every voxel is evaluated (robust; once per load step)
    {
        ProfScope ps(ctx, "ref_material");
        if (zero_trace) {
            if (ctx->dim != 6) return fgb_fail(ctx, FGB_EINVAL, "zero_trace reference material is defined for dim 6 only");
            k_ref_material<6, 1><<<grid, 128, 0, ctx->stream>>>(S.a, S.g, M, ctx->d_partials, linear, ctx->d_flag);
        } else {
            DISPATCH_D(ctx, (k_ref_material<3, 0><<<grid, 128, 0, ctx->stream>>>(S.a, S.g, M, ctx->d_partials, linear, ctx->d_flag)),
                       (k_ref_material<6, 0><<<grid, 128, 0, ctx->stream>>>(S.a, S.g, M, ctx->d_partials, linear, ctx->d_flag)),
                       (k_ref_material<9, 0><<<grid, 128, 0, ctx->stream>>>(S.a, S.g, M, ctx->d_partials, linear, ctx->d_flag)));
        }
        FGB_CHECK_LAUNCH(ctx, "k_ref_material");
    }
    double mm[2];
    rc = fgb_reduce_finish(ctx, grid, 2, 1, mm);
    if (rc) return rc;
    *lmin = mm[0];
    *lmax = -mm[1];
    return FGB_OK;
}
