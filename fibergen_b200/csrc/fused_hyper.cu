// Fused sweeps of the inner CG iteration of the Newton-CG scheme for Neo-Hooke phases (BASELINE config 4: runCGHyper fg:22699-23130,
// ApplyOperator fg:23132-23150, calcStressDeriv fg:18425-18478, NeoHookeMaterialLaw::dPK1 fg:11789-11856, Voigt mixing fg:12763).
//
// The tangent is evaluated at the deformation gradient F of the OUTER Newton iteration, which does not change during the inner CG
// solve.  Everything that depends on F only is therefore computed once per Newton iteration into a per-voxel cache,
//     Finv (9 values), a = sum_p phi_p mu_p - 2 mu0, b = sum_p phi_p lambda_p, c = sum_p phi_p (mu_p - lambda_p ln J),
// and an inner iteration applies   dP = a W + b tr(Finv^T W^T) Finv^T + c Finv^T W^T Finv^T   (the Voigt sum of fg:11795-11843 with
// the phase-independent factors taken out) without the 3x3 inverse, determinant and logarithm per voxel and phase.
//
//   k_nh_cache       : F, phi -> cache (once per Newton iteration)
//   k_nh_dir_tangent : Q = R + beta*Q (fg:23086) ; sigma = (dP/dF(F) - C0) : Q (fg:18425)            one elementwise sweep
//   k_hyper_cg_u     : W = grad_h u (epsOperatorStaggeredHyper fg:18763) is not stored: <Q, Q - W> (fg:22905), then
//                      X += alpha Q ; R -= alpha (Q - W) ; <R, R> (fg:22933, fg:23064-23068) re-evaluate it from u
#include "material.cuh"
#include "reduce.cuh"
#include <cstdlib>

#define NH_CACHE_PLANES 12

struct NhPhases {
    int n;
    const double* phi[FGB_MAX_PHASES];
    double mu[FGB_MAX_PHASES], lam[FGB_MAX_PHASES];
};

__global__ void __launch_bounds__(256) k_nh_cache(const double* __restrict__ F, double* __restrict__ cache, GridDev g, NhPhases M, double beta,
                                                  int* flag) {
    const unsigned nvox = (unsigned)g.lnx * (unsigned)g.ny * (unsigned)g.nz;
    for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += gridDim.x * blockDim.x) {
        const unsigned row_ = v / (unsigned)g.nz;
        const size_t o = (size_t)row_ * g.nzp + (v - row_ * (unsigned)g.nz);
        double Fv[9], Fi[9];
#pragma unroll
        for (int d = 0; d < 9; d++) Fv[d] = F[(size_t)d * g.plane + o];
        inv9(Fv, Fi);
        const double lnJ = checked_log(det9(Fv), flag);
        double a = 0, b = 0, c = 0;
        for (int p = 0; p < M.n; p++) {
            const double phi = M.phi[p][o];
            if (phi <= FGB_VOIGT_THRESHOLD) continue;
            a += phi * M.mu[p];
            b += phi * M.lam[p];
            c += phi * (M.mu[p] - M.lam[p] * lnJ);
        }
#pragma unroll
        for (int d = 0; d < 9; d++) cache[(size_t)d * g.plane + o] = Fi[d];
        cache[(size_t)9 * g.plane + o] = a + beta;
        cache[(size_t)10 * g.plane + o] = b;
        cache[(size_t)11 * g.plane + o] = c;
    }
}

// Q = R + cgbeta*Q_old (skipped when R == null), sigma = a Q + b tr(A) Finv^T + c B (+ gamma tr(Q) on the diagonal)
template <int UPDATE>
__global__ void __launch_bounds__(256) k_nh_dir_tangent(const double* __restrict__ R, const double* __restrict__ Q_old, double* __restrict__ Q_new,
                                                        const double* __restrict__ cache, double* __restrict__ sigma, GridDev g, double cgbeta,
                                                        double gamma, const double* __restrict__ scal) {
    if (scal) cgbeta = scal[1];
    const unsigned nvox = (unsigned)g.lnx * (unsigned)g.ny * (unsigned)g.nz;
    const int Ti[9] = {0, 1, 2, 6, 7, 8, 3, 4, 5};
    for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += gridDim.x * blockDim.x) {
        const unsigned row_ = v / (unsigned)g.nz;
        const size_t o = (size_t)row_ * g.nzp + (v - row_ * (unsigned)g.nz);
        double W[9], Fi[9], A[9], B[9];
#pragma unroll
        for (int d = 0; d < 9; d++) {
            const size_t oo = (size_t)d * g.plane + o;
            W[d] = UPDATE ? (R[oo] + cgbeta * Q_old[oo]) : Q_old[oo];
            Fi[d] = cache[oo];
        }
        if (UPDATE) {
#pragma unroll
            for (int d = 0; d < 9; d++) Q_new[(size_t)d * g.plane + o] = W[d];
        }
        const double a = cache[(size_t)9 * g.plane + o], b = cache[(size_t)10 * g.plane + o], c = cache[(size_t)11 * g.plane + o];
        nh_products(Fi, W, A, B);
        const double c_tr = b * (A[0] + A[1] + A[2]);
        const double gtr = gamma * (W[0] + W[1] + W[2]);
#pragma unroll
        for (int d = 0; d < 9; d++) {
            double s = a * W[d] + c_tr * Fi[Ti[d]] + c * B[d];
            if (d < 3 && gamma != 0) s += gtr;
            sigma[(size_t)d * g.plane + o] = s;
        }
    }
}

// the 9 components of grad_h u (epsOperatorStaggeredHyper fg:18784-18841, E = 0); two voxels (k, k+1) per thread so that X, R, Q and
// u move as 16-byte accesses; halos as k_eps (stencil.cu)
template <int DOT_ONLY>
__global__ void __launch_bounds__(256, 3) k_hyper_cg_u(const double* __restrict__ u, const double* __restrict__ Q, double* __restrict__ X,
                                                    double* __restrict__ R, double a, GridDev g, double* __restrict__ partials,
                                                    const double* __restrict__ halo_lo, const double* __restrict__ halo_hi, size_t hslot,
                                                    const double* __restrict__ scal) {
    if (!DOT_ONLY && scal) a = scal[2];
    const unsigned nzh = (unsigned)(g.nz + 1) / 2;
    const unsigned npairs = (unsigned)g.lnx * (unsigned)g.ny * nzh;
    const size_t us = 2 * (size_t)g.unzcs;
    const double* u0p = u;
    const double* u1p = u + g.uplane;
    const double* u2p = u + 2 * g.uplane;
    double acc = 0;
    for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < npairs; v += gridDim.x * blockDim.x) {
        const unsigned row_ = v / nzh;
        const int k = 2 * (int)(v - row_ * nzh);
        const int i = (int)(row_ / (unsigned)g.ny);
        const int j = (int)(row_ - (unsigned)i * (unsigned)g.ny);
        const bool second = k + 1 < g.nz;
        const int im = (i == 0) ? g.lnx - 1 : i - 1, ip = (i + 1 == g.lnx) ? 0 : i + 1;
        const int jm = (j == 0) ? g.ny - 1 : j - 1, jp = (j + 1 == g.ny) ? 0 : j + 1;
        const int km = (k == 0) ? g.nz - 1 : k - 1;
        const int kp2 = (k + 2 >= g.nz) ? k + 2 - g.nz : k + 2;
        const size_t rowo = (size_t)row_ * us;
        const size_t o = rowo + k;
        const size_t o_im = ((size_t)im * g.ny + j) * us + k, o_ip = ((size_t)ip * g.ny + j) * us + k;
        const size_t o_jm = ((size_t)i * g.ny + jm) * us + k, o_jp = ((size_t)i * g.ny + jp) * us + k;
        const size_t oh = (size_t)j * us + k;
        const bool lo_h = halo_lo != nullptr && i == 0, hi_h = halo_hi != nullptr && i + 1 == g.lnx;
#define LD2(ptr) (*reinterpret_cast<const double2*>(ptr))
        const double2 u0 = LD2(u0p + o), u1 = LD2(u1p + o), u2 = LD2(u2p + o);
        const double2 u0_ip = hi_h ? LD2(halo_hi + oh) : LD2(u0p + o_ip);
        const double2 u1_im = lo_h ? LD2(halo_lo + hslot + oh) : LD2(u1p + o_im);
        const double2 u2_im = lo_h ? LD2(halo_lo + 2 * hslot + oh) : LD2(u2p + o_im);
        const double2 u1_jp = LD2(u1p + o_jp), u0_jm = LD2(u0p + o_jm), u2_jm = LD2(u2p + o_jm);
#undef LD2
        const double u0_km = u0p[rowo + km], u1_km = u1p[rowo + km];
        const double u2_kp2 = u2p[rowo + kp2];
        const double u2_k1 = second ? u2.y : u2p[rowo];
        double e0[9], e1[9];
        e0[0] = (u0_ip.x - u0.x) * g.hx;   e1[0] = (u0_ip.y - u0.y) * g.hx;
        e0[1] = (u1_jp.x - u1.x) * g.hy;   e1[1] = (u1_jp.y - u1.y) * g.hy;
        e0[2] = (u2_k1 - u2.x) * g.hz;     e1[2] = (u2_kp2 - u2.y) * g.hz;
        e0[3] = (u1.x - u1_km) * g.hz;     e1[3] = (u1.y - u1.x) * g.hz;
        e0[4] = (u0.x - u0_km) * g.hz;     e1[4] = (u0.y - u0.x) * g.hz;
        e0[5] = (u0.x - u0_jm.x) * g.hy;   e1[5] = (u0.y - u0_jm.y) * g.hy;
        e0[6] = (u2.x - u2_jm.x) * g.hy;   e1[6] = (u2.y - u2_jm.y) * g.hy;
        e0[7] = (u2.x - u2_im.x) * g.hx;   e1[7] = (u2.y - u2_im.y) * g.hx;
        e0[8] = (u1.x - u1_im.x) * g.hx;   e1[8] = (u1.y - u1_im.y) * g.hx;
        const size_t eo = (size_t)row_ * g.nzp + k;
        double s0 = 0, s1 = 0;
#pragma unroll
        for (int d = 0; d < 9; d++) {
            const size_t oo = (size_t)d * g.plane + eo;
            const double2 q = *reinterpret_cast<const double2*>(Q + oo);
            if (DOT_ONLY) {
                s0 += q.x * (q.x - e0[d]);
                s1 += q.y * (q.y - e1[d]);
                continue;
            }
            double2 xv = *reinterpret_cast<double2*>(X + oo);
            double2 rv = *reinterpret_cast<double2*>(R + oo);
            xv.x = xv.x + a * q.x;
            xv.y = xv.y + a * q.y;
            rv.x = rv.x + (-a) * (q.x - e0[d]);
            rv.y = rv.y + (-a) * (q.y - e1[d]);
            if (!second) { xv.y = 0.0; rv.y = 0.0; }
            *reinterpret_cast<double2*>(X + oo) = xv;
            *reinterpret_cast<double2*>(R + oo) = rv;
            s0 += rv.x * rv.x;
            s1 += rv.y * rv.y;
        }
        acc += s0;
        if (second) acc += s1;
    }
    double vals[1] = {acc};
    block_reduce_store<1, 0>(vals, partials);
}

// ---- host side -------------------------------------------------------------------------------------------------------
int fgb_fused_nh_applicable(const fgb_ctx* ctx) {
    static const bool off = getenv("FGB_NO_FUSED_NH") != nullptr;
    if (off || ctx->dim != 9 || ctx->scheme != FGB_GAMMA_STAGGERED || ctx->dfg || ctx->nphases < 1 || ctx->mix != FGB_MIX_VOIGT) return 0;
    if (ctx->bc_active || ctx->bc_relax != 1.0) return 0;
    if (ctx->nranks > 1 && !ctx->nccl_comm) return 0;
    for (int p = 0; p < ctx->nphases; p++)
        if (ctx->laws[p].id != FGB_LAW_NH || !ctx->phi[p]) return 0;
    return 1;
}

// per-voxel tangent cache at the deformation gradient F (once per Newton iteration)
int fgb_k_nh_cache(fgb_ctx* ctx, const double* F, double mu0) {
    const GridDev& g = ctx->g;
    if (!ctx->nh_cache) {
        cudaError_t e = cudaMalloc(&ctx->nh_cache, sizeof(double) * g.plane * NH_CACHE_PLANES);
        if (e != cudaSuccess) { ctx->nh_cache = nullptr; return fgb_fail(ctx, FGB_ENOMEM, "cannot allocate the Neo-Hooke tangent cache (12 planes)"); }
        FGB_CUDA(ctx, cudaMemsetAsync(ctx->nh_cache, 0, sizeof(double) * g.plane * NH_CACHE_PLANES, ctx->stream));
    }
    NhPhases M;
    M.n = ctx->nphases;
    for (int p = 0; p < ctx->nphases; p++) { M.phi[p] = ctx->phi[p]; M.mu[p] = ctx->laws[p].p[0]; M.lam[p] = ctx->laws[p].p[1]; }
    const size_t nvox = (size_t)g.lnx * g.ny * g.nz;
    const unsigned grid = fgb_wave_grid(ctx, (const void*)k_nh_cache, 256, nvox, (size_t)ctx->sm_count * 16);
    ProfScope ps(ctx, "nh_tangent_cache");
    k_nh_cache<<<grid, 256, 0, ctx->stream>>>(F, ctx->nh_cache, g, M, -2 * mu0, ctx->d_flag);
    FGB_CHECK_LAUNCH(ctx, "k_nh_cache");
    return FGB_OK;
}

int fgb_k_nh_dir_tangent(fgb_ctx* ctx, const double* R, double cgbeta, const double* Q_old, double* Q_new, double* sigma, double lambda0) {
    const GridDev& g = ctx->g;
    const size_t nvox = (size_t)g.lnx * g.ny * g.nz;
    const double* scal = (R && ctx->cg_dev) ? ctx->d_scalars : nullptr;
    ProfScope ps(ctx, R ? "nh_dir_tangent" : "nh_tangent");
    if (R) {
        const unsigned grid = fgb_wave_grid(ctx, (const void*)k_nh_dir_tangent<1>, 256, nvox, (size_t)ctx->sm_count * 16);
        k_nh_dir_tangent<1><<<grid, 256, 0, ctx->stream>>>(R, Q_old, Q_new, ctx->nh_cache, sigma, g, cgbeta, -lambda0, scal);
    } else {
        const unsigned grid = fgb_wave_grid(ctx, (const void*)k_nh_dir_tangent<0>, 256, nvox, (size_t)ctx->sm_count * 16);
        k_nh_dir_tangent<0><<<grid, 256, 0, ctx->stream>>>(nullptr, Q_old, nullptr, ctx->nh_cache, sigma, g, 0.0, -lambda0, nullptr);
    }
    FGB_CHECK_LAUNCH(ctx, "k_nh_dir_tangent");
    return FGB_OK;
}

int fgb_k_hyper_cg_u(fgb_ctx* ctx, bool dot_only, double* X, double* R, const double* Q, double a, double* out) {
    const GridDev& g = ctx->g;
    const size_t npairs = (size_t)g.lnx * g.ny * ((g.nz + 1) / 2);
    const unsigned grid = fgb_wave_grid(ctx, dot_only ? (const void*)k_hyper_cg_u<1> : (const void*)k_hyper_cg_u<0>, 256, npairs, ctx->red_blocks);
    {
        ProfScope ps(ctx, dot_only ? "eps_dot_implicit" : "cg_update_implicit");
        const double* lo = (ctx->nranks > 1) ? ctx->halo : nullptr;
        const double* hi = (ctx->nranks > 1) ? ctx->halo + 3 * ctx->halo_slot : nullptr;
        const double* scal = ctx->cg_dev ? ctx->d_scalars : nullptr;
        if (dot_only) k_hyper_cg_u<1><<<grid, 256, 0, ctx->stream>>>(ctx->ubuf, Q, X, R, a, g, ctx->d_partials, lo, hi, ctx->halo_slot, scal);
        else k_hyper_cg_u<0><<<grid, 256, 0, ctx->stream>>>(ctx->ubuf, Q, X, R, a, g, ctx->d_partials, lo, hi, ctx->halo_slot, scal);
        FGB_CHECK_LAUNCH(ctx, "k_hyper_cg_u");
    }
    int rc = fgb_reduce_finish(ctx, grid, 1, 0, out);
    if (rc) return rc;
    out[0] /= (double)g.nx * g.ny * g.nz;
    return FGB_OK;
}
