// Staggered-grid finite-difference operators.
//   div : divOperatorStaggered fg:18853-18908, ...Heat fg:18914-18968, ...Hyper fg:19016-19071
//   eps : epsOperatorStaggered fg:18614-18692, ...Heat fg:18697-18757, ...Hyper fg:18763-18846
// D+ a(i) = (a(i+1)-a(i))*h, D- a(i) = (a(i)-a(i-1))*h, periodic (offset tables fg:14867-14891), h = n/L.
// Unlike the reference (in place, barriers fg:18642) the rhs / displacement lives in its own buffer `u`,
// so there is no in-place hazard.  With an x-slab partition the i-1 / i+1 planes of the neighbour ranks
// come from the halo buffers filled by comm.cu.
#include "fgb_internal.h"

struct StencilArgs {
    GridDev g;
    // halo planes (ny*nzp doubles per component) for i = -1 and i = lnx; null when nranks == 1
    const double* halo_lo;   // components packed as needed by the kernel
    const double* halo_hi;
};

__device__ __forceinline__ const double* plane_ptr(const double* f, const GridDev& g, int c) { return f + (size_t)c * g.plane; }

// value of component plane `p` at (i+di, j, k) with periodic wrap in x or halo access
__device__ __forceinline__ double at_x(const double* p, const GridDev& g, int i, int j, int k, int di, const double* halo_lo,
                                       const double* halo_hi, size_t rs) {
    int ii = i + di;
    if (halo_lo == nullptr) {
        if (ii < 0) ii += g.lnx;
        if (ii >= g.lnx) ii -= g.lnx;
        return p[((size_t)ii * g.ny + j) * rs + k];
    }
    if (ii < 0) return halo_lo[(size_t)j * rs + k];
    if (ii >= g.lnx) return halo_hi[(size_t)j * rs + k];
    return p[((size_t)ii * g.ny + j) * rs + k];
}

template <int D>
__global__ void __launch_bounds__(256) k_div(const double* __restrict__ tau, double* __restrict__ u, GridDev g,
                                             const double* __restrict__ halo_lo, const double* __restrict__ halo_hi, size_t hp) {
    const unsigned nvox = (unsigned)g.lnx * (unsigned)g.ny * (unsigned)g.nz;
    for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += gridDim.x * blockDim.x) {
        const unsigned row_ = v / (unsigned)g.nz;
        const int k = (int)(v - row_ * (unsigned)g.nz);
        const int i = (int)(row_ / (unsigned)g.ny);
        const int j = (int)(row_ - (unsigned)i * (unsigned)g.ny);
        const size_t o = ((size_t)i * g.ny + j) * g.nzp + k;
        const size_t uo = ((size_t)i * g.ny + j) * (2 * (size_t)g.unzcs) + k;
        const int jp = (j + 1 == g.ny) ? 0 : j + 1, jm = (j == 0) ? g.ny - 1 : j - 1;
        const int kp = (k + 1 == g.nz) ? 0 : k + 1, km = (k == 0) ? g.nz - 1 : k - 1;
        const size_t o_jp = ((size_t)i * g.ny + jp) * g.nzp + k, o_jm = ((size_t)i * g.ny + jm) * g.nzp + k;
        const size_t o_kp = ((size_t)i * g.ny + j) * g.nzp + kp, o_km = ((size_t)i * g.ny + j) * g.nzp + km;
#define T_(c) plane_ptr(tau, g, c)
        if (D == 3) {
            // halo_lo: [tau0 at i=-1]
            double f = (T_(0)[o] - at_x(T_(0), g, i, j, k, -1, halo_lo, halo_hi, (size_t)g.nzp)) * g.hx;
            f += (T_(1)[o] - T_(1)[o_jm]) * g.hy;
            f += (T_(2)[o] - T_(2)[o_km]) * g.hz;
            u[uo] = f;
        } else {
            // shear operand indices: elasticity (5,4 | 5,3 | 4,3), hyper (5,4 | 8,3 | 7,6)
            const int c1x = (D == 6) ? 5 : 8, c2x = (D == 6) ? 4 : 7, c2y = (D == 6) ? 3 : 6;
            // halo_lo: [tau0 at -1]; halo_hi: [tau(c1x) at lnx, tau(c2x) at lnx]
            const double f0 = (T_(0)[o] - at_x(T_(0), g, i, j, k, -1, halo_lo, halo_hi, (size_t)g.nzp)) * g.hx + (T_(5)[o_jp] - T_(5)[o]) * g.hy +
                              (T_(4)[o_kp] - T_(4)[o]) * g.hz;
            const double f1 = (at_x(T_(c1x), g, i, j, k, +1, halo_lo, halo_hi, (size_t)g.nzp) - T_(c1x)[o]) * g.hx + (T_(1)[o] - T_(1)[o_jm]) * g.hy +
                              (T_(3)[o_kp] - T_(3)[o]) * g.hz;
            const double f2 = (at_x(T_(c2x), g, i, j, k, +1, halo_lo, halo_hi ? halo_hi + hp : nullptr, (size_t)g.nzp) - T_(c2x)[o]) * g.hx +
                              (T_(c2y)[o_jp] - T_(c2y)[o]) * g.hy + (T_(2)[o] - T_(2)[o_km]) * g.hz;
            u[uo] = f0;
            u[g.uplane + uo] = f1;
            u[2 * g.uplane + uo] = f2;
        }
#undef T_
    }
}

struct Const9 {
    double v[9];
};

template <int D>
__global__ void __launch_bounds__(256) k_eps(const double* __restrict__ u, double* __restrict__ eta, GridDev g, Const9 E,
                                             const double* __restrict__ halo_lo, const double* __restrict__ halo_hi, size_t hp) {
    const unsigned nvox = (unsigned)g.lnx * (unsigned)g.ny * (unsigned)g.nz;
    for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += gridDim.x * blockDim.x) {
        const unsigned row_ = v / (unsigned)g.nz;
        const int k = (int)(v - row_ * (unsigned)g.nz);
        const int i = (int)(row_ / (unsigned)g.ny);
        const int j = (int)(row_ - (unsigned)i * (unsigned)g.ny);
        const size_t eo = ((size_t)i * g.ny + j) * g.nzp + k;           // output (field layout)
        const size_t us = 2 * (size_t)g.unzcs;                            // u row stride
        const int jp = (j + 1 == g.ny) ? 0 : j + 1, jm = (j == 0) ? g.ny - 1 : j - 1;
        const int kp = (k + 1 == g.nz) ? 0 : k + 1, km = (k == 0) ? g.nz - 1 : k - 1;
        const size_t o = ((size_t)i * g.ny + j) * us + k;
        const size_t o_jp = ((size_t)i * g.ny + jp) * us + k, o_jm = ((size_t)i * g.ny + jm) * us + k;
        const size_t o_kp = ((size_t)i * g.ny + j) * us + kp, o_km = ((size_t)i * g.ny + j) * us + km;
#define U_(c) (u + (size_t)(c)*g.uplane)
        // halo_lo: [u0,u1,u2 at i=-1], halo_hi: [u0 at i=lnx]
        if (D == 3) {
            const double u0 = U_(0)[o];
            eta[eo] = E.v[0] + (at_x(U_(0), g, i, j, k, +1, halo_lo, halo_hi, us) - u0) * g.hx;
            eta[g.plane + eo] = E.v[1] + (U_(0)[o_jp] - u0) * g.hy;
            eta[2 * g.plane + eo] = E.v[2] + (U_(0)[o_kp] - u0) * g.hz;
        } else {
            const double u0 = U_(0)[o], u1 = U_(1)[o], u2 = U_(2)[o];
            const double u0_xm = at_x(U_(0), g, i, j, k, -1, halo_lo, halo_hi, us);
            const double u1_xm = at_x(U_(1), g, i, j, k, -1, halo_lo ? halo_lo + hp : nullptr, halo_hi, us);
            const double u2_xm = at_x(U_(2), g, i, j, k, -1, halo_lo ? halo_lo + 2 * hp : nullptr, halo_hi, us);
            (void)u0_xm;
            const double e0 = E.v[0] + (at_x(U_(0), g, i, j, k, +1, halo_lo, halo_hi, us) - u0) * g.hx;
            const double e1 = E.v[1] + (U_(1)[o_jp] - u1) * g.hy;
            const double e2 = E.v[2] + (U_(2)[o_kp] - u2) * g.hz;
            eta[eo] = e0;
            eta[g.plane + eo] = e1;
            eta[2 * g.plane + eo] = e2;
            if (D == 6) {
                eta[3 * g.plane + eo] = E.v[3] + 0.5 * ((u2 - U_(2)[o_jm]) * g.hy + (u1 - U_(1)[o_km]) * g.hz);
                eta[4 * g.plane + eo] = E.v[4] + 0.5 * ((u2 - u2_xm) * g.hx + (u0 - U_(0)[o_km]) * g.hz);
                eta[5 * g.plane + eo] = E.v[5] + 0.5 * ((u1 - u1_xm) * g.hx + (u0 - U_(0)[o_jm]) * g.hy);
            } else {
                eta[3 * g.plane + eo] = E.v[3] + (u1 - U_(1)[o_km]) * g.hz;
                eta[4 * g.plane + eo] = E.v[4] + (u0 - U_(0)[o_km]) * g.hz;
                eta[5 * g.plane + eo] = E.v[5] + (u0 - U_(0)[o_jm]) * g.hy;
                eta[6 * g.plane + eo] = E.v[6] + (u2 - U_(2)[o_jm]) * g.hy;
                eta[7 * g.plane + eo] = E.v[7] + (u2 - u2_xm) * g.hx;
                eta[8 * g.plane + eo] = E.v[8] + (u1 - u1_xm) * g.hx;
            }
        }
#undef U_
    }
}


// halo layout is owned by comm.cu: ctx->halo = [3 lo slots][3 hi slots], each slot ctx->halo_slot doubles
static void halo_ptrs(fgb_ctx* ctx, const double** lo, const double** hi) {
    if (ctx->nranks > 1 && ctx->halo) {
        *lo = ctx->halo;
        *hi = ctx->halo + 3 * ctx->halo_slot;
    } else {
        *lo = nullptr;
        *hi = nullptr;
    }
}

int fgb_k_div(fgb_ctx* ctx, const double* tau, double* u) {
    ctx->implicit_w_of = -1;          // the u buffer is overwritten
    const GridDev& g = ctx->g;
    const size_t nvox = (size_t)g.lnx * g.ny * g.nz;
    const double *lo, *hi;
    halo_ptrs(ctx, &lo, &hi);
    ProfScope ps(ctx, "div_staggered");
    const void* kp = ctx->dim == 3 ? (const void*)k_div<3> : ctx->dim == 6 ? (const void*)k_div<6> : (const void*)k_div<9>;
    const unsigned grid = fgb_wave_grid(ctx, kp, 256, nvox, (size_t)ctx->sm_count * 16);
    if (ctx->dim == 3) k_div<3><<<grid, 256, 0, ctx->stream>>>(tau, u, g, lo, hi, ctx->halo_slot);
    else if (ctx->dim == 6) k_div<6><<<grid, 256, 0, ctx->stream>>>(tau, u, g, lo, hi, ctx->halo_slot);
    else k_div<9><<<grid, 256, 0, ctx->stream>>>(tau, u, g, lo, hi, ctx->halo_slot);
    FGB_CHECK_LAUNCH(ctx, "k_div");
    return FGB_OK;
}

int fgb_k_eps(fgb_ctx* ctx, const double* u, double* eta, const double* Econst) {
    const GridDev& g = ctx->g;
    const size_t nvox = (size_t)g.lnx * g.ny * g.nz;
    Const9 E;
    for (int i = 0; i < 9; i++) E.v[i] = (i < ctx->dim) ? Econst[i] : 0.0;
    const double *lo, *hi;
    halo_ptrs(ctx, &lo, &hi);
    ProfScope ps(ctx, "eps_staggered");
    const void* kp = ctx->dim == 3 ? (const void*)k_eps<3> : ctx->dim == 6 ? (const void*)k_eps<6> : (const void*)k_eps<9>;
    const unsigned grid = fgb_wave_grid(ctx, kp, 256, nvox, (size_t)ctx->sm_count * 16);
    if (ctx->dim == 3) k_eps<3><<<grid, 256, 0, ctx->stream>>>(u, eta, g, E, lo, hi, ctx->halo_slot);
    else if (ctx->dim == 6) k_eps<6><<<grid, 256, 0, ctx->stream>>>(u, eta, g, E, lo, hi, ctx->halo_slot);
    else k_eps<9><<<grid, 256, 0, ctx->stream>>>(u, eta, g, E, lo, hi, ctx->halo_slot);
    FGB_CHECK_LAUNCH(ctx, "k_eps");
    return FGB_OK;
}

// divVector fg:19983-20003: b = alpha * sum_a (f_a(x) - f_a(x + e_a)) / h_a of the 3-component buffer u (forward differences with the
// sign of the reference), written as one component in the u layout.  halo_hi: [u0 at i = lnx].
__global__ void __launch_bounds__(256) k_div_vector(const double* __restrict__ u, double* __restrict__ b, GridDev g, double alpha,
                                                    const double* __restrict__ halo_hi) {
    const unsigned nvox = (unsigned)g.lnx * (unsigned)g.ny * (unsigned)g.nz;
    const size_t us = 2 * (size_t)g.unzcs;
    const double chx = alpha * g.hx, chy = alpha * g.hy, chz = alpha * g.hz;
    for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += gridDim.x * blockDim.x) {
        const unsigned row_ = v / (unsigned)g.nz;
        const int k = (int)(v - row_ * (unsigned)g.nz);
        const int i = (int)(row_ / (unsigned)g.ny);
        const int j = (int)(row_ - (unsigned)i * (unsigned)g.ny);
        const int jp = (j + 1 == g.ny) ? 0 : j + 1;
        const int kp = (k + 1 == g.nz) ? 0 : k + 1;
        const size_t o = ((size_t)i * g.ny + j) * us + k;
        const double* u0 = u;
        const double* u1 = u + g.uplane;
        const double* u2 = u + 2 * g.uplane;
        const double u0_xp = at_x(u0, g, i, j, k, +1, halo_hi, halo_hi, us);
        b[o] = (u0[o] - u0_xp) * chx + (u1[o] - u1[((size_t)i * g.ny + jp) * us + k]) * chy + (u2[o] - u2[((size_t)i * g.ny + j) * us + kp]) * chz;
    }
}

int fgb_k_div_vector(fgb_ctx* ctx, const double* u, double* b, double alpha) {
    const GridDev& g = ctx->g;
    const size_t nvox = (size_t)g.lnx * g.ny * g.nz;
    const double *lo, *hi;
    halo_ptrs(ctx, &lo, &hi);
    ProfScope ps(ctx, "div_vector");
    const unsigned grid = fgb_wave_grid(ctx, (const void*)k_div_vector, 256, nvox, (size_t)ctx->sm_count * 16);
    k_div_vector<<<grid, 256, 0, ctx->stream>>>(u, b, g, alpha, hi);
    FGB_CHECK_LAUNCH(ctx, "k_div_vector");
    return FGB_OK;
}

// mxpyTensor fg:20590-20596: r = -(x + y) over a whole component plane (padding included, as in the reference)
__global__ void __launch_bounds__(256) k_mxpy(double* __restrict__ r, const double* __restrict__ x, const double* __restrict__ y, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) r[i] = -(x[i] + y[i]);
}

int fgb_k_mxpy(fgb_ctx* ctx, double* r, const double* x, const double* y) {
    const size_t n = ctx->g.plane;
    const unsigned grid = fgb_wave_grid(ctx, (const void*)k_mxpy, 256, n, (size_t)ctx->sm_count * 16);
    k_mxpy<<<grid, 256, 0, ctx->stream>>>(r, x, y, n);
    FGB_CHECK_LAUNCH(ctx, "k_mxpy");
    return FGB_OK;
}

// ---- doubly fine grid transfer (half_staggered / full_staggered schemes) ------------------------------------------------
// component-wise shifts of the staggered positions, stored component order 11,22,33,23,13,12,32,31,21 (fg:14232-14234)
__constant__ int c_dfg_si[9] = {0, 0, 0, 0, 1, 1, 0, 1, 1};
__constant__ int c_dfg_sj[9] = {0, 0, 0, 1, 0, 1, 1, 0, 1};
__constant__ int c_dfg_sk[9] = {0, 0, 0, 1, 1, 0, 1, 1, 0};

// prolongate_to_dfg fg:14216-14268: fine(i,j,k) = coarse(((i + si) mod fnx)/2, ((j + sj) mod fny)/2, ((k + sk) mod fnz)/2) per component
__global__ void __launch_bounds__(256) k_prolongate(const double* __restrict__ c, double* __restrict__ f, GridDev gc, GridDev gf, int dim) {
    const unsigned nvox = (unsigned)gf.lnx * (unsigned)gf.ny * (unsigned)gf.nz;
    for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += gridDim.x * blockDim.x) {
        const unsigned row_ = v / (unsigned)gf.nz;
        const int k = (int)(v - row_ * (unsigned)gf.nz);
        const int i = (int)(row_ / (unsigned)gf.ny);
        const int j = (int)(row_ - (unsigned)i * (unsigned)gf.ny);
        const size_t o = (size_t)row_ * gf.nzp + k;
        for (int d = 0; d < dim; d++) {
            const int ii = ((i + gf.lnx + c_dfg_si[d]) % gf.lnx) / 2;
            const int jj = ((j + gf.ny + c_dfg_sj[d]) % gf.ny) / 2;
            const int kk = ((k + gf.nz + c_dfg_sk[d]) % gf.nz) / 2;
            f[(size_t)d * gf.plane + o] = c[(size_t)d * gc.plane + ((size_t)ii * gc.ny + jj) * gc.nzp + kk];
        }
    }
}

// restrict_from_dfg fg:14273-14339: coarse(i,j,k) = mean of the 8 fine values at (2i + a - si, 2j + b - sj, 2k + c - sk) mod n, a,b,c in {0,1}
__global__ void __launch_bounds__(256) k_restrict(const double* __restrict__ f, double* __restrict__ c, GridDev gc, GridDev gf, int dim) {
    const unsigned nvox = (unsigned)gc.lnx * (unsigned)gc.ny * (unsigned)gc.nz;
    for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += gridDim.x * blockDim.x) {
        const unsigned row_ = v / (unsigned)gc.nz;
        const int k = (int)(v - row_ * (unsigned)gc.nz);
        const int i = (int)(row_ / (unsigned)gc.ny);
        const int j = (int)(row_ - (unsigned)i * (unsigned)gc.ny);
        const size_t o = (size_t)row_ * gc.nzp + k;
        for (int d = 0; d < dim; d++) {
            const size_t i0 = (size_t)((2 * i + gf.lnx - c_dfg_si[d]) % gf.lnx), i1 = (size_t)((2 * i + 1 + gf.lnx - c_dfg_si[d]) % gf.lnx);
            const size_t j0 = (size_t)((2 * j + gf.ny - c_dfg_sj[d]) % gf.ny), j1 = (size_t)((2 * j + 1 + gf.ny - c_dfg_sj[d]) % gf.ny);
            const size_t k0 = (size_t)((2 * k + gf.nz - c_dfg_sk[d]) % gf.nz), k1 = (size_t)((2 * k + 1 + gf.nz - c_dfg_sk[d]) % gf.nz);
            const double* s = f + (size_t)d * gf.plane;
#define FQ(a, b, cc) s[((a)*gf.ny + (b)) * gf.nzp + (cc)]
            c[(size_t)d * gc.plane + o] = 0.125 * (FQ(i0, j0, k0) + FQ(i1, j0, k0) + FQ(i0, j1, k0) + FQ(i1, j1, k0) + FQ(i0, j0, k1) +
                                                   FQ(i1, j0, k1) + FQ(i0, j1, k1) + FQ(i1, j1, k1));
#undef FQ
        }
    }
}

int fgb_k_prolongate(fgb_ctx* ctx, const double* coarse, double* fine) {
    const GridDev& gf = ctx->gf;
    const size_t nvox = (size_t)gf.lnx * gf.ny * gf.nz;
    ProfScope ps(ctx, "prolongate_to_dfg");
    const unsigned grid = fgb_wave_grid(ctx, (const void*)k_prolongate, 256, nvox, (size_t)ctx->sm_count * 16);
    k_prolongate<<<grid, 256, 0, ctx->stream>>>(coarse, fine, ctx->g, gf, ctx->dim);
    FGB_CHECK_LAUNCH(ctx, "k_prolongate");
    return FGB_OK;
}

int fgb_k_restrict(fgb_ctx* ctx, const double* fine, double* coarse) {
    const GridDev& g = ctx->g;
    const size_t nvox = (size_t)g.lnx * g.ny * g.nz;
    ProfScope ps(ctx, "restrict_from_dfg");
    const unsigned grid = fgb_wave_grid(ctx, (const void*)k_restrict, 256, nvox, (size_t)ctx->sm_count * 16);
    k_restrict<<<grid, 256, 0, ctx->stream>>>(fine, coarse, g, ctx->gf, ctx->dim);
    FGB_CHECK_LAUNCH(ctx, "k_restrict");
    return FGB_OK;
}

// initFullStageredRawPhases fg:17648-17680: piecewise constant injection of one coarse plane, fine(i,j,k) = coarse(i/2, j/2, k/2)
__global__ void __launch_bounds__(256) k_inject(const double* __restrict__ c, double* __restrict__ f, GridDev gc, GridDev gf) {
    const unsigned nvox = (unsigned)gf.lnx * (unsigned)gf.ny * (unsigned)gf.nz;
    for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += gridDim.x * blockDim.x) {
        const unsigned row_ = v / (unsigned)gf.nz;
        const int k = (int)(v - row_ * (unsigned)gf.nz);
        const int i = (int)(row_ / (unsigned)gf.ny);
        const int j = (int)(row_ - (unsigned)i * (unsigned)gf.ny);
        f[(size_t)row_ * gf.nzp + k] = c[((size_t)(i / 2) * gc.ny + (j / 2)) * gc.nzp + (k / 2)];
    }
}

int fgb_k_inject_phase(fgb_ctx* ctx, const double* coarse, double* fine) {
    const GridDev& gf = ctx->gf;
    const size_t nvox = (size_t)gf.lnx * gf.ny * gf.nz;
    const unsigned grid = fgb_wave_grid(ctx, (const void*)k_inject, 256, nvox, (size_t)ctx->sm_count * 16);
    k_inject<<<grid, 256, 0, ctx->stream>>>(coarse, fine, ctx->g, gf);
    FGB_CHECK_LAUNCH(ctx, "k_inject");
    return FGB_OK;
}
