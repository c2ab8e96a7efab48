// TensorField BLAS-1 (fg:9799-10066), Voigt-weighted inner products (fg:20871-21036), per-component
// reductions (fg:10088-10208) and the fused CG update sweep (fg:23221-23245).
// All sweeps run over voxels (padding skipped), are vectorised along z where alignment allows and
// reduce deterministically (warp shuffle -> block -> fixed-order finish).
#include "fgb_internal.h"
#include "reduce.cuh"

unsigned fgb_wave_grid(fgb_ctx* ctx, const void* kernel, int block, size_t n, size_t cap) {
    size_t b = (n + block - 1) / block;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    auto it = ctx->occupancy.find(kernel);
    int occ;
    if (it == ctx->occupancy.end()) {
        occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, block, 0) != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 0; }
        ctx->occupancy[kernel] = occ;
    } else occ = it->second;
    const size_t wave = (size_t)occ * ctx->sm_count;
    if (wave > 0 && b > wave) b = (b / wave) * wave;
    return (unsigned)b;
}

static unsigned grid_for(const fgb_ctx* ctx, size_t n, int block) {
    size_t b = (n + block - 1) / block;
    size_t cap = (size_t)ctx->red_blocks;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

struct Const9 {
    double v[9];
};

// iterate over voxel pairs (k, k+1) so that loads are 128-bit (nzp is even, rows are 16-byte aligned)
#define FGB_VOXEL_PAIR_LOOP(g)                                                                                  \
    const int nzh = (g.nz + 1) / 2;                                                                             \
    const size_t npairs = (size_t)g.lnx * g.ny * nzh;                                                           \
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < npairs; v += (size_t)gridDim.x * blockDim.x)

#define FGB_PAIR_INDEX(g)                                             \
    const int kh = (int)(v % nzh);                                    \
    const size_t row = v / nzh;                                       \
    const size_t o = row * g.nzp + 2 * (size_t)kh;                    \
    const bool second = (2 * kh + 1) < g.nz;

template <int D>
__global__ void __launch_bounds__(256) k_set_constant(double* __restrict__ f, GridDev g, Const9 c, int add) {
    // like the reference this also writes the padding (fg:10045)
    const size_t n2 = g.plane / 2;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n2; v += (size_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int d = 0; d < D; d++) {
            double2* p = reinterpret_cast<double2*>(f + (size_t)d * g.plane) + v;
            if (add) {
                double2 x = *p;
                x.x += c.v[d];
                x.y += c.v[d];
                *p = x;
            } else {
                *p = make_double2(c.v[d], c.v[d]);
            }
        }
    }
}

// r = x + a*(y - z) (z may be null -> r = x + a*y); also used for copy (a = 0)
template <int MODE>
__global__ void __launch_bounds__(256) k_axpy(double* __restrict__ r, const double* __restrict__ x, double a,
                                              const double* __restrict__ y, const double* __restrict__ z, size_t n2,
                                              const double* __restrict__ scal) {
    if (scal) a = *scal;          // device-resident CG scalar (fgb_cgdev_*)
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n2; v += (size_t)gridDim.x * blockDim.x) {
        const double2 xv = reinterpret_cast<const double2*>(x)[v];
        double2 o;
        if (MODE == 0) {
            o = xv;
        } else if (MODE == 1) {
            const double2 yv = reinterpret_cast<const double2*>(y)[v];
            o = make_double2(xv.x + a * yv.x, xv.y + a * yv.y);
        } else {
            const double2 yv = reinterpret_cast<const double2*>(y)[v];
            const double2 zv = reinterpret_cast<const double2*>(z)[v];
            o = make_double2(xv.x + a * (yv.x - zv.x), xv.y + a * (yv.y - zv.y));
        }
        reinterpret_cast<double2*>(r)[v] = o;
    }
}

template <int D>
__global__ void __launch_bounds__(256) k_adjust_residual(double* __restrict__ r, GridDev g, Const9 E, const double* __restrict__ z) {
    const size_t n2 = g.plane / 2;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n2; v += (size_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int d = 0; d < D; d++) {
            double2* p = reinterpret_cast<double2*>(r + (size_t)d * g.plane) + v;
            const double2 zv = reinterpret_cast<const double2*>(z + (size_t)d * g.plane)[v];
            double2 x = *p;
            x.x += E.v[d] - zv.x;
            x.y += E.v[d] - zv.y;
            *p = x;
        }
    }
}

// tau = C0 : eps (calcStressConst fg:17973), over the whole padded array like the reference
template <int D>
__global__ void __launch_bounds__(256) k_stress_const(const double* __restrict__ e, double* __restrict__ t, GridDev g, double two_mu,
                                                       double lambda) {
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < g.plane; v += (size_t)gridDim.x * blockDim.x) {
        double ev[D];
#pragma unroll
        for (int d = 0; d < D; d++) ev[d] = e[(size_t)d * g.plane + v];
        const double ltr = lambda * (ev[0] + ev[1] + ev[2]);
#pragma unroll
        for (int d = 0; d < D; d++) t[(size_t)d * g.plane + v] = ev[d] * two_mu + (d < 3 ? ltr : 0.0);
    }
}

// Voigt-weighted a:(b-c) summed over voxels
template <int D, int HASC>
__global__ void __launch_bounds__(256) k_inner(const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ c,
                                               GridDev g, double* __restrict__ partials) {
    double s = 0;
    FGB_VOXEL_PAIR_LOOP(g) {
        FGB_PAIR_INDEX(g)
        double s0 = 0, s1 = 0;
#pragma unroll
        for (int d = 0; d < D; d++) {
            const double2 av = *reinterpret_cast<const double2*>(a + (size_t)d * g.plane + o);
            double2 bv = *reinterpret_cast<const double2*>(b + (size_t)d * g.plane + o);
            if (HASC) {
                const double2 cv = *reinterpret_cast<const double2*>(c + (size_t)d * g.plane + o);
                bv.x -= cv.x;
                bv.y -= cv.y;
            }
            const double w = (D == 6 && d >= 3) ? 2.0 : 1.0;
            s0 += w * av.x * bv.x;
            s1 += w * av.y * bv.y;
        }
        s += s0;
        if (second) s += s1;
    }
    double vals[1] = {s};
    block_reduce_store<1, 0>(vals, partials);
}

// per-component sum of a_j*b_j (b == null: sum of a_j)
template <int D, int MEAN>
__global__ void __launch_bounds__(256) k_component_dot(const double* __restrict__ a, const double* __restrict__ b, GridDev g,
                                                       double* __restrict__ partials) {
    double s[D];
#pragma unroll
    for (int d = 0; d < D; d++) s[d] = 0;
    FGB_VOXEL_PAIR_LOOP(g) {
        FGB_PAIR_INDEX(g)
#pragma unroll
        for (int d = 0; d < D; d++) {
            const double2 av = *reinterpret_cast<const double2*>(a + (size_t)d * g.plane + o);
            if (MEAN) {
                s[d] += av.x;
                if (second) s[d] += av.y;
            } else {
                const double2 bv = *reinterpret_cast<const double2*>(b + (size_t)d * g.plane + o);
                s[d] += av.x * bv.x;
                if (second) s[d] += av.y * bv.y;
            }
        }
    }
    block_reduce_store<D, 0>(s, partials);
}

// fused CG update: x += a*p ; r -= a*(p - w) ; delta = <r,r>   (fg:23221, fg:23237, fg:23240)
template <int D>
__global__ void __launch_bounds__(256) k_cg_update(double* __restrict__ x, double* __restrict__ r, const double* __restrict__ p,
                                                   const double* __restrict__ w, double a, GridDev g, double* __restrict__ partials,
                                                   const double* __restrict__ scal) {
    if (scal) a = scal[2];
    double s = 0;
    FGB_VOXEL_PAIR_LOOP(g) {
        FGB_PAIR_INDEX(g)
        double s0 = 0, s1 = 0;
#pragma unroll
        for (int d = 0; d < D; d++) {
            const size_t oo = (size_t)d * g.plane + o;
            const double2 pv = *reinterpret_cast<const double2*>(p + oo);
            const double2 wv = *reinterpret_cast<const double2*>(w + oo);
            double2 xv = *reinterpret_cast<double2*>(x + oo);
            double2 rv = *reinterpret_cast<double2*>(r + oo);
            xv.x = xv.x + a * pv.x;
            xv.y = xv.y + a * pv.y;
            rv.x = rv.x + (-a) * (pv.x - wv.x);
            rv.y = rv.y + (-a) * (pv.y - wv.y);
            *reinterpret_cast<double2*>(x + oo) = xv;
            *reinterpret_cast<double2*>(r + oo) = rv;
            const double wgt = (D == 6 && d >= 3) ? 2.0 : 1.0;
            s0 += wgt * rv.x * rv.x;
            s1 += wgt * rv.y * rv.y;
        }
        s += s0;
        if (second) s += s1;
    }
    double vals[1] = {s};
    block_reduce_store<1, 0>(vals, partials);
}

// combine the per-block partials in a fixed order
__global__ void k_reduce_finish(const double* __restrict__ partials, int nblocks, int nvals, int op, double* __restrict__ out) {
    const int v = blockIdx.x;
    if (v >= nvals) return;
    __shared__ double sh[256];
    double acc = (op == 0) ? 0.0 : (op == 1 ? INFINITY : -INFINITY);
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) {
        const double x = partials[(size_t)b * FGB_RED_MAXV + v];
        acc = (op == 0) ? acc + x : (op == 1 ? fmin(acc, x) : fmax(acc, x));
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            const double x = sh[threadIdx.x + s];
            sh[threadIdx.x] = (op == 0) ? sh[threadIdx.x] + x : (op == 1 ? fmin(sh[threadIdx.x], x) : fmax(sh[threadIdx.x], x));
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[v] = sh[0];
}

int fgb_reduce_finish(fgb_ctx* ctx, int nblocks, int nvals, int op, double* host_out) {
    k_reduce_finish<<<nvals, 256, 0, ctx->stream>>>(ctx->d_partials, nblocks, nvals, op, ctx->d_result);
    FGB_CHECK_LAUNCH(ctx, "k_reduce_finish");
    if (ctx->reduce_on_device) {
        // fgb_cgdev_*: the sum stays on the device (d_result, or d_gather with one entry per rank); no host synchronisation
        for (int i = 0; i < nvals; i++) host_out[i] = 0.0;
        return ctx->nranks > 1 ? fgb_allgather_dev(ctx, nvals) : FGB_OK;
    }
    // slab partition: the per-rank results are gathered on the device and combined in rank order (one host synchronisation)
    if (ctx->nranks > 1) return fgb_allreduce_host(ctx, host_out, nvals, op);
    FGB_CUDA(ctx, cudaMemcpyAsync(ctx->h_result, ctx->d_result, sizeof(double) * nvals, cudaMemcpyDeviceToHost, ctx->stream));
    FGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < nvals; i++) host_out[i] = ctx->h_result[i];
    return FGB_OK;
}

#define DISPATCH_D(ctx, CALL3, CALL6, CALL9) \
    do {                                     \
        if ((ctx)->dim == 3) { CALL3; }      \
        else if ((ctx)->dim == 6) { CALL6; } \
        else { CALL9; }                      \
    } while (0)

int fgb_k_set_constant(fgb_ctx* ctx, double* f, const double* c, int add) {
    Const9 C;
    for (int i = 0; i < 9; i++) C.v[i] = i < ctx->dim ? c[i] : 0;
    const unsigned grid = grid_for(ctx, ctx->g.plane / 2, 256);
    ProfScope ps(ctx, "set_constant");
    DISPATCH_D(ctx, (k_set_constant<3><<<grid, 256, 0, ctx->stream>>>(f, ctx->g, C, add)),
               (k_set_constant<6><<<grid, 256, 0, ctx->stream>>>(f, ctx->g, C, add)),
               (k_set_constant<9><<<grid, 256, 0, ctx->stream>>>(f, ctx->g, C, add)));
    FGB_CHECK_LAUNCH(ctx, "k_set_constant");
    return FGB_OK;
}

int fgb_k_copy(fgb_ctx* ctx, const double* src, double* dst, int ncomp) {
    if (src == dst) return FGB_OK;
    FGB_CUDA(ctx, cudaMemcpyAsync(dst, src, sizeof(double) * ctx->g.plane * ncomp, cudaMemcpyDeviceToDevice, ctx->stream));
    return FGB_OK;
}

int fgb_k_xpay(fgb_ctx* ctx, double* r, const double* x, double a, const double* y) {
    const size_t n2 = ctx->g.plane * ctx->dim / 2;
    ProfScope ps(ctx, "xpay");
    k_axpy<1><<<grid_for(ctx, n2, 256), 256, 0, ctx->stream>>>(r, x, a, y, nullptr, n2, nullptr);
    FGB_CHECK_LAUNCH(ctx, "k_axpy<1>");
    return FGB_OK;
}

int fgb_k_xpay_dev(fgb_ctx* ctx, double* r, const double* x, int scal_index, const double* y) {
    const size_t n2 = ctx->g.plane * ctx->dim / 2;
    ProfScope ps(ctx, "xpay");
    k_axpy<1><<<grid_for(ctx, n2, 256), 256, 0, ctx->stream>>>(r, x, 0.0, y, nullptr, n2, ctx->d_scalars + scal_index);
    FGB_CHECK_LAUNCH(ctx, "k_axpy<1>");
    return FGB_OK;
}

int fgb_k_xpaymz(fgb_ctx* ctx, double* r, const double* x, double a, const double* y, const double* z) {
    const size_t n2 = ctx->g.plane * ctx->dim / 2;
    ProfScope ps(ctx, "xpaymz");
    k_axpy<2><<<grid_for(ctx, n2, 256), 256, 0, ctx->stream>>>(r, x, a, y, z, n2, nullptr);
    FGB_CHECK_LAUNCH(ctx, "k_axpy<2>");
    return FGB_OK;
}

int fgb_k_adjust_residual(fgb_ctx* ctx, double* r, const double* E, const double* z) {
    Const9 C;
    for (int i = 0; i < 9; i++) C.v[i] = i < ctx->dim ? E[i] : 0;
    const unsigned grid = grid_for(ctx, ctx->g.plane / 2, 256);
    ProfScope ps(ctx, "adjust_residual");
    DISPATCH_D(ctx, (k_adjust_residual<3><<<grid, 256, 0, ctx->stream>>>(r, ctx->g, C, z)),
               (k_adjust_residual<6><<<grid, 256, 0, ctx->stream>>>(r, ctx->g, C, z)),
               (k_adjust_residual<9><<<grid, 256, 0, ctx->stream>>>(r, ctx->g, C, z)));
    FGB_CHECK_LAUNCH(ctx, "k_adjust_residual");
    return FGB_OK;
}

int fgb_k_calc_stress_const(fgb_ctx* ctx, const double* src, double* dst, double mu0, double lambda0) {
    const unsigned grid = grid_for(ctx, ctx->g.plane, 256);
    ProfScope ps(ctx, "calc_stress_const");
    DISPATCH_D(ctx, (k_stress_const<3><<<grid, 256, 0, ctx->stream>>>(src, dst, ctx->g, 2 * mu0, lambda0)),
               (k_stress_const<6><<<grid, 256, 0, ctx->stream>>>(src, dst, ctx->g, 2 * mu0, lambda0)),
               (k_stress_const<9><<<grid, 256, 0, ctx->stream>>>(src, dst, ctx->g, 2 * mu0, lambda0)));
    FGB_CHECK_LAUNCH(ctx, "k_stress_const");
    return FGB_OK;
}

static size_t npairs_of(const fgb_ctx* ctx) { return (size_t)ctx->g.lnx * ctx->g.ny * ((ctx->g.nz + 1) / 2); }

int fgb_k_inner(fgb_ctx* ctx, const double* a, const double* b, const double* c, double* out) {
    const void* kp = (c ? (ctx->dim == 3 ? (const void*)k_inner<3, 1> : ctx->dim == 6 ? (const void*)k_inner<6, 1> : (const void*)k_inner<9, 1>)
                        : (ctx->dim == 3 ? (const void*)k_inner<3, 0> : ctx->dim == 6 ? (const void*)k_inner<6, 0> : (const void*)k_inner<9, 0>));
    const unsigned grid = fgb_wave_grid(ctx, kp, 256, npairs_of(ctx), ctx->red_blocks);
    {
        ProfScope ps(ctx, "inner_product");
        if (c) {
            DISPATCH_D(ctx, (k_inner<3, 1><<<grid, 256, 0, ctx->stream>>>(a, b, c, ctx->g, ctx->d_partials)),
                       (k_inner<6, 1><<<grid, 256, 0, ctx->stream>>>(a, b, c, ctx->g, ctx->d_partials)),
                       (k_inner<9, 1><<<grid, 256, 0, ctx->stream>>>(a, b, c, ctx->g, ctx->d_partials)));
        } else {
            DISPATCH_D(ctx, (k_inner<3, 0><<<grid, 256, 0, ctx->stream>>>(a, b, c, ctx->g, ctx->d_partials)),
                       (k_inner<6, 0><<<grid, 256, 0, ctx->stream>>>(a, b, c, ctx->g, ctx->d_partials)),
                       (k_inner<9, 0><<<grid, 256, 0, ctx->stream>>>(a, b, c, ctx->g, ctx->d_partials)));
        }
        FGB_CHECK_LAUNCH(ctx, "k_inner");
    }
    int rc = fgb_reduce_finish(ctx, grid, 1, 0, out);
    if (rc) return rc;
    out[0] /= (double)ctx->g.nx * ctx->g.ny * ctx->g.nz;
    return FGB_OK;
}

int fgb_k_component_dot(fgb_ctx* ctx, const double* a, const double* b, double* out, int mean_only) {
    const void* kp = (mean_only ? (ctx->dim == 3 ? (const void*)k_component_dot<3, 1> : ctx->dim == 6 ? (const void*)k_component_dot<6, 1> : (const void*)k_component_dot<9, 1>)
                           : (ctx->dim == 3 ? (const void*)k_component_dot<3, 0> : ctx->dim == 6 ? (const void*)k_component_dot<6, 0> : (const void*)k_component_dot<9, 0>));
    const unsigned grid = fgb_wave_grid(ctx, kp, 256, npairs_of(ctx), ctx->red_blocks);
    {
        ProfScope ps(ctx, "component_dot");
        if (mean_only) {
            DISPATCH_D(ctx, (k_component_dot<3, 1><<<grid, 256, 0, ctx->stream>>>(a, b, ctx->g, ctx->d_partials)),
                       (k_component_dot<6, 1><<<grid, 256, 0, ctx->stream>>>(a, b, ctx->g, ctx->d_partials)),
                       (k_component_dot<9, 1><<<grid, 256, 0, ctx->stream>>>(a, b, ctx->g, ctx->d_partials)));
        } else {
            DISPATCH_D(ctx, (k_component_dot<3, 0><<<grid, 256, 0, ctx->stream>>>(a, b, ctx->g, ctx->d_partials)),
                       (k_component_dot<6, 0><<<grid, 256, 0, ctx->stream>>>(a, b, ctx->g, ctx->d_partials)),
                       (k_component_dot<9, 0><<<grid, 256, 0, ctx->stream>>>(a, b, ctx->g, ctx->d_partials)));
        }
        FGB_CHECK_LAUNCH(ctx, "k_component_dot");
    }
    int rc = fgb_reduce_finish(ctx, grid, ctx->dim, 0, out);
    if (rc) return rc;
    for (int i = 0; i < ctx->dim; i++) out[i] /= (double)ctx->g.nx * ctx->g.ny * ctx->g.nz;
    return FGB_OK;
}

int fgb_k_cg_update(fgb_ctx* ctx, double* x, double* r, const double* p, const double* w, double a, double* delta) {
    const void* kp = ctx->dim == 3 ? (const void*)k_cg_update<3> : ctx->dim == 6 ? (const void*)k_cg_update<6> : (const void*)k_cg_update<9>;
    const unsigned grid = fgb_wave_grid(ctx, kp, 256, npairs_of(ctx), ctx->red_blocks);
    const double* scal = ctx->cg_dev ? ctx->d_scalars : nullptr;
    {
        ProfScope ps(ctx, "cg_update");
        DISPATCH_D(ctx, (k_cg_update<3><<<grid, 256, 0, ctx->stream>>>(x, r, p, w, a, ctx->g, ctx->d_partials, scal)),
                   (k_cg_update<6><<<grid, 256, 0, ctx->stream>>>(x, r, p, w, a, ctx->g, ctx->d_partials, scal)),
                   (k_cg_update<9><<<grid, 256, 0, ctx->stream>>>(x, r, p, w, a, ctx->g, ctx->d_partials, scal)));
        FGB_CHECK_LAUNCH(ctx, "k_cg_update");
    }
    int rc = fgb_reduce_finish(ctx, grid, 1, 0, delta);
    if (rc) return rc;
    delta[0] /= (double)ctx->g.nx * ctx->g.ny * ctx->g.nz;
    return FGB_OK;
}

// extrapolateLoadstepPolynomial fg:21493-21512: p = Vinv * f, F = <tpowers, p>, in the reference's summation order
struct ExtrapArgs {
    int n;
    const double* f[8];
    double Vinv[64], tp[8];
};
__global__ void __launch_bounds__(256) k_extrapolate_poly(double* __restrict__ dst, ExtrapArgs A, size_t ntot) {
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < ntot; v += (size_t)gridDim.x * blockDim.x) {
        double f[8];
        for (int i = 0; i < A.n; i++) f[i] = A.f[i][v];
        double acc = 0;
        for (int i = 0; i < A.n; i++) {
            double pi = 0;
            for (int j = 0; j < A.n; j++) pi += A.Vinv[i * A.n + j] * f[j];
            acc += A.tp[i] * pi;
        }
        dst[v] = acc;
    }
}

int fgb_k_extrapolate_poly(fgb_ctx* ctx, int n, const double* const* fields, const double* Vinv, const double* tpowers, double* dst) {
    ExtrapArgs A;
    A.n = n;
    for (int i = 0; i < n; i++) { A.f[i] = fields[i]; A.tp[i] = tpowers[i]; }
    for (int i = 0; i < n * n; i++) A.Vinv[i] = Vinv[i];
    const size_t ntot = ctx->g.plane * ctx->dim;
    k_extrapolate_poly<<<grid_for(ctx, ntot, 256), 256, 0, ctx->stream>>>(dst, A, ntot);
    FGB_CHECK_LAUNCH(ctx, "k_extrapolate_poly");
    return FGB_OK;
}

// Device-resident CG scalars.  The arithmetic is the host loop's (runCGElasticity fg:23211-23245), operation by operation, so that
// both forms of the loop produce the same bits: sum / nxyz, + tiny, quotient.
__global__ void k_cg_scalars(int mode, int nranks, const double* __restrict__ sums, double* __restrict__ scal, double nxyz,
                             double* __restrict__ ring, const int* __restrict__ flag) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double s = sums[0];
    for (int r = 1; r < nranks; r++) s = s + sums[r];          // rank order (fgb_allreduce_host)
    s /= nxyz;
    const double tiny = 2.2250738585072014e-308;              // boost::numeric::bounds<double>::smallest()
    if (mode == 0) {
        scal[3] = s;
        const double a = s + tiny;
        scal[2] = scal[0] / a;
    } else {
        ring[0] = scal[0];
        ring[1] = scal[3];
        ring[2] = scal[2];
        ring[3] = s;
        ring[4] = (double)*flag;                               // material-law domain errors / peer time-outs raised so far
        const double delta = s + tiny;
        scal[4] = delta;
        scal[1] = delta / scal[0];
        scal[0] = delta;
    }
}

int fgb_k_cg_scalars(fgb_ctx* ctx, int mode, int ring_slot) {
    const double nxyz = (double)ctx->g.nx * ctx->g.ny * ctx->g.nz;
    const double* sums = ctx->nranks > 1 ? ctx->d_gather : ctx->d_result;
    double* ring = ctx->d_scalars + 8;
    k_cg_scalars<<<1, 32, 0, ctx->stream>>>(mode, ctx->nranks, sums, ctx->d_scalars, nxyz, ring, ctx->d_flag);
    FGB_CHECK_LAUNCH(ctx, "k_cg_scalars");
    if (mode == 1) {
        FGB_CUDA(ctx, cudaMemcpyAsync(ctx->h_ring + 8 * ring_slot, ring, sizeof(double) * 8, cudaMemcpyDeviceToHost, ctx->stream));
        FGB_CUDA(ctx, cudaEventRecord(ctx->ring_ev[ring_slot], ctx->stream));
    }
    return FGB_OK;
}
