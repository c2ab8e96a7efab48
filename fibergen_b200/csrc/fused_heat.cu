// Fused sweeps of the CG iteration for heat conduction / porous flow on the staggered grid (BASELINE config 3: d = 3 tensor
// components, u = 1 transformed component, laminate mixing at interface voxels).
//
// For the linear laws of these modes the mixed law of a voxel is a fixed linear map: Voigt and Reuss by construction, the laminate
// rule because its Newton iteration takes exactly one full step from a = 0 (fg:13196, fg:13367-13370), whose Hessian does not depend
// on F.  With scalar-isotropic phases (ScalarLinearIsotropicMaterialLaw fg:11161) the map is diagonal, q_c = K_c g_c.  K is
// evaluated ONCE per phase / normal / law change by running the generic mixed law (Mixed<3>::PK1, the code the unfused path runs
// every iteration) on the three unit vectors -- so it carries the reference's behaviour to the letter, including the Inf/NaN of the
// component-wise jump formula when a normal component vanishes (fg:13236-13239, fg:9375) -- and the iteration reads 3 doubles per
// voxel instead of phase fractions + normals + a Newton solve.
//
//   k_heat_march  : p_new = r + beta*p_old (fg:23245) ; tau = (K - 2 mu0) p_new (calcStress fg:18134) ; f = div_h tau (fg:18914)
//                   x-marching tile as k_dsd_march (fused.cu): tau is never written
//   k_heat_cg_u   : eta = E + grad_h T (epsOperatorStaggeredHeat fg:18697) is not stored: <p, p - eta> (fg:20871), then
//                   x += alpha p ; r -= alpha (p - eta) ; <r, r> (fg:23221-23240) re-evaluate it from the temperature field
#include "material.cuh"
#include "reduce.cuh"
#include <cstdlib>

// K_c = [P_mix(e_c)]_c and the off-diagonal entries' magnitude (must vanish for the fused path to apply)
__global__ void __launch_bounds__(256) k_heat_tangent(double* __restrict__ K, GridDev g, MaterialDev M, double* __restrict__ partials, int* flag) {
    const unsigned nvox = (unsigned)g.lnx * (unsigned)g.ny * (unsigned)g.nz;
    double off = 0;
    for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += gridDim.x * blockDim.x) {
        const unsigned row_ = v / (unsigned)g.nz;
        const size_t o = (size_t)row_ * g.nzp + (v - row_ * (unsigned)g.nz);
        for (int c = 0; c < 3; c++) {
            double F[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, P[9];
            F[c] = 1.0;
            Mixed<3>::PK1(M, o, F, 1.0, P, flag);
            K[(size_t)c * g.plane + o] = P[c];
            for (int d = 0; d < 3; d++)
                if (d != c && P[d] != 0) off = fmax(off, fabs(P[d]));          // NaN entries compare false: they stay on the diagonal path
        }
    }
    double vals[1] = {off};
    block_reduce_store<1, 2>(vals, partials);
}

template <int UPDATE>
__device__ __forceinline__ double hval(const double* __restrict__ r, const double* __restrict__ p, size_t idx, double cgbeta) {
    if (UPDATE) return __ldg(r + idx) + cgbeta * __ldg(p + idx);
    return __ldg(p + idx);
}

// A thread owns one k column of BJ consecutive y rows and walks a segment of x planes: tau_0 of the previous plane is carried in a
// register, tau_1 of row j-1 is the thread's own previous row (one halo row per plane), tau_2 at k-1 comes through shared memory.
template <int UPDATE, int BJ, int HALO>
__global__ void __launch_bounds__(256, 3) k_heat_march(const double* __restrict__ r, const double* __restrict__ p_old, double* __restrict__ p_new,
                                                    double* __restrict__ u, const double* __restrict__ K, GridDev g, double cgbeta,
                                                    double beta, int SEG, const double* __restrict__ halo, size_t hslot,
                                                    const double* __restrict__ scal) {
    if (scal) cgbeta = scal[1];
    extern __shared__ double zx[];          // [2][BJ][blockDim.x]
    const int NT = blockDim.x;
    const int tid = threadIdx.x;
    const int j0 = blockIdx.x * BJ;
    const int i0 = blockIdx.y * SEG;
    const int i1 = min(i0 + SEG, g.lnx);
    const int k = blockIdx.z * NT + tid;
    const bool active = k < g.nz;
    const int kc = active ? k : 0;
    const size_t P = g.plane;
    const int jm0 = (j0 == 0) ? g.ny - 1 : j0 - 1;
    const int km = (kc == 0) ? g.nz - 1 : kc - 1;
    const bool whole = (gridDim.z == 1);                        // the CTA holds the whole z row: the wrap is a slot index
    const bool edge_lo = !whole && (tid == 0 || kc == 0);
    const int nb_lo = whole ? ((kc == 0) ? g.nz - 1 : tid - 1) : max(tid - 1, 0);
#define ROW(i, j) (((size_t)(i) * g.ny + (j)) * g.nzp)
    double t0_prev[BJ];
    {
        const int im = (i0 == 0) ? g.lnx - 1 : i0 - 1;
#pragma unroll
        for (int jr = 0; jr < BJ; jr++) {
            if (HALO && i0 == 0) {
                // halo slots: [r_0][p_0][K_0] of the left neighbour's last plane (comm.cu: fgb_comm_halo_heat)
                const size_t oh = (size_t)(j0 + jr) * g.nzp + kc;
                const double e0 = hval<UPDATE>(halo, halo + hslot, oh, cgbeta);
                t0_prev[jr] = (__ldg(halo + 2 * hslot + oh) + beta) * e0;
            } else {
                const size_t o = ROW(im, j0 + jr) + kc;
                t0_prev[jr] = (__ldg(K + o) + beta) * hval<UPDATE>(r, p_old, o, cgbeta);
            }
        }
    }
    for (int i = i0; i < i1; i++) {
        double t1_m;
        {
            const size_t o = ROW(i, jm0) + kc;
            t1_m = (__ldg(K + P + o) + beta) * hval<UPDATE>(r, p_old, P + o, cgbeta);
        }
        double t0[BJ], t1[BJ], t2[BJ];
#pragma unroll
        for (int jr = 0; jr < BJ; jr++) {
            const size_t o = ROW(i, j0 + jr) + kc;
            const double e0 = hval<UPDATE>(r, p_old, o, cgbeta), e1 = hval<UPDATE>(r, p_old, P + o, cgbeta), e2 = hval<UPDATE>(r, p_old, 2 * P + o, cgbeta);
            if (UPDATE && active) { p_new[o] = e0; p_new[P + o] = e1; p_new[2 * P + o] = e2; }
            t0[jr] = (__ldg(K + o) + beta) * e0;
            t1[jr] = (__ldg(K + P + o) + beta) * e1;
            t2[jr] = (__ldg(K + 2 * P + o) + beta) * e2;
        }
        double* zb = zx + (size_t)((i - i0) & 1) * (BJ * NT);
#pragma unroll
        for (int jr = 0; jr < BJ; jr++) zb[jr * NT + tid] = t2[jr];
        __syncthreads();
#pragma unroll
        for (int jr = 0; jr < BJ; jr++) {
            double t2_km = zb[jr * NT + nb_lo];
            if (edge_lo) {
                const size_t o = ROW(i, j0 + jr) + km;
                t2_km = (__ldg(K + 2 * P + o) + beta) * hval<UPDATE>(r, p_old, 2 * P + o, cgbeta);
            }
            const double t1_jm = (jr == 0) ? t1_m : t1[jr > 0 ? jr - 1 : 0];
            // divOperatorStaggeredHeat fg:18924-18962
            const double f = (t0[jr] - t0_prev[jr]) * g.hx + (t1[jr] - t1_jm) * g.hy + (t2[jr] - t2_km) * g.hz;
            if (active) u[((size_t)i * g.ny + (j0 + jr)) * (2 * (size_t)g.unzcs) + k] = f;
            t0_prev[jr] = t0[jr];
        }
    }
#undef ROW
}

struct Const3h {
    double v[3];
};

// two voxels (k, k+1) per thread, 16-byte accesses; DOT_ONLY: sum p.(p - eta); else x += a p ; r -= a (p - eta) ; sum r.r
template <int DOT_ONLY>
__global__ void __launch_bounds__(256, 4) k_heat_cg_u(const double* __restrict__ u, const double* __restrict__ p, double* __restrict__ x,
                                                   double* __restrict__ r, double a, GridDev g, Const3h E, double* __restrict__ partials,
                                                   const double* __restrict__ halo_hi, const double* __restrict__ scal) {
    if (!DOT_ONLY && scal) a = scal[2];
    const unsigned nzh = (unsigned)(g.nz + 1) / 2;
    const unsigned npairs = (unsigned)g.lnx * (unsigned)g.ny * nzh;
    const size_t us = 2 * (size_t)g.unzcs;
    double acc = 0;
    for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < npairs; v += gridDim.x * blockDim.x) {
        const unsigned row_ = v / nzh;
        const int k = 2 * (int)(v - row_ * nzh);
        const int i = (int)(row_ / (unsigned)g.ny);
        const int j = (int)(row_ - (unsigned)i * (unsigned)g.ny);
        const bool second = k + 1 < g.nz;
        const int ip = (i + 1 == g.lnx) ? 0 : i + 1;
        const int jp = (j + 1 == g.ny) ? 0 : j + 1;
        const int kp2 = (k + 2 >= g.nz) ? k + 2 - g.nz : k + 2;
        const size_t o = (size_t)row_ * us + k;
        const bool hi_h = halo_hi != nullptr && i + 1 == g.lnx;
#define LD2(ptr) (*reinterpret_cast<const double2*>(ptr))
        const double2 t = LD2(u + o);
        const double2 t_ip = hi_h ? LD2(halo_hi + (size_t)j * us + k) : LD2(u + ((size_t)ip * g.ny + j) * us + k);
        const double2 t_jp = LD2(u + ((size_t)i * g.ny + jp) * us + k);
        const double t_kp2 = u[(size_t)row_ * us + kp2];
        const double t_k1 = second ? t.y : u[(size_t)row_ * us];
#undef LD2
        // epsOperatorStaggeredHeat fg:18717-18752
        const double e0[3] = {E.v[0] + (t_ip.x - t.x) * g.hx, E.v[1] + (t_jp.x - t.x) * g.hy, E.v[2] + (t_k1 - t.x) * g.hz};
        const double e1[3] = {E.v[0] + (t_ip.y - t.y) * g.hx, E.v[1] + (t_jp.y - t.y) * g.hy, E.v[2] + (t_kp2 - t.y) * g.hz};
        const size_t eo = (size_t)row_ * g.nzp + k;
        double s0 = 0, s1 = 0;
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const size_t oo = (size_t)d * g.plane + eo;
            const double2 pv = *reinterpret_cast<const double2*>(p + oo);
            if (DOT_ONLY) {
                s0 += pv.x * (pv.x - e0[d]);
                s1 += pv.y * (pv.y - e1[d]);
                continue;
            }
            double2 xv = *reinterpret_cast<double2*>(x + oo);
            double2 rv = *reinterpret_cast<double2*>(r + oo);
            xv.x = xv.x + a * pv.x;
            xv.y = xv.y + a * pv.y;
            rv.x = rv.x + (-a) * (pv.x - e0[d]);
            rv.y = rv.y + (-a) * (pv.y - e1[d]);
            if (!second) { xv.y = 0.0; rv.y = 0.0; }
            *reinterpret_cast<double2*>(x + oo) = xv;
            *reinterpret_cast<double2*>(r + oo) = rv;
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
// 1 if the fused heat path applies: heat / porous mode, staggered grid, scalar-isotropic phases, lambda_0 = 0 (always, unless a
// <ref> material sets it), no doubly fine grid; the diagonal form of the mixed law is verified when K is built
int fgb_fused_heat_applicable(const fgb_ctx* ctx) {
    static const bool off = getenv("FGB_NO_FUSED_HEAT") != nullptr;
    if (off || ctx->dim != 3 || ctx->scheme != FGB_GAMMA_STAGGERED || ctx->dfg || ctx->nphases < 1) return 0;
    if (ctx->nranks > 1 && !ctx->nccl_comm) return 0;
    for (int p = 0; p < ctx->nphases; p++)
        if (ctx->laws[p].id != FGB_LAW_SCALAR || !ctx->phi[p]) return 0;
    if (ctx->mix == FGB_MIX_LAMINATE && !ctx->normals) return 0;
    return 1;
}

// builds (or reuses) the per-voxel diagonal conductivity; returns FGB_OK and *diag = 0 if the mixed law is not diagonal
int fgb_heat_tangent(fgb_ctx* ctx, int* diag) {
    if (ctx->heatK_valid) { *diag = ctx->heatK_diag; return FGB_OK; }
    const GridDev& g = ctx->g;
    if (!ctx->heatK) {
        cudaError_t e = cudaMalloc(&ctx->heatK, sizeof(double) * g.plane * 3);
        if (e != cudaSuccess) { ctx->heatK = nullptr; return fgb_fail(ctx, FGB_ENOMEM, "cannot allocate the effective conductivity planes"); }
        FGB_CUDA(ctx, cudaMemsetAsync(ctx->heatK, 0, sizeof(double) * g.plane * 3, ctx->stream));
    }
    const MaterialDev M = fgb_material_dev(ctx);
    const size_t nvox = (size_t)g.lnx * g.ny * g.nz;
    unsigned grid = (unsigned)((nvox + 255) / 256);
    if (grid > (unsigned)ctx->red_blocks) grid = ctx->red_blocks;
    {
        ProfScope ps(ctx, "heat_tangent");
        k_heat_tangent<<<grid, 256, 0, ctx->stream>>>(ctx->heatK, g, M, ctx->d_partials, ctx->d_flag);
        FGB_CHECK_LAUNCH(ctx, "k_heat_tangent");
    }
    double off = 0;
    int rc = fgb_reduce_finish(ctx, grid, 1, 2, &off);
    if (rc) return rc;
    ctx->heatK_diag = (off == 0.0) ? 1 : 0;
    ctx->heatK_valid = true;
    *diag = ctx->heatK_diag;
    return FGB_OK;
}

int fgb_k_heat_march(fgb_ctx* ctx, const double* r, double cgbeta, const double* p_old, double* p_new, double mu0, double alpha) {
    ctx->implicit_w_of = -1;
    const GridDev& g = ctx->g;
    const double beta = -alpha * 2 * mu0;
    int threads = 256;
    while (threads > 32 && threads / 2 >= g.nz) threads /= 2;
    const int kchunks = (g.nz + threads - 1) / threads;
    int BJ = (g.ny % 4 == 0) ? 4 : (g.ny % 2 == 0) ? 2 : 1;
    if (const char* e = getenv("FGB_HEAT_BJ")) { const int b = atoi(e); if ((b == 1 || b == 2 || b == 4) && g.ny % b == 0) BJ = b; }
    // segments: whole waves of 3 CTAs per SM, one warm-up plane each
    int SEG = g.lnx;
    {
        double best = 0;
        for (int nseg = 1; nseg <= (g.lnx + 7) / 8; nseg++) {
            const int cand = (g.lnx + nseg - 1) / nseg;
            const long ctas = (long)(g.ny / BJ) * ((g.lnx + cand - 1) / cand) * kchunks;
            const double waves = (double)ctas / (3.0 * ctx->sm_count);
            const double score = waves / ceil(waves) * cand / (cand + 1.0);
            if (score > best * 1.005) { best = score; SEG = cand; }
        }
    }
    const int segs = (g.lnx + SEG - 1) / SEG;
    const double* halo = (ctx->nranks > 1) ? ctx->halo : nullptr;
    const double* scal = (r && ctx->cg_dev) ? ctx->d_scalars : nullptr;
    const size_t smem = sizeof(double) * 2 * BJ * threads;
    dim3 grid(g.ny / BJ, segs, kchunks);
    ProfScope ps(ctx, r ? "heat_dir_flux_div" : "heat_flux_div");
#define HM(U_, BJ_, H_) k_heat_march<U_, BJ_, H_><<<grid, threads, smem, ctx->stream>>>(r, p_old, p_new, ctx->ubuf, ctx->heatK, g, cgbeta, beta, SEG, halo, ctx->halo_slot, scal)
#define HM_BJ(U_, H_)            \
    do {                         \
        if (BJ == 4) HM(U_, 4, H_); \
        else if (BJ == 2) HM(U_, 2, H_); \
        else HM(U_, 1, H_);      \
    } while (0)
    if (r) { if (halo) HM_BJ(1, 1); else HM_BJ(1, 0); }
    else { if (halo) HM_BJ(0, 1); else HM_BJ(0, 0); }
#undef HM_BJ
#undef HM
    FGB_CHECK_LAUNCH(ctx, "k_heat_march");
    return FGB_OK;
}

int fgb_k_heat_cg_u(fgb_ctx* ctx, bool dot_only, const double* Econst, double* x, double* r, const double* p, double a, double* out) {
    const GridDev& g = ctx->g;
    Const3h E;
    for (int i = 0; i < 3; i++) E.v[i] = Econst[i];
    const size_t npairs = (size_t)g.lnx * g.ny * ((g.nz + 1) / 2);
    const unsigned grid = fgb_wave_grid(ctx, dot_only ? (const void*)k_heat_cg_u<1> : (const void*)k_heat_cg_u<0>, 256, npairs, ctx->red_blocks);
    {
        ProfScope ps(ctx, dot_only ? "eps_dot_implicit" : "cg_update_implicit");
        const double* hi = (ctx->nranks > 1) ? ctx->halo + 3 * ctx->halo_slot : nullptr;
        const double* scal = ctx->cg_dev ? ctx->d_scalars : nullptr;
        if (dot_only) k_heat_cg_u<1><<<grid, 256, 0, ctx->stream>>>(ctx->ubuf, p, x, r, a, g, E, ctx->d_partials, hi, scal);
        else k_heat_cg_u<0><<<grid, 256, 0, ctx->stream>>>(ctx->ubuf, p, x, r, a, g, E, ctx->d_partials, hi, scal);
        FGB_CHECK_LAUNCH(ctx, "k_heat_cg_u");
    }
    int rc = fgb_reduce_finish(ctx, grid, 1, 0, out);
    if (rc) return rc;
    out[0] /= (double)g.nx * g.ny * g.nz;
    return FGB_OK;
}
