// Fused sweeps of the CG iteration for the staggered scheme (the bandwidth plan of DESIGN.md):
//
//   k_dir_stress_div_iso : p_new = r + beta*p_old                      (TensorField::xpay fg:9819, fg:23245)
//                          tau   = (C - C0) : p_new                    (calcStress fg:18134, Voigt mixing fg:12752,
//                                                                       LinearIsotropic fg:11375)
//                          f     = div_h tau -> u buffer               (divOperatorStaggered fg:18853)
//                          one pass: reads r, p_old, phi; writes p_new and the 3 rhs components.  tau never goes to HBM;
//                          the stencil neighbours are re-evaluated from L1/L2-resident data (z neighbours by warp shuffle).
//   k_eps_dot            : eta = E + sym-grad_h u                      (epsOperatorStaggered fg:18614, applyBCProjector fg:20263)
//                          <p, p - eta>                                (innerProductL2 fg:20871) in the same pass.
#include "fgb_internal.h"
#include <cmath>
#include <cstdlib>
#include "reduce.cuh"

struct IsoPhases {
    int n;
    const double* phi[FGB_MAX_PHASES];
    double mu[FGB_MAX_PHASES], lam[FGB_MAX_PHASES];
};

struct Const9f {
    double v[9];
};

#define FGB_VOIGT_THR (10 * 2.220446049250313e-16)

// Voigt-mixed isotropic stress, accumulated phase by phase like VoigtMixedMaterialLaw::PK1 (fg:12752) with
// LinearIsotropicMaterialLaw::PK1 (fg:11375): S += E*two_mu_p + lambda_p*tr, two_mu_p = 2*phi_p*mu_p.  No per-voxel
// arrays (they would live in local memory): every helper walks the phases once and keeps its sums in registers.

// tau_a, tau_b for two shear components at voxel o
__device__ __forceinline__ void tau_shear2(const IsoPhases& M, size_t o, double ea, double eb, double beta, double& ta, double& tb) {
    ta = 0;
    tb = 0;
    for (int p = 0; p < M.n; p++) {
        const double phi = __ldg(M.phi[p] + o);
        if (phi <= FGB_VOIGT_THR) continue;
        const double two_mu = 2 * phi * M.mu[p];
        ta += ea * two_mu;
        tb += eb * two_mu;
    }
    if (beta != 0) { ta += beta * ea; tb += beta * eb; }
}

// tau_0..2 from the three diagonal strains at voxel o
__device__ __forceinline__ void tau_diag(const IsoPhases& M, size_t o, double e0, double e1, double e2, double beta, double gamma,
                                         double& t0, double& t1, double& t2) {
    const double tr = e0 + e1 + e2;
    t0 = t1 = t2 = 0;
    for (int p = 0; p < M.n; p++) {
        const double phi = __ldg(M.phi[p] + o);
        if (phi <= FGB_VOIGT_THR) continue;
        const double two_mu = 2 * phi * M.mu[p];
        const double ltr = (phi * M.lam[p]) * tr;
        t0 += e0 * two_mu + ltr;
        t1 += e1 * two_mu + ltr;
        t2 += e2 * two_mu + ltr;
    }
    if (beta != 0) { t0 += beta * e0; t1 += beta * e1; t2 += beta * e2; }
    if (gamma != 0) { t0 += gamma * tr; t1 += gamma * tr; t2 += gamma * tr; }
}

// all six components at voxel o
__device__ __forceinline__ void tau_all(const IsoPhases& M, size_t o, const double* e, double beta, double gamma, double* t) {
    const double tr = e[0] + e[1] + e[2];
#pragma unroll
    for (int d = 0; d < 6; d++) t[d] = 0;
    for (int p = 0; p < M.n; p++) {
        const double phi = __ldg(M.phi[p] + o);
        if (phi <= FGB_VOIGT_THR) continue;
        const double two_mu = 2 * phi * M.mu[p];
        const double ltr = (phi * M.lam[p]) * tr;
        t[0] += e[0] * two_mu + ltr;
        t[1] += e[1] * two_mu + ltr;
        t[2] += e[2] * two_mu + ltr;
        t[3] += e[3] * two_mu;
        t[4] += e[4] * two_mu;
        t[5] += e[5] * two_mu;
    }
    if (beta != 0) {
#pragma unroll
        for (int d = 0; d < 6; d++) t[d] += beta * e[d];
    }
    if (gamma != 0) { t[0] += gamma * tr; t[1] += gamma * tr; t[2] += gamma * tr; }
}

template <int UPDATE>
__device__ __forceinline__ double pval(const double* __restrict__ r, const double* __restrict__ p, size_t idx, double cgbeta) {
    if (UPDATE) return __ldg(r + idx) + cgbeta * __ldg(p + idx);
    return __ldg(p + idx);
}

// one CTA walks JB consecutive y rows of one x plane; a thread owns the voxels k = threadIdx.x + m*blockDim.x of each row
template <int UPDATE>
__global__ void __launch_bounds__(256) k_dir_stress_div_iso(const double* __restrict__ r, const double* __restrict__ p_old,
                                                            double* __restrict__ p_new, double* __restrict__ u, GridDev g, IsoPhases M,
                                                            double cgbeta, double beta, double gamma, int JB, const double* __restrict__ scal) {
    if (scal) cgbeta = scal[1];          // device-resident CG scalar (fgb_cgdev_*)
    const int jblocks = (g.ny + JB - 1) / JB;
    const int i = blockIdx.x / jblocks;
    const int j0 = (blockIdx.x - i * jblocks) * JB;
    const int im = (i == 0) ? g.lnx - 1 : i - 1, ip = (i + 1 == g.lnx) ? 0 : i + 1;
    const size_t P = g.plane;
    const int lane = threadIdx.x & 31;
    for (int jj = 0; jj < JB; jj++) {
        const int j = j0 + jj;
        if (j >= g.ny) break;
        const int jm = (j == 0) ? g.ny - 1 : j - 1, jp = (j + 1 == g.ny) ? 0 : j + 1;
        const size_t row = ((size_t)i * g.ny + j) * g.nzp;
        const size_t row_im = ((size_t)im * g.ny + j) * g.nzp, row_ip = ((size_t)ip * g.ny + j) * g.nzp;
        const size_t row_jm = ((size_t)i * g.ny + jm) * g.nzp, row_jp = ((size_t)i * g.ny + jp) * g.nzp;
        const size_t urow = ((size_t)i * g.ny + j) * (2 * (size_t)g.unzcs);
        for (int k0 = 0; k0 < g.nz; k0 += blockDim.x) {
            const int k = k0 + threadIdx.x;
            const bool active = k < g.nz;
            double t[6] = {0, 0, 0, 0, 0, 0};
            const size_t o = row + k;
            if (active) {
                double e[6];
#pragma unroll
                for (int d = 0; d < 6; d++) e[d] = pval<UPDATE>(r, p_old, d * P + o, cgbeta);
                if (UPDATE) {
#pragma unroll
                    for (int d = 0; d < 6; d++) p_new[d * P + o] = e[d];
                }
                tau_all(M, o, e, beta, gamma, t);
            }
            const double t0 = t[0], t1 = t[1], t2 = t[2], t3 = t[3], t4 = t[4], t5 = t[5];
            // z neighbours: tau4, tau3 at k+1 and tau2 at k-1 come from the adjacent lanes whenever they hold them
            double t4_kp = __shfl_down_sync(0xffffffffu, t4, 1);
            double t3_kp = __shfl_down_sync(0xffffffffu, t3, 1);
            double t2_km = __shfl_up_sync(0xffffffffu, t2, 1);
            if (!active) continue;
            double dmy0, dmy1;
            if (lane == 31 || k + 1 >= g.nz) {
                const size_t on = row + ((k + 1 == g.nz) ? 0 : k + 1);
                tau_shear2(M, on, pval<UPDATE>(r, p_old, 4 * P + on, cgbeta), pval<UPDATE>(r, p_old, 3 * P + on, cgbeta), beta, t4_kp, t3_kp);
            }
            if (lane == 0 || k == 0) {
                const size_t on = row + ((k == 0) ? g.nz - 1 : k - 1);
                tau_diag(M, on, pval<UPDATE>(r, p_old, on, cgbeta), pval<UPDATE>(r, p_old, P + on, cgbeta),
                         pval<UPDATE>(r, p_old, 2 * P + on, cgbeta), beta, gamma, dmy0, dmy1, t2_km);
            }
            // x and y neighbours: re-evaluate the needed stress components from r, p_old, phi (L1/L2 hits)
            double t0_im, t1_jm, t5_ip, t4_ip, t5_jp, t3_jp;
            {
                size_t on = row_im + k;
                tau_diag(M, on, pval<UPDATE>(r, p_old, on, cgbeta), pval<UPDATE>(r, p_old, P + on, cgbeta),
                         pval<UPDATE>(r, p_old, 2 * P + on, cgbeta), beta, gamma, t0_im, dmy0, dmy1);
                on = row_jm + k;
                tau_diag(M, on, pval<UPDATE>(r, p_old, on, cgbeta), pval<UPDATE>(r, p_old, P + on, cgbeta),
                         pval<UPDATE>(r, p_old, 2 * P + on, cgbeta), beta, gamma, dmy0, t1_jm, dmy1);
                on = row_ip + k;
                tau_shear2(M, on, pval<UPDATE>(r, p_old, 5 * P + on, cgbeta), pval<UPDATE>(r, p_old, 4 * P + on, cgbeta), beta, t5_ip, t4_ip);
                on = row_jp + k;
                tau_shear2(M, on, pval<UPDATE>(r, p_old, 5 * P + on, cgbeta), pval<UPDATE>(r, p_old, 3 * P + on, cgbeta), beta, t5_jp, t3_jp);
            }
            // divOperatorStaggered fg:18863-18901
            const double f0 = (t0 - t0_im) * g.hx + (t5_jp - t5) * g.hy + (t4_kp - t4) * g.hz;
            const double f1 = (t5_ip - t5) * g.hx + (t1 - t1_jm) * g.hy + (t3_kp - t3) * g.hz;
            const double f2 = (t4_ip - t4) * g.hx + (t3_jp - t3) * g.hy + (t2 - t2_km) * g.hz;
            u[urow + k] = f0;
            u[g.uplane + urow + k] = f1;
            u[2 * g.uplane + urow + k] = f2;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// x-marching version of the fused sweep.  A thread owns one k column of BJ consecutive y rows and walks along x:
//   * y neighbours (tau_1 at j-1, tau_5 / tau_3 at j+1) are its own registers (plus one halo row on either side),
//   * the previous x plane's tau_0 and the next x plane's tau_5 / tau_4 (with its phi) are carried in registers,
//   * z neighbours are exchanged through a double-buffered shared-memory row (one barrier per plane),
// so every r, p_old and phi value of the CTA's own rows is loaded exactly once.
// (Tried and rejected on B200: software L2 prefetch of the next plane, 0.67 -> 0.81 ms; selecting halo planes by pointer
// instead of by branch, 0.64 -> 0.67 ms.)
template <int NP>
__device__ __forceinline__ void phi_load(const IsoPhases& M, size_t o, double* phi) {
#pragma unroll
    for (int p = 0; p < NP; p++) phi[p] = __ldg(M.phi[p] + o);
}
template <int NP>
__device__ __forceinline__ double shear_from(const IsoPhases& M, const double* phi, double e, double beta) {
    double t = 0;
#pragma unroll
    for (int p = 0; p < NP; p++) {
        const double two_mu = (phi[p] > FGB_VOIGT_THR) ? 2 * phi[p] * M.mu[p] : 0.0;
        t += e * two_mu;
    }
    if (beta != 0) t += beta * e;
    return t;
}
template <int NP>
__device__ __forceinline__ void diag_from(const IsoPhases& M, const double* phi, double e0, double e1, double e2, double beta, double gamma,
                                          double& t0, double& t1, double& t2) {
    const double tr = e0 + e1 + e2;
    t0 = t1 = t2 = 0;
#pragma unroll
    for (int p = 0; p < NP; p++) {
        const bool on = phi[p] > FGB_VOIGT_THR;
        const double two_mu = on ? 2 * phi[p] * M.mu[p] : 0.0;
        const double ltr = on ? (phi[p] * M.lam[p]) * tr : 0.0;
        t0 += e0 * two_mu + ltr;
        t1 += e1 * two_mu + ltr;
        t2 += e2 * two_mu + ltr;
    }
    if (beta != 0) { t0 += beta * e0; t1 += beta * e1; t2 += beta * e2; }
    if (gamma != 0) { t0 += gamma * tr; t1 += gamma * tr; t2 += gamma * tr; }
}

static __device__ __forceinline__ void cp_async8(double* smem_dst, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(src) : "memory");
}

template <bool B>
struct BoolTag {
    static constexpr bool value = B;
};

template <int UPDATE, int NP, int BJ, int HALO, int ZW, int NT, int MB>
__global__ void __launch_bounds__(NT, MB) k_dsd_march(const double* __restrict__ r, const double* __restrict__ p_old, double* __restrict__ p_new,
                                                   double* __restrict__ u, GridDev g, IsoPhases M, double cgbeta, double beta, double gamma,
                                                   int SEG, const double* __restrict__ halo, const double* __restrict__ scal) {
    if (scal) cgbeta = scal[1];          // device-resident CG scalar (fgb_cgdev_*)
    // halo != null (slab partition): planes i = -1 and i = lnx come from the neighbour ranks, layout
    // [r_lo 3][p_lo 3][r_hi 2][p_hi 2][phi_lo MAXP][phi_hi MAXP], each ny*nzp doubles (comm.cu: fgb_comm_halo_iso)
    const size_t pe = (size_t)g.ny * g.nzp;
    const int j0 = blockIdx.x * BJ;
    // slab partition: the last plane (whose x neighbour is the hi halo) is the work of an extra row of CTAs, the "tail"; the
    // regular segments cover planes [0, lnx-1) and never look at the hi halo.  (Taking the halo branch inside the loop, or
    // peeling the last plane after the loop, left halo state live across the loop and cost the slab variant 15-25 %.)
    const bool tail = HALO && (blockIdx.y + 1 == gridDim.y);
    const int i0 = tail ? g.lnx - 1 : blockIdx.y * SEG;
    const int i1 = tail ? g.lnx : min(i0 + SEG, HALO ? g.lnx - 1 : g.lnx);
    const int k = blockIdx.z * blockDim.x + threadIdx.x;
    const bool active = k < g.nz;
    const int kc = active ? k : 0;                      // inactive lanes shadow k = 0 (no stores)
    const size_t P = g.plane;
    const int jm0 = (j0 == 0) ? g.ny - 1 : j0 - 1;
    const int jpB = (j0 + BJ == g.ny) ? 0 : j0 + BJ;
    const int kp = (kc + 1 == g.nz) ? 0 : kc + 1, km = (kc == 0) ? g.nz - 1 : kc - 1;
    // z neighbours travel through shared memory (double-buffered: one barrier per plane).  ZW: the CTA holds the whole z row, the
    // periodic wrap is a slot index; otherwise the two threads at the CTA's ends recompute their neighbour from memory.
    // NT = 512 (one CTA per SM) keeps a 257..512-voxel z row inside one CTA, so that no thread takes the slow edge path
    extern __shared__ double zx[];          // [2][3 * BJ][NT]
    const int tid = threadIdx.x;
    const bool edge_hi = !ZW && active && ((tid == (int)blockDim.x - 1) || (k + 1 >= g.nz));          // one thread per CTA end
    const bool edge_lo = !ZW && (tid == 0);
    double* ex = zx + 2 * 3 * BJ * NT;          // !ZW: [2][BJ][10] raw values of the outside z neighbours (hi end, lo end)
    const int nb_hi = ZW ? ((kc + 1 == g.nz) ? 0 : tid + 1) : min(tid + 1, (int)blockDim.x - 1);
    const int nb_lo = ZW ? ((kc == 0) ? g.nz - 1 : tid - 1) : max(tid - 1, 0);
#define ROW(i, j) (((size_t)(i) * g.ny + (j)) * g.nzp)

    double t0_prev[BJ], t5n[BJ], t4n[BJ], phin[BJ][NP];
    // warm-up: tau_0 of plane i0-1, and tau_5 / tau_4 / phi of plane i0
    {
        const int im = (i0 == 0) ? g.lnx - 1 : i0 - 1;
        const bool from_halo = HALO && (i0 == 0);
#pragma unroll
        for (int jr = 0; jr < BJ; jr++) {
            double ph[NP], d1, d2;
            size_t o = ROW(im, j0 + jr) + kc;
            if (from_halo) {
                const size_t oh = (size_t)(j0 + jr) * g.nzp + kc;
#pragma unroll
                for (int q = 0; q < NP; q++) ph[q] = __ldg(halo + (10 + q) * pe + oh);
                diag_from<NP>(M, ph, pval<UPDATE>(halo, halo + 3 * pe, oh, cgbeta), pval<UPDATE>(halo + pe, halo + 4 * pe, oh, cgbeta),
                              pval<UPDATE>(halo + 2 * pe, halo + 5 * pe, oh, cgbeta), beta, gamma, t0_prev[jr], d1, d2);
            } else {
                phi_load<NP>(M, o, ph);
                diag_from<NP>(M, ph, pval<UPDATE>(r, p_old, o, cgbeta), pval<UPDATE>(r, p_old, P + o, cgbeta),
                              pval<UPDATE>(r, p_old, 2 * P + o, cgbeta), beta, gamma, t0_prev[jr], d1, d2);
            }
            o = ROW(i0, j0 + jr) + kc;
            phi_load<NP>(M, o, phin[jr]);
            const double e5 = pval<UPDATE>(r, p_old, 5 * P + o, cgbeta), e4 = pval<UPDATE>(r, p_old, 4 * P + o, cgbeta);
            if (UPDATE && active) { p_new[5 * P + o] = e5; p_new[4 * P + o] = e4; }
            t5n[jr] = shear_from<NP>(M, phin[jr], e5, beta);
            t4n[jr] = shear_from<NP>(M, phin[jr], e4, beta);
        }
    }
    // one plane of the march; LAST = the plane i+1 is the neighbour rank's (hi halo)
    auto plane_step = [&](const int i, auto last_tag) __attribute__((always_inline)) {
        constexpr bool LAST = decltype(last_tag)::value;
        const int ip = (i + 1 == g.lnx) ? 0 : i + 1;
        // halo rows of plane i
        double t1_m, t5_p, t3_p;
        {
            double ph[NP], d0, d2;
            size_t o = ROW(i, jm0) + kc;
            phi_load<NP>(M, o, ph);
            diag_from<NP>(M, ph, pval<UPDATE>(r, p_old, o, cgbeta), pval<UPDATE>(r, p_old, P + o, cgbeta),
                          pval<UPDATE>(r, p_old, 2 * P + o, cgbeta), beta, gamma, d0, t1_m, d2);
            o = ROW(i, jpB) + kc;
            phi_load<NP>(M, o, ph);
            t5_p = shear_from<NP>(M, ph, pval<UPDATE>(r, p_old, 5 * P + o, cgbeta), beta);
            t3_p = shear_from<NP>(M, ph, pval<UPDATE>(r, p_old, 3 * P + o, cgbeta), beta);
        }
        // z row split over several CTAs: the two end threads need their outside neighbour from memory.  They start 8-byte
        // asynchronous copies of the raw values into a shared-memory scratch now (no registers, no waiting) and pick them up after
        // the barrier, so the DRAM latency overlaps the plane's own loads instead of following them
        if (!ZW) {
            if (edge_hi) {
#pragma unroll
                for (int jr = 0; jr < BJ; jr++) {
                    const size_t o = ROW(i, j0 + jr) + kp;
                    double* d = ex + jr * 10;
#pragma unroll
                    for (int q = 0; q < NP; q++) cp_async8(d + q, M.phi[q] + o);
                    cp_async8(d + 3, p_old + 4 * P + o); cp_async8(d + 4, p_old + 3 * P + o);
                    if (UPDATE) { cp_async8(d + 5, r + 4 * P + o); cp_async8(d + 6, r + 3 * P + o); }
                }
            }
            if (edge_lo) {
#pragma unroll
                for (int jr = 0; jr < BJ; jr++) {
                    const size_t o = ROW(i, j0 + jr) + km;
                    double* d = ex + (BJ + jr) * 10;
#pragma unroll
                    for (int q = 0; q < NP; q++) cp_async8(d + q, M.phi[q] + o);
                    cp_async8(d + 3, p_old + o); cp_async8(d + 4, p_old + P + o); cp_async8(d + 5, p_old + 2 * P + o);
                    if (UPDATE) { cp_async8(d + 6, r + o); cp_async8(d + 7, r + P + o); cp_async8(d + 8, r + 2 * P + o); }
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        // own rows of plane i: components 0..3 from memory, 4 and 5 carried from the previous step
        double tc[BJ][6];
#pragma unroll
        for (int jr = 0; jr < BJ; jr++) {
            const size_t o = ROW(i, j0 + jr) + kc;
            const double e0 = pval<UPDATE>(r, p_old, o, cgbeta), e1 = pval<UPDATE>(r, p_old, P + o, cgbeta);
            const double e2 = pval<UPDATE>(r, p_old, 2 * P + o, cgbeta), e3 = pval<UPDATE>(r, p_old, 3 * P + o, cgbeta);
            if (UPDATE && active) { p_new[o] = e0; p_new[P + o] = e1; p_new[2 * P + o] = e2; p_new[3 * P + o] = e3; }
            diag_from<NP>(M, phin[jr], e0, e1, e2, beta, gamma, tc[jr][0], tc[jr][1], tc[jr][2]);
            tc[jr][3] = shear_from<NP>(M, phin[jr], e3, beta);
            tc[jr][4] = t4n[jr];
            tc[jr][5] = t5n[jr];
        }
        // next plane: phi, tau_5, tau_4 (and the p_new values of those two components)
#pragma unroll
        for (int jr = 0; jr < BJ; jr++) {
            const size_t o = ROW(ip, j0 + jr) + kc;
            double e5, e4;
            if (LAST) {
                const size_t oh = (size_t)(j0 + jr) * g.nzp + kc;
#pragma unroll
                for (int q = 0; q < NP; q++) phin[jr][q] = __ldg(halo + (10 + FGB_MAX_PHASES + q) * pe + oh);
                e5 = pval<UPDATE>(halo + 6 * pe, halo + 8 * pe, oh, cgbeta);
                e4 = pval<UPDATE>(halo + 7 * pe, halo + 9 * pe, oh, cgbeta);
            } else {
                phi_load<NP>(M, o, phin[jr]);
                e5 = pval<UPDATE>(r, p_old, 5 * P + o, cgbeta);
                e4 = pval<UPDATE>(r, p_old, 4 * P + o, cgbeta);
            }
            if (UPDATE && active && (i + 1 < i1)) { p_new[5 * P + o] = e5; p_new[4 * P + o] = e4; }
            t5n[jr] = shear_from<NP>(M, phin[jr], e5, beta);
            t4n[jr] = shear_from<NP>(M, phin[jr], e4, beta);
        }
        double* zb = zx + (size_t)((i - i0) & 1) * (3 * BJ * NT);
#pragma unroll
        for (int jr = 0; jr < BJ; jr++) {
            zb[(3 * jr) * NT + tid] = tc[jr][4];
            zb[(3 * jr + 1) * NT + tid] = tc[jr][3];
            zb[(3 * jr + 2) * NT + tid] = tc[jr][2];
        }
        __syncthreads();
#pragma unroll
        for (int jr = 0; jr < BJ; jr++) {
            // z neighbours
            double t4_kp = zb[(3 * jr) * NT + nb_hi];
            double t3_kp = zb[(3 * jr + 1) * NT + nb_hi];
            double t2_km = zb[(3 * jr + 2) * NT + nb_lo];
            if (edge_hi) {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                const double* d = ex + jr * 10;
                double ph[NP];
#pragma unroll
                for (int q = 0; q < NP; q++) ph[q] = d[q];
                const double e4 = UPDATE ? d[5] + cgbeta * d[3] : d[3], e3 = UPDATE ? d[6] + cgbeta * d[4] : d[4];
                t4_kp = shear_from<NP>(M, ph, e4, beta);
                t3_kp = shear_from<NP>(M, ph, e3, beta);
            }
            if (edge_lo) {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                const double* d = ex + (BJ + jr) * 10;
                double ph[NP], d0, d1;
#pragma unroll
                for (int q = 0; q < NP; q++) ph[q] = d[q];
                const double e0 = UPDATE ? d[6] + cgbeta * d[3] : d[3], e1 = UPDATE ? d[7] + cgbeta * d[4] : d[4];
                const double e2 = UPDATE ? d[8] + cgbeta * d[5] : d[5];
                diag_from<NP>(M, ph, e0, e1, e2, beta, gamma, d0, d1, t2_km);
            }
            const double t1_jm = (jr == 0) ? t1_m : tc[jr > 0 ? jr - 1 : 0][1];
            const double t5_jp = (jr == BJ - 1) ? t5_p : tc[jr < BJ - 1 ? jr + 1 : jr][5];
            const double t3_jp = (jr == BJ - 1) ? t3_p : tc[jr < BJ - 1 ? jr + 1 : jr][3];
            // divOperatorStaggered fg:18863-18901
            const double f0 = (tc[jr][0] - t0_prev[jr]) * g.hx + (t5_jp - tc[jr][5]) * g.hy + (t4_kp - tc[jr][4]) * g.hz;
            const double f1 = (t5n[jr] - tc[jr][5]) * g.hx + (tc[jr][1] - t1_jm) * g.hy + (t3_kp - tc[jr][3]) * g.hz;
            const double f2 = (t4n[jr] - tc[jr][4]) * g.hx + (t3_jp - tc[jr][3]) * g.hy + (tc[jr][2] - t2_km) * g.hz;
            if (active) {
                const size_t uo = ((size_t)i * g.ny + (j0 + jr)) * (2 * (size_t)g.unzcs) + k;
                u[uo] = f0;
                u[g.uplane + uo] = f1;
                u[2 * g.uplane + uo] = f2;
            }
            t0_prev[jr] = tc[jr][0];
        }
    };
    if (tail) {
        plane_step(i0, BoolTag<true>{});
    } else {
#pragma unroll 1
        for (int i = i0; i < i1; i++) plane_step(i, BoolTag<false>{});
    }
#undef ROW
}

// marching tiles: block size along z, rows per thread, and the x segment length.  Each segment pays one warm-up plane, and the
// grid should fill whole waves of 2 CTAs per SM.
struct MarchTile {
    int threads, kchunks, BJ, SEG, segs, mb;
};
static MarchTile march_tile(const fgb_ctx* ctx) {
    const GridDev& g = ctx->g;
    MarchTile t;
    // 512-thread CTAs (one per SM) when they waste no more lanes than 256-thread ones: a 257..512-voxel z row then stays inside one
    // CTA, and a longer row has half as many CTA ends (whose threads fetch their outside neighbour from memory)
    const int waste512 = (g.nz + 511) / 512 * 512 - g.nz, waste256 = (g.nz + 255) / 256 * 256 - g.nz;
    t.threads = (g.nz > 256 && waste512 <= waste256 && !getenv("FGB_MARCH_NT256")) ? 512 : 256;
    while (t.threads > 32 && t.threads / 2 >= g.nz) t.threads /= 2;
    t.kchunks = (g.nz + t.threads - 1) / t.threads;
    // two rows per thread (measured at 256^3: 0.645 ms with 4 rows, 0.516 ms with 2, 0.572 ms with 1; at 512^3: 5.04 / 4.24 ms)
    t.BJ = (g.ny % 2 == 0) ? 2 : 1;
    if (const char* e = getenv("FGB_MARCH_BJ")) { const int b = atoi(e); if ((b == 1 || b == 2 || b == 4) && g.ny % b == 0) t.BJ = b; }
    // (measured at 256^3, 2 rows: 0.516 ms when the compiler may use 128 registers, 0.630 ms when held to 80 for a third resident CTA)
    t.mb = 2;
    if (const char* e = getenv("FGB_MARCH_MB")) t.mb = atoi(e) == 3 ? 3 : 2;
    const double resident = (t.threads > 256) ? 1.0 : (t.BJ == 4 ? 2.0 : (t.BJ == 2 ? (double)t.mb : 4.0));          // CTAs per SM (launch bounds of k_dsd_march)
    t.SEG = 16;
    if (const char* e = getenv("FGB_MARCH_SEG")) t.SEG = atoi(e);
    else {
        double best = 0;
        for (int nseg = 1; nseg <= (g.lnx + 7) / 8; nseg++) {
            const int cand = (g.lnx + nseg - 1) / nseg;          // balanced segments
            const long ctas = (long)(g.ny / t.BJ) * ((g.lnx + cand - 1) / cand) * t.kchunks;
            const double waves = (double)ctas / (resident * ctx->sm_count);
            const double score = waves / std::ceil(waves) * cand / (cand + 1.0);
            if (score > best * 1.005) { best = score; t.SEG = cand; }
        }
    }
    if (g.lnx < t.SEG) t.SEG = g.lnx;
    if (t.SEG < 1) t.SEG = 1;
    // slab partition: regular segments over planes [0, lnx-1) plus one row of tail CTAs for the last plane (k_dsd_march)
    const bool slab = ctx->nranks > 1 || getenv("FGB_MARCH_FAKE_HALO");
    t.segs = slab ? (g.lnx - 1 + t.SEG - 1) / t.SEG + 1 : (g.lnx + t.SEG - 1) / t.SEG;
    return t;
}

template <int UPDATE, int NP>
static int launch_march(fgb_ctx* ctx, const double* r, const double* p_old, double* p_new, const IsoPhases& M, double cgbeta, double beta,
                        double gamma) {
    const GridDev& g = ctx->g;
    const double* halo = (ctx->nranks > 1) ? ctx->iso_halo : nullptr;
    // timing experiments only (wrong results): run the slab variant of the kernel on one GPU against a zero halo
    static const bool fake_halo = getenv("FGB_MARCH_FAKE_HALO") != nullptr;
    if (!halo && fake_halo) {
        static double* fh = nullptr;
        if (!fh) {
            const size_t nb = sizeof(double) * (10 + 2 * FGB_MAX_PHASES) * (size_t)g.ny * g.nzp;
            FGB_CUDA(ctx, cudaMalloc(&fh, nb));
            FGB_CUDA(ctx, cudaMemset(fh, 0, nb));
        }
        halo = fh;
    }
    const MarchTile mt = march_tile(ctx);
    const double* scal = (UPDATE && ctx->cg_dev) ? ctx->d_scalars : nullptr;
    const int threads = mt.threads, kchunks = mt.kchunks, SEG = mt.SEG;
    const int segs = mt.segs;
#define LAUNCH_MARCH6(BJ_, H_, Z_, NT_, MB_)                                                                         \
    do {                                                                                                             \
        const size_t smem = sizeof(double) * (2 * 3 * BJ_ * NT_ + (Z_ ? 0 : 2 * BJ_ * 10));                         \
        if (smem > 48 * 1024)                                                                                        \
            FGB_CUDA(ctx, cudaFuncSetAttribute(k_dsd_march<UPDATE, NP, BJ_, H_, Z_, NT_, MB_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_dsd_march<UPDATE, NP, BJ_, H_, Z_, NT_, MB_><<<grid, threads, smem, ctx->stream>>>(r, p_old, p_new, ctx->ubuf, g, M, cgbeta, beta, gamma, SEG, halo, scal); \
    } while (0)
    // resident CTAs per SM asked of the compiler: 256-thread tiles 2 (4 rows), 3 or 2 (2 rows, FGB_MARCH_MB), 4 (1 row); 512-thread tiles 1
#define LAUNCH_MARCH5(BJ_, H_, Z_, NT_)                                                \
    do {                                                                               \
        if constexpr (NT_ == 512) LAUNCH_MARCH6(BJ_, H_, Z_, NT_, 1);                  \
        else if constexpr (BJ_ == 4) LAUNCH_MARCH6(BJ_, H_, Z_, NT_, 2);               \
        else if constexpr (BJ_ == 2) {                                                 \
            if (mt.mb == 2) LAUNCH_MARCH6(BJ_, H_, Z_, NT_, 2);                        \
            else LAUNCH_MARCH6(BJ_, H_, Z_, NT_, 3);                                   \
        } else LAUNCH_MARCH6(BJ_, H_, Z_, NT_, 4);                                     \
    } while (0)
#define LAUNCH_MARCH(BJ_)                                     \
    do {                                                      \
        dim3 grid(g.ny / BJ_, segs, kchunks);                 \
        if (threads > 256 && kchunks == 1) {                  \
            if (halo) LAUNCH_MARCH5(BJ_, 1, 1, 512);          \
            else LAUNCH_MARCH5(BJ_, 0, 1, 512);               \
        } else if (threads > 256) {                           \
            if (halo) LAUNCH_MARCH5(BJ_, 1, 0, 512);          \
            else LAUNCH_MARCH5(BJ_, 0, 0, 512);               \
        } else if (halo) {                                    \
            if (kchunks == 1) LAUNCH_MARCH5(BJ_, 1, 1, 256);  \
            else LAUNCH_MARCH5(BJ_, 1, 0, 256);               \
        } else {                                              \
            if (kchunks == 1) LAUNCH_MARCH5(BJ_, 0, 1, 256);  \
            else LAUNCH_MARCH5(BJ_, 0, 0, 256);               \
        }                                                     \
    } while (0)
    if (mt.BJ == 4) LAUNCH_MARCH(4);
    else if (mt.BJ == 2) LAUNCH_MARCH(2);
    else LAUNCH_MARCH(1);
#undef LAUNCH_MARCH
#undef LAUNCH_MARCH5
#undef LAUNCH_MARCH6
    FGB_CHECK_LAUNCH(ctx, "k_dsd_march");
    return FGB_OK;
}

// eta = E + sym-grad_h u (elasticity), and sum_voxels p:(p - eta) with Voigt weights
// MODE 0: store eta, sum p:(p - eta).  MODE 1: the sum only (eta stays implicit in u).  MODE 2: the CG update with the implicit eta:
// x += a p ; r -= a (p - eta) ; sum r:r  (fg:23221, fg:23237, fg:23240) -- eta is re-evaluated with the same expressions (scalar
// variant, kept for comparison: FGB_CGU_SCALAR; the production kernel is k_cg_update_u6 below).
template <int MODE>
__global__ void __launch_bounds__(256) k_eps_dot6(const double* __restrict__ u, double* __restrict__ eta, const double* __restrict__ p,
                                                  GridDev g, Const9f E, double* __restrict__ partials, const double* __restrict__ halo_lo,
                                                  const double* __restrict__ halo_hi, size_t hslot, double* __restrict__ x,
                                                  double* __restrict__ r, double a, const double* __restrict__ scal) {
    if (MODE == 2 && scal) a = scal[2];
    const unsigned nvox = (unsigned)g.lnx * (unsigned)g.ny * (unsigned)g.nz;
    const size_t us = 2 * (size_t)g.unzcs;
    double acc = 0;
    for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += gridDim.x * blockDim.x) {
        const unsigned row_ = v / (unsigned)g.nz;
        const int k = (int)(v - row_ * (unsigned)g.nz);
        const int i = (int)(row_ / (unsigned)g.ny);
        const int j = (int)(row_ - (unsigned)i * (unsigned)g.ny);
        const int im = (i == 0) ? g.lnx - 1 : i - 1, ip = (i + 1 == g.lnx) ? 0 : i + 1;
        const int jm = (j == 0) ? g.ny - 1 : j - 1, jp = (j + 1 == g.ny) ? 0 : j + 1;
        const int km = (k == 0) ? g.nz - 1 : k - 1, kp = (k + 1 == g.nz) ? 0 : k + 1;
        const size_t o = (size_t)row_ * us + k;
        const size_t o_im = ((size_t)im * g.ny + j) * us + k, o_ip = ((size_t)ip * g.ny + j) * us + k;
        const size_t o_jm = ((size_t)i * g.ny + jm) * us + k, o_jp = ((size_t)i * g.ny + jp) * us + k;
        const size_t o_km = (size_t)row_ * us + km, o_kp = (size_t)row_ * us + kp;
        const double* u0p = u;
        const double* u1p = u + g.uplane;
        const double* u2p = u + 2 * g.uplane;
        const double u0 = u0p[o], u1 = u1p[o], u2 = u2p[o];
        // slab partition: planes i-1 / i+1 outside the slab come from the halo slots (comm.cu: fgb_comm_halo_u)
        const size_t oh = (size_t)j * us + k;
        const bool lo_h = halo_lo != nullptr && i == 0, hi_h = halo_hi != nullptr && i + 1 == g.lnx;
        const double u0_ip = hi_h ? halo_hi[oh] : u0p[o_ip];
        const double u0_im = lo_h ? halo_lo[oh] : u0p[o_im];
        const double u1_im = lo_h ? halo_lo[hslot + oh] : u1p[o_im];
        const double u2_im = lo_h ? halo_lo[2 * hslot + oh] : u2p[o_im];
        (void)u0_im;
        double e[6];
        e[0] = E.v[0] + (u0_ip - u0) * g.hx;
        e[1] = E.v[1] + (u1p[o_jp] - u1) * g.hy;
        e[2] = E.v[2] + (u2p[o_kp] - u2) * g.hz;
        e[3] = E.v[3] + 0.5 * ((u2 - u2p[o_jm]) * g.hy + (u1 - u1p[o_km]) * g.hz);
        e[4] = E.v[4] + 0.5 * ((u2 - u2_im) * g.hx + (u0 - u0p[o_km]) * g.hz);
        e[5] = E.v[5] + 0.5 * ((u1 - u1_im) * g.hx + (u0 - u0p[o_jm]) * g.hy);
        const size_t eo = (size_t)row_ * g.nzp + k;
        double s = 0;
#pragma unroll
        for (int d = 0; d < 6; d++) {
            const double pv = __ldg(p + d * g.plane + eo);
            if (MODE == 2) {
                x[d * g.plane + eo] = x[d * g.plane + eo] + a * pv;
                const double rv = r[d * g.plane + eo] + (-a) * (pv - e[d]);
                r[d * g.plane + eo] = rv;
                s += ((d >= 3) ? 2.0 : 1.0) * rv * rv;
            } else {
                if (MODE == 0) eta[d * g.plane + eo] = e[d];
                s += ((d >= 3) ? 2.0 : 1.0) * pv * (pv - e[d]);
            }
        }
        acc += s;
    }
    double vals[1] = {acc};
    block_reduce_store<1, 0>(vals, partials);
}

// ------------------------------------------------------------------------------------------------
// returns 1 if the fused path applies to this context (D = 6 elasticity, staggered, Voigt, all phases isotropic, one rank)
int fgb_fused_iso_applicable(const fgb_ctx* ctx) {
    if (ctx->dim != 6 || ctx->mode != FGB_MODE_ELASTICITY || ctx->scheme != FGB_GAMMA_STAGGERED) return 0;
    if (ctx->mix != FGB_MIX_VOIGT || ctx->nphases < 1 || ctx->dfg) return 0;
    if (ctx->nranks > 1 && (ctx->nphases > 3 || !ctx->nccl_comm || getenv("FGB_NO_MARCH"))) return 0;   // slab runs need the marching kernel
    for (int p = 0; p < ctx->nphases; p++)
        if (ctx->laws[p].id != FGB_LAW_ISO || !ctx->phi[p]) return 0;
    return 1;
}

int fgb_k_dir_stress_div_iso(fgb_ctx* ctx, const double* r, double cgbeta, const double* p_old, double* p_new, double mu0, double lambda0,
                             double alpha) {
    ctx->implicit_w_of = -1;          // the u buffer is overwritten
    IsoPhases M;
    M.n = ctx->nphases;
    for (int p = 0; p < ctx->nphases; p++) {
        M.phi[p] = ctx->phi[p];
        M.mu[p] = alpha * ctx->laws[p].p[0];
        M.lam[p] = alpha * ctx->laws[p].p[1];
    }
    const double beta = -alpha * 2 * mu0, gamma = -alpha * lambda0;
    const GridDev& g = ctx->g;
    const int JB = 4;
    const unsigned grid = (unsigned)(g.lnx * ((g.ny + JB - 1) / JB));
    int threads = 256;
    while (threads > 32 && threads / 2 >= g.nz) threads /= 2;
    ProfScope ps(ctx, r ? "cg_direction_stress_div" : "stress_div");
    if (ctx->nphases <= 3 && !getenv("FGB_NO_MARCH")) {
        switch (ctx->nphases) {
            case 1: return r ? launch_march<1, 1>(ctx, r, p_old, p_new, M, cgbeta, beta, gamma) : launch_march<0, 1>(ctx, nullptr, p_old, nullptr, M, 0.0, beta, gamma);
            case 2: return r ? launch_march<1, 2>(ctx, r, p_old, p_new, M, cgbeta, beta, gamma) : launch_march<0, 2>(ctx, nullptr, p_old, nullptr, M, 0.0, beta, gamma);
            case 3: return r ? launch_march<1, 3>(ctx, r, p_old, p_new, M, cgbeta, beta, gamma) : launch_march<0, 3>(ctx, nullptr, p_old, nullptr, M, 0.0, beta, gamma);
        }
    }
    if (r) k_dir_stress_div_iso<1><<<grid, threads, 0, ctx->stream>>>(r, p_old, p_new, ctx->ubuf, g, M, cgbeta, beta, gamma, JB, ctx->cg_dev ? ctx->d_scalars : nullptr);
    else k_dir_stress_div_iso<0><<<grid, threads, 0, ctx->stream>>>(nullptr, p_old, nullptr, ctx->ubuf, g, M, 0.0, beta, gamma, JB, nullptr);
    FGB_CHECK_LAUNCH(ctx, "k_dir_stress_div_iso");
    return FGB_OK;
}

// The two sweeps of the implicit-operator-result form, two voxels (k, k+1) per thread so that x, r and p move as 16-byte accesses:
//   eta = sym-grad_h u (same expressions as k_eps_dot6);  DOT_ONLY: sum p:(p - eta);  else: x += a p ; r -= a (p - eta) ; sum r:r
template <int DOT_ONLY>
__global__ void __launch_bounds__(256, 4) k_cg_update_u6(const double* __restrict__ u, const double* __restrict__ p, double* __restrict__ x,
                                                      double* __restrict__ r, double a, GridDev g, Const9f E, double* __restrict__ partials,
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
        const int kp2 = (k + 2 >= g.nz) ? k + 2 - g.nz : k + 2;          // neighbour above the second voxel
        const size_t o = (size_t)row_ * us + k;
        const size_t o_im = ((size_t)im * g.ny + j) * us + k, o_ip = ((size_t)ip * g.ny + j) * us + k;
        const size_t o_jm = ((size_t)i * g.ny + jm) * us + k, o_jp = ((size_t)i * g.ny + jp) * us + k;
        const size_t oh = (size_t)j * us + k;
        const bool lo_h = halo_lo != nullptr && i == 0, hi_h = halo_hi != nullptr && i + 1 == g.lnx;
#define LD2(ptr) (*reinterpret_cast<const double2*>(ptr))
        // rows are 16-byte aligned and k is even; the element after the last voxel of a row is padding (never used when !second)
        const double2 u0 = LD2(u0p + o), u1 = LD2(u1p + o), u2 = LD2(u2p + o);
        const double2 u0_ip = hi_h ? LD2(halo_hi + oh) : LD2(u0p + o_ip);
        const double2 u1_im = lo_h ? LD2(halo_lo + hslot + oh) : LD2(u1p + o_im);
        const double2 u2_im = lo_h ? LD2(halo_lo + 2 * hslot + oh) : LD2(u2p + o_im);
        const double2 u1_jp = LD2(u1p + o_jp), u2_jm = LD2(u2p + o_jm), u0_jm = LD2(u0p + o_jm);
        const double u0_km = u0p[(size_t)row_ * us + km], u1_km = u1p[(size_t)row_ * us + km];
        const double u2_kp2 = u2p[(size_t)row_ * us + kp2];
        double e0[6], e1[6];
        // first voxel (k): neighbours k-1 -> *_km, k+1 -> .y of the pair (or the wrap when nz == k+1)
        const double u2_k1 = second ? u2.y : u2p[(size_t)row_ * us];      // u2 at (k+1) mod nz
        e0[0] = E.v[0] + (u0_ip.x - u0.x) * g.hx;
        e0[1] = E.v[1] + (u1_jp.x - u1.x) * g.hy;
        e0[2] = E.v[2] + (u2_k1 - u2.x) * g.hz;
        e0[3] = E.v[3] + 0.5 * ((u2.x - u2_jm.x) * g.hy + (u1.x - u1_km) * g.hz);
        e0[4] = E.v[4] + 0.5 * ((u2.x - u2_im.x) * g.hx + (u0.x - u0_km) * g.hz);
        e0[5] = E.v[5] + 0.5 * ((u1.x - u1_im.x) * g.hx + (u0.x - u0_jm.x) * g.hy);
        // second voxel (k+1): neighbours k -> .x of the pair, k+2 -> u2_kp2
        e1[0] = E.v[0] + (u0_ip.y - u0.y) * g.hx;
        e1[1] = E.v[1] + (u1_jp.y - u1.y) * g.hy;
        e1[2] = E.v[2] + (u2_kp2 - u2.y) * g.hz;
        e1[3] = E.v[3] + 0.5 * ((u2.y - u2_jm.y) * g.hy + (u1.y - u1.x) * g.hz);
        e1[4] = E.v[4] + 0.5 * ((u2.y - u2_im.y) * g.hx + (u0.y - u0.x) * g.hz);
        e1[5] = E.v[5] + 0.5 * ((u1.y - u1_im.y) * g.hx + (u0.y - u0_jm.y) * g.hy);
#undef LD2
        const size_t eo = (size_t)row_ * g.nzp + k;
        double s0 = 0, s1 = 0;
#pragma unroll
        for (int d = 0; d < 6; d++) {
            const size_t oo = (size_t)d * g.plane + eo;
            const double2 pv = *reinterpret_cast<const double2*>(p + oo);
            const double wgt = (d >= 3) ? 2.0 : 1.0;
            if (DOT_ONLY) {
                s0 += wgt * pv.x * (pv.x - e0[d]);
                s1 += wgt * pv.y * (pv.y - e1[d]);
                continue;
            }
            double2 xv = *reinterpret_cast<double2*>(x + oo);
            double2 rv = *reinterpret_cast<double2*>(r + oo);
            xv.x = xv.x + a * pv.x;
            xv.y = xv.y + a * pv.y;
            rv.x = rv.x + (-a) * (pv.x - e0[d]);
            rv.y = rv.y + (-a) * (pv.y - e1[d]);
            if (!second) { xv.y = 0.0; rv.y = 0.0; }          // padding column: keep it finite (the explicit update leaves garbage here too)
            *reinterpret_cast<double2*>(x + oo) = xv;
            *reinterpret_cast<double2*>(r + oo) = rv;
            s0 += wgt * rv.x * rv.x;
            s1 += wgt * rv.y * rv.y;
        }
        acc += s0;
        if (second) acc += s1;
    }
    double vals[1] = {acc};
    block_reduce_store<1, 0>(vals, partials);
}

static int eps_dot_launch(fgb_ctx* ctx, int mode, const double* u, double* eta, const double* Econst, const double* p, double* x, double* r,
                          double a, double* out) {
    const GridDev& g = ctx->g;
    Const9f E;
    for (int i = 0; i < 9; i++) E.v[i] = i < 6 ? Econst[i] : 0.0;
    const size_t nvox = (size_t)g.lnx * g.ny * g.nz;
    const void* kp = mode == 0 ? (const void*)k_eps_dot6<0> : mode == 1 ? (const void*)k_eps_dot6<1> : (const void*)k_eps_dot6<2>;
    const unsigned grid = fgb_wave_grid(ctx, kp, 256, nvox, ctx->red_blocks);
    {
        ProfScope ps(ctx, mode == 2 ? "cg_update_implicit" : (mode == 1 ? "eps_dot_implicit" : "eps_dot"));
        const double* lo = (ctx->nranks > 1) ? ctx->halo : nullptr;
        const double* hi = (ctx->nranks > 1) ? ctx->halo + 3 * ctx->halo_slot : nullptr;
        const double* scal = ctx->cg_dev ? ctx->d_scalars : nullptr;
        if (mode == 0) k_eps_dot6<0><<<grid, 256, 0, ctx->stream>>>(u, eta, p, g, E, ctx->d_partials, lo, hi, ctx->halo_slot, x, r, a, scal);
        else if (mode == 1) k_eps_dot6<1><<<grid, 256, 0, ctx->stream>>>(u, eta, p, g, E, ctx->d_partials, lo, hi, ctx->halo_slot, x, r, a, scal);
        else k_eps_dot6<2><<<grid, 256, 0, ctx->stream>>>(u, eta, p, g, E, ctx->d_partials, lo, hi, ctx->halo_slot, x, r, a, scal);
        FGB_CHECK_LAUNCH(ctx, "k_eps_dot6");
    }
    int rc = fgb_reduce_finish(ctx, grid, 1, 0, out);
    if (rc) return rc;
    out[0] /= (double)g.nx * g.ny * g.nz;
    return FGB_OK;
}

// eta == nullptr: eta is not stored (it stays implicit in u for fgb_k_cg_update_implicit)
static int implicit_sweep(fgb_ctx* ctx, bool dot_only, const double* u, const double* Econst, double* x, double* r, const double* p, double a,
                          double* out);
int fgb_k_eps_dot(fgb_ctx* ctx, const double* u, double* eta, const double* Econst, const double* p, double* pAp) {
    if (!eta && !getenv("FGB_CGU_SCALAR")) return implicit_sweep(ctx, true, u, Econst, nullptr, nullptr, p, 0.0, pAp);
    return eps_dot_launch(ctx, eta ? 0 : 1, u, eta, Econst, p, nullptr, nullptr, 0.0, pAp);
}

static int implicit_sweep(fgb_ctx* ctx, bool dot_only, const double* u, const double* Econst, double* x, double* r, const double* p, double a,
                          double* out) {
    const GridDev& g = ctx->g;
    Const9f E;
    for (int i = 0; i < 9; i++) E.v[i] = i < 6 ? Econst[i] : 0.0;
    const size_t npairs = (size_t)g.lnx * g.ny * ((g.nz + 1) / 2);
    const unsigned grid = fgb_wave_grid(ctx, dot_only ? (const void*)k_cg_update_u6<1> : (const void*)k_cg_update_u6<0>, 256, npairs, ctx->red_blocks);
    {
        ProfScope ps(ctx, dot_only ? "eps_dot_implicit" : "cg_update_implicit");
        const double* lo = (ctx->nranks > 1) ? ctx->halo : nullptr;
        const double* hi = (ctx->nranks > 1) ? ctx->halo + 3 * ctx->halo_slot : nullptr;
        const double* scal = ctx->cg_dev ? ctx->d_scalars : nullptr;
        if (dot_only) k_cg_update_u6<1><<<grid, 256, 0, ctx->stream>>>(u, p, x, r, a, g, E, ctx->d_partials, lo, hi, ctx->halo_slot, scal);
        else k_cg_update_u6<0><<<grid, 256, 0, ctx->stream>>>(u, p, x, r, a, g, E, ctx->d_partials, lo, hi, ctx->halo_slot, scal);
        FGB_CHECK_LAUNCH(ctx, "k_cg_update_u6");
    }
    int rc = fgb_reduce_finish(ctx, grid, 1, 0, out);
    if (rc) return rc;
    out[0] /= (double)g.nx * g.ny * g.nz;
    return FGB_OK;
}

int fgb_k_cg_update_implicit(fgb_ctx* ctx, const double* u, const double* Econst, double* x, double* r, const double* p, double a, double* delta) {
    if (getenv("FGB_CGU_SCALAR")) return eps_dot_launch(ctx, 2, u, nullptr, Econst, p, x, r, a, delta);
    return implicit_sweep(ctx, false, u, Econst, x, r, p, a, delta);
}
