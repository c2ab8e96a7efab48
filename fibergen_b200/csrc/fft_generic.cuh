// Hand-written batched multi-component 3-D real<->complex FFT for sm_100a, with the Green operator
// fused between the forward and the inverse x pass.
//
// Replaces FFT3<double>::forward/backward (FFTW3 r2c/c2r, fg:7204-7245), the 1/nxyz scaling sweeps of
// fftVector/fftTensor (fg:18501-18506, fg:18548-18553) and the Fourier-space operators
// G0OperatorFourierStaggeredGeneral(Heat) (fg:19778-19927) and GammaOperatorFourierCollocated
// (Heat/Hyper) (fg:19302-19745).  Conventions: forward = sign -1 and scaled by 1/nxyz, backward unscaled.
//
// Data movement per 3-D transform of one component (in place, reference layout):
//   z pass : rows of nzp doubles are contiguous; two real rows are packed into one complex pencil
//            (real/imag), transformed in shared memory and unpacked into two half spectra.
//   y pass : pencils with element stride nzc; a CTA owns a tile of T consecutive k for all j so that
//            global accesses are T*16-byte contiguous segments.
//   x pass : as y with stride ny*nzc; the CTA holds all tensor components of its tile, applies the
//            Green operator at every frequency and transforms back before anything returns to HBM.
// The pencil transform is a Stockham autosort FFT in shared memory (radix 4/2 butterflies in registers,
// generic O(p^2) stages for odd prime factors so that every n the reference accepts works).
#pragma once
#include "fgb_internal.h"

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cscale(double s, double2 a) { return make_double2(s * a.x, s * a.y); }

template <int DIR>
__device__ __forceinline__ double2 twiddle(const double2* __restrict__ tw, int idx) {
    double2 w = __ldg(tw + idx);
    if (DIR > 0) w.y = -w.y;
    return w;
}

// Stockham FFT of T interleaved pencils: element e of lane t lives at buf[e*TS + t].
// Returns the buffer holding the result (in or out).  All threads of the CTA must call it.
template <int DIR>
__device__ double2* fft_tile(double2* in, double2* out, const FftPlanDev& P, int T, int TS) {
    const int n = P.n;
    const int nthreads = blockDim.x;
    const int tid = threadIdx.x;
    int Ns = 1;
    for (int s = 0; s < P.nstages; s++) {
        const int r = P.radix[s];
        const int m = n / r;
        const int tws = n / (Ns * r);
        if (r == 4) {
            for (int idx = tid; idx < m * T; idx += nthreads) {
                const int t = idx % T, j = idx / T;
                const int k = j % Ns;
                double2 v0 = in[(j)*TS + t];
                double2 v1 = in[(j + m) * TS + t];
                double2 v2 = in[(j + 2 * m) * TS + t];
                double2 v3 = in[(j + 3 * m) * TS + t];
                if (k) {
                    v1 = cmul(v1, twiddle<DIR>(P.tw, k * tws));
                    v2 = cmul(v2, twiddle<DIR>(P.tw, 2 * k * tws));
                    v3 = cmul(v3, twiddle<DIR>(P.tw, 3 * k * tws));
                }
                const double2 a0 = cadd(v0, v2), a1 = csub(v0, v2), a2 = cadd(v1, v3);
                double2 d = csub(v1, v3);
                // multiply by -i (forward) or +i (inverse)
                const double2 a3 = (DIR < 0) ? make_double2(d.y, -d.x) : make_double2(-d.y, d.x);
                const int j0 = (j / Ns) * Ns * 4 + k;
                out[(j0)*TS + t] = cadd(a0, a2);
                out[(j0 + Ns) * TS + t] = cadd(a1, a3);
                out[(j0 + 2 * Ns) * TS + t] = csub(a0, a2);
                out[(j0 + 3 * Ns) * TS + t] = csub(a1, a3);
            }
        } else if (r == 2) {
            for (int idx = tid; idx < m * T; idx += nthreads) {
                const int t = idx % T, j = idx / T;
                const int k = j % Ns;
                double2 v0 = in[(j)*TS + t];
                double2 v1 = in[(j + m) * TS + t];
                if (k) v1 = cmul(v1, twiddle<DIR>(P.tw, k * tws));
                const int j0 = (j / Ns) * Ns * 2 + k;
                out[(j0)*TS + t] = cadd(v0, v1);
                out[(j0 + Ns) * TS + t] = csub(v0, v1);
            }
        } else {
            // generic radix-r stage: every output is an r-term sum (covers 3,5,7 and large primes)
            for (int idx = tid; idx < n * T; idx += nthreads) {
                const int t = idx % T, o = idx / T;
                const int k = o % Ns;
                const int qo = (o / Ns) % r;
                const int j = (o / (Ns * r)) * Ns + k;
                const int step = (k * tws + qo * m) % n;
                double2 acc = in[j * TS + t];
                int e = 0;
                for (int q = 1; q < r; q++) {
                    e += step;
                    if (e >= n) e -= n;
                    acc = cadd(acc, cmul(in[(j + q * m) * TS + t], twiddle<DIR>(P.tw, e)));
                }
                out[o * TS + t] = acc;
            }
        }
        __syncthreads();
        double2* tmp = in;
        in = out;
        out = tmp;
        Ns *= r;
    }
    return in;
}

// ------------------------------------------------------------------------------------------------
// z pass: in-place r2c / c2r of contiguous rows, two rows per complex pencil
// ------------------------------------------------------------------------------------------------
template <int FWD>
__global__ void __launch_bounds__(256) k_fft_z(double* __restrict__ base, long rows, int nz, int nzc, long nzp,
                                               FftPlanDev P, int T, int TS, double scale) {
    extern __shared__ double2 smem[];
    double2* a = smem;
    double2* b = smem + (size_t)nz * TS;
    const long pair0 = (long)blockIdx.x * T;
    const int tid = threadIdx.x, nth = blockDim.x;

    if (FWD) {
        for (int idx = tid; idx < T * nz; idx += nth) {
            const int t = idx / nz, z = idx % nz;
            const long r0 = 2 * (pair0 + t), r1 = r0 + 1;
            double re = 0, im = 0;
            if (r0 < rows) re = base[r0 * nzp + z];
            if (r1 < rows) im = base[r1 * nzp + z];
            a[z * TS + t] = make_double2(re, im);
        }
        __syncthreads();
        double2* res = fft_tile<-1>(a, b, P, T, TS);
        for (int idx = tid; idx < T * nzc; idx += nth) {
            const int t = idx / nzc, k = idx % nzc;
            const long r0 = 2 * (pair0 + t), r1 = r0 + 1;
            if (r0 >= rows) continue;
            const double2 zk = res[k * TS + t];
            const double2 zn = res[((nz - k) % nz) * TS + t];
            const double h = 0.5 * scale;
            // A = (Z[k] + conj(Z[n-k]))/2 ; B = (Z[k] - conj(Z[n-k]))/(2i)
            const double2 A = make_double2(h * (zk.x + zn.x), h * (zk.y - zn.y));
            const double2 B = make_double2(h * (zk.y + zn.y), -h * (zk.x - zn.x));
            reinterpret_cast<double2*>(base + r0 * nzp)[k] = A;
            if (r1 < rows) reinterpret_cast<double2*>(base + r1 * nzp)[k] = B;
        }
    } else {
        // c2r: imaginary parts of the DC (and Nyquist) bins are ignored, like FFTW's c2r
        for (int idx = tid; idx < T * nzc; idx += nth) {
            const int t = idx / nzc, k = idx % nzc;
            const long r0 = 2 * (pair0 + t), r1 = r0 + 1;
            double2 A = make_double2(0, 0), B = make_double2(0, 0);
            if (r0 < rows) A = reinterpret_cast<const double2*>(base + r0 * nzp)[k];
            if (r1 < rows) B = reinterpret_cast<const double2*>(base + r1 * nzp)[k];
            const bool selfconj = (k == 0) || (2 * k == nz);
            if (selfconj) {
                a[k * TS + t] = make_double2(A.x, B.x);
            } else {
                a[k * TS + t] = make_double2(A.x - B.y, A.y + B.x);
                a[(nz - k) * TS + t] = make_double2(A.x + B.y, B.x - A.y);
            }
        }
        __syncthreads();
        double2* res = fft_tile<+1>(a, b, P, T, TS);
        for (int idx = tid; idx < T * nz; idx += nth) {
            const int t = idx / nz, z = idx % nz;
            const long r0 = 2 * (pair0 + t), r1 = r0 + 1;
            const double2 v = res[z * TS + t];
            if (r0 < rows) base[r0 * nzp + z] = v.x;
            if (r1 < rows) base[r1 * nzp + z] = v.y;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// strided pass (y, and x without Green operator): one component per CTA
// ------------------------------------------------------------------------------------------------
template <int DIR>
__global__ void __launch_bounds__(256) k_fft_strided(const double2* __restrict__ src, double2* __restrict__ dst, FftPlanDev P,
                                                     PencilMap mi, PencilMap mo, int ninner, int T, PeerTable pt) {
    extern __shared__ double2 smem[];
    const int n = P.n;
    double2* a = smem;
    double2* b = smem + (size_t)n * T;
    const int tid = threadIdx.x, nth = blockDim.x;
    const int inner0 = blockIdx.x * T;
    const double2* gi = src + (long)blockIdx.z * mi.cstride + (long)blockIdx.y * mi.ostride + inner0;
    const long coff = (long)blockIdx.z * mo.cstride + (long)blockIdx.y * mo.ostride + inner0;
    const int tmax = min(T, ninner - inner0);
    for (int idx = tid; idx < n * T; idx += nth) {
        const int e = idx / T, t = idx % T;
        a[idx] = (t < tmax) ? gi[mi.at(e) + t] : make_double2(0, 0);
    }
    __syncthreads();
    double2* res = fft_tile<DIR>(a, b, P, T, T);
    for (int idx = tid; idx < n * T; idx += nth) {
        const int e = idx / T, t = idx % T;
        if (t < tmax) {
            double2* b = pt.n ? pt.p[e / mo.seglen] : dst;
            b[coff + mo.at(e) + t] = res[idx];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Green operators at one frequency
// ------------------------------------------------------------------------------------------------
struct GreenDev {
    int kind;
    double c10, c20, beta, alpha;
    double dc[9];
    int freq_hack;
    int nx, ny, nz;
    const double* kpm[3];
    const double2* kp[3];
    const double* xi[3];
    const double* xi2pi[3];      // 2 pi m / L per index (fg:20159, fg:22073, fg:19089)
    const double2* wex[3];       // Willot-R: 1 + exp(i q), q = xi2pi * L/n (fg:19132)
    const double* wtan[3];       // Willot-R: 0.25 tan(q/2)               (fg:19153)
    double wvox[3];              // Willot-R: voxel size L/n              (fg:19115-19117)
    const double* pois[3];       // poisson_solve fg:23454: (n/L)^2 (cos(2 pi i / n) - 1) per index
};

// a / b from a reciprocal estimate r ~ 1/b by one Markstein correction step: the correctly rounded quotient (as the reference's
// division) except in astronomically rare half-way cases, at 3 FMAs instead of a second division sequence
__device__ __forceinline__ double div_by_rcp(double a, double b, double r) {
    const double q = a * r;
    const double rem = fma(-q, b, a);
    return fma(rem, r, q);
}

// G0OperatorFourierStaggeredGeneral, fg:19834-19927
__device__ __forceinline__ void green_staggered(const GreenDev& G, int ii, int jj, int kk, double2* f) {
    const double s0 = __ldg(G.kpm[0] + ii), s1 = __ldg(G.kpm[1] + jj), s2 = __ldg(G.kpm[2] + kk);
    const double norm = s0 * s0 + s1 * s1 + s2 * s2;
    // c1 = c10/norm, c2 = c20/(norm*norm) (fg:19899-19900): one reciprocal serves both quotients
    const double r = __drcp_rn(norm);
    const double n2 = norm * norm;
    const double c1 = div_by_rcp(G.c10, norm, r);
    const double c2 = div_by_rcp(G.c20, n2, r * r);
    const double2 kp0 = __ldg(G.kp[0] + ii), kp1 = __ldg(G.kp[1] + jj), kp2 = __ldg(G.kp[2] + kk);
    const double2 fkp = cadd(cadd(cmul(f[0], kp0), cmul(f[1], kp1)), cmul(f[2], kp2));
    const double2 c2fkp = cscale(c2, fkp);
    f[0] = cadd(cscale(c1, f[0]), cmul(c2fkp, make_double2(-kp0.x, kp0.y)));
    f[1] = cadd(cscale(c1, f[1]), cmul(c2fkp, make_double2(-kp1.x, kp1.y)));
    f[2] = cadd(cscale(c1, f[2]), cmul(c2fkp, make_double2(-kp2.x, kp2.y)));
}

// G0OperatorFourierStaggeredGeneralHeat, fg:19778-19830
__device__ __forceinline__ void green_staggered_heat(const GreenDev& G, int ii, int jj, int kk, double2* f) {
    const double s0 = __ldg(G.kpm[0] + ii), s1 = __ldg(G.kpm[1] + jj), s2 = __ldg(G.kpm[2] + kk);
    const double norm = s0 * s0 + s1 * s1 + s2 * s2;
    f[0] = cscale(G.c10 / norm, f[0]);
}

// the 21 coefficients of APPLY_GAMMA_CALC_G, fg:19435-19456
__device__ __forceinline__ void gamma_el_coeffs(double* g, double c1, double c2, double xi0, double xi1, double xi2,
                                                double S0, double S1, double S2, double s, bool accumulate) {
    const double xi00 = xi0 * xi0, xi11 = xi1 * xi1, xi22 = xi2 * xi2;
    const double xi01 = xi0 * xi1, xi02 = xi0 * xi2, xi12 = xi1 * xi2;
    const double c12 = c1 * 2;
    const double c3 = c12 + c2 * xi00, c4 = c12 + c2 * xi11, c5 = c12 + c2 * xi22;
    double v[21];
    v[0] = (c12 + c3) * xi00;                // 00
    v[1] = c2 * xi00 * xi11;                 // 10
    v[2] = c2 * xi00 * xi22;                 // 20
    v[3] = c2 * xi00 * xi12 * S1 * S2;       // 30
    v[4] = c3 * xi02 * S0 * S2;              // 40
    v[5] = c3 * xi01 * S0 * S1;              // 50
    v[6] = (c12 + c4) * xi11;                // 11
    v[7] = c2 * xi11 * xi22;                 // 21
    v[8] = c4 * xi12 * S1 * S2;              // 31
    v[9] = c2 * xi11 * xi02 * S0 * S2;       // 41
    v[10] = c4 * xi01 * S0 * S1;             // 51
    v[11] = (c12 + c5) * xi22;               // 22
    v[12] = c5 * xi12 * S1 * S2;             // 32
    v[13] = c5 * xi02 * S0 * S2;             // 42
    v[14] = c2 * xi22 * xi01 * S0 * S1;      // 52
    v[15] = c1 * (xi11 + xi22) + c2 * xi11 * xi22;   // 33
    v[16] = (c1 + c2 * xi22) * xi01 * S0 * S1;       // 43
    v[17] = (c1 + c2 * xi11) * xi02 * S0 * S2;       // 53
    v[18] = c1 * (xi00 + xi22) + c2 * xi00 * xi22;   // 44
    v[19] = (c1 + c2 * xi00) * xi12 * S1 * S2;       // 54
    v[20] = c1 * (xi00 + xi11) + c2 * xi00 * xi11;   // 55
#pragma unroll
    for (int i = 0; i < 21; i++) g[i] = accumulate ? g[i] + s * v[i] : v[i];
}

// index of symmetric entry (i>=j) in the 21-vector above (column-major lower triangle)
__device__ __forceinline__ int sym21(int i, int j) {
    if (i < j) { int t = i; i = j; j = t; }
    // column j starts at offset j*6 - j*(j-1)/2
    return j * 6 - (j * (j - 1)) / 2 + (i - j);
}

// GammaOperatorFourierCollocated, fg:19381-19608
__device__ __forceinline__ void green_colloc_el(const GreenDev& G, int ii, int jj, int kk, double2* f) {
    const double xi0 = __ldg(G.xi[0] + ii), xi1 = __ldg(G.xi[1] + jj), xi2 = __ldg(G.xi[2] + kk);
    const double norm = xi0 * xi0 + xi1 * xi1 + xi2 * xi2;
    const double c1 = G.c10 / norm;
    const double c2 = G.c20 / (norm * norm);
    double g[21];
    const bool fi = G.freq_hack && (G.nx % 2 == 0) && ii == G.nx / 2;
    const bool fj = G.freq_hack && (G.ny % 2 == 0) && jj == G.ny / 2;
    const bool fk = G.freq_hack && (G.nz % 2 == 0) && kk == G.nz / 2;
    if (fi || fj || fk) {
#pragma unroll
        for (int i = 0; i < 21; i++) g[i] = 0;
        double s = 1;
        if (fi) s *= 0.5;
        if (fj) s *= 0.5;
        if (fk) s *= 0.5;
        for (int i = 1; i >= (fi ? -1 : 1); i -= 2)
            for (int j = 1; j >= (fj ? -1 : 1); j -= 2)
                for (int k2 = 1; k2 >= (fk ? -1 : 1); k2 -= 2)
                    gamma_el_coeffs(g, c1, c2, xi0, xi1, xi2, (double)i, (double)j, (double)k2, s, true);
    } else {
        gamma_el_coeffs(g, c1, c2, xi0, xi1, xi2, 1.0, 1.0, 1.0, 1.0, false);
    }
    double2 ey[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        double2 a = cadd(cadd(cscale(g[sym21(i, 0)], f[0]), cscale(g[sym21(i, 1)], f[1])), cscale(g[sym21(i, 2)], f[2]));
        double2 b = cadd(cadd(cscale(g[sym21(i, 3)], f[3]), cscale(g[sym21(i, 4)], f[4])), cscale(g[sym21(i, 5)], f[5]));
        ey[i] = cadd(a, cscale(2.0, b));
    }
#pragma unroll
    for (int i = 0; i < 6; i++) f[i] = cadd(ey[i], cscale(G.beta, f[i]));
}

// GammaOperatorFourierCollocatedHeat, fg:19302-19377
__device__ __forceinline__ void green_colloc_heat(const GreenDev& G, int ii, int jj, int kk, double2* f) {
    double xi[3] = {__ldg(G.xi[0] + ii), __ldg(G.xi[1] + jj), __ldg(G.xi[2] + kk)};
    const double norm = xi[0] * xi[0] + xi[1] * xi[1] + xi[2] * xi[2];
    const double c1 = G.c10 / norm;
    double2 ey[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        double2 c = make_double2(0, 0);
#pragma unroll
        for (int j = 0; j < 3; j++) c = cadd(c, cscale(c1 * xi[i] * xi[j], f[j]));
        ey[i] = c;
    }
#pragma unroll
    for (int i = 0; i < 3; i++) f[i] = cadd(ey[i], cscale(G.beta, f[i]));
}

// GammaOperatorFourierCollocatedHyper, fg:19619-19745
__device__ __forceinline__ void green_colloc_hyper(const GreenDev& G, int ii, int jj, int kk, double2* f) {
    const int vi[9] = {0, 1, 2, 1, 0, 0, 2, 2, 1};
    const int vj[9] = {0, 1, 2, 2, 2, 1, 1, 0, 0};
    double xi[3] = {__ldg(G.xi[0] + ii), __ldg(G.xi[1] + jj), __ldg(G.xi[2] + kk)};
    const double norm = xi[0] * xi[0] + xi[1] * xi[1] + xi[2] * xi[2];
    const double c1 = G.c10 / norm;
    const double c2 = G.c20 / (norm * norm);
    double2 ey[9];
#pragma unroll
    for (int i = 0; i < 9; i++) {
        double2 c = make_double2(0, 0);
#pragma unroll
        for (int j = 0; j < 9; j++) {
            const double gij = c1 * ((vi[i] == vi[j]) ? xi[vj[i]] * xi[vj[j]] : 0.0) +
                               c2 * (xi[vi[i]] * xi[vj[i]] * xi[vi[j]] * xi[vj[j]]);
            c = cadd(c, cscale(gij, f[j]));
        }
        ey[i] = c;
    }
#pragma unroll
    for (int i = 0; i < 9; i++) f[i] = cadd(ey[i], cscale(G.beta, f[i]));
}

// G0DivOperatorFourierHyper, fg:20155-20218: u^ = G0^ (i xi . tau^) with xi = 2 pi m / L, 9 -> 3 components (the remaining six are
// left untouched: the caller only transforms components 0..2 back)
__device__ __forceinline__ void green_g0div_hyper(const GreenDev& G, int ii, int jj, int kk, double2* f) {
    const double x0 = __ldg(G.xi2pi[0] + ii), x1 = __ldg(G.xi2pi[1] + jj), x2 = __ldg(G.xi2pi[2] + kk);
    const double norm = x0 * x0 + x1 * x1 + x2 * x2;
    const double c1 = G.c10 / norm;
    const double c2 = G.c20 / (norm * norm);
    // imag * (xi0*a + xi1*b + xi2*c)
    const double2 s1 = cadd(cadd(cscale(x0, f[0]), cscale(x1, f[5])), cscale(x2, f[4]));
    const double2 s2 = cadd(cadd(cscale(x0, f[8]), cscale(x1, f[1])), cscale(x2, f[3]));
    const double2 s3 = cadd(cadd(cscale(x0, f[7]), cscale(x1, f[6])), cscale(x2, f[2]));
    const double2 f1 = make_double2(-s1.y, s1.x), f2 = make_double2(-s2.y, s2.x), f3 = make_double2(-s3.y, s3.x);
    f[0] = cadd(cscale(c1, f1), cscale(c2, cadd(cadd(cscale(x0 * x0, f1), cscale(x0 * x1, f2)), cscale(x0 * x2, f3))));
    f[1] = cadd(cscale(c1, f2), cscale(c2, cadd(cadd(cscale(x1 * x0, f1), cscale(x1 * x1, f2)), cscale(x1 * x2, f3))));
    f[2] = cadd(cscale(c1, f3), cscale(c2, cadd(cadd(cscale(x2 * x0, f1), cscale(x2 * x1, f2)), cscale(x2 * x2, f3))));
}

// GradOperatorFourierHyper, fg:22069-22116: W^ = i xi (x) q^, 3 -> 9 components
__device__ __forceinline__ void green_grad_hyper(const GreenDev& G, int ii, int jj, int kk, double2* f) {
    const double x0 = __ldg(G.xi2pi[0] + ii), x1 = __ldg(G.xi2pi[1] + jj), x2 = __ldg(G.xi2pi[2] + kk);
    const double2 q0 = f[0], q1 = f[1], q2 = f[2];
#define IXQ(xv, q) make_double2(-(xv) * (q).y, (xv) * (q).x)
    f[0] = IXQ(x0, q0); f[1] = IXQ(x1, q1); f[2] = IXQ(x2, q2);
    f[3] = IXQ(x2, q1); f[4] = IXQ(x2, q0); f[5] = IXQ(x1, q0);
    f[6] = IXQ(x1, q2); f[7] = IXQ(x0, q2); f[8] = IXQ(x0, q1);
#undef IXQ
}

// GammaOperatorFourierWillotR, fg:19083-19299 (the branch compiled in the reference: WILLOT_ALLOW_NONZERO_LAMBDA, "defined for
// lambda_0 -> infinity").  G.c10 = mu_0, G.c20 = mu_0/lambda_0.
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ double wr_s(const double2* r, int a, int b, int c) {
    // the s.. terms (fg:19178-19215): a == b -> 4 Im(r_c conj r_a)^2, else -4 Im(r_a conj r_b) Im(r_a conj r_c)
    if (a == b) {
        const double t = cmul(r[c], cconj(r[a])).y;
        return 4.0 * t * t;
    }
    return -4.0 * cmul(r[a], cconj(r[b])).y * cmul(r[a], cconj(r[c])).y;
}
static __device__ __noinline__ void green_willot(const GreenDev& G, int ii, int jj, int kk, double2* f) {
    const int vi[6] = {0, 1, 2, 1, 0, 0};
    const int vj[6] = {0, 1, 2, 2, 2, 1};
    const double2 e0 = __ldg(G.wex[0] + ii), e1 = __ldg(G.wex[1] + jj), e2 = __ldg(G.wex[2] + kk);
    const double2 e012 = cmul(cmul(e0, e1), e2);
    const double t[3] = {__ldg(G.wtan[0] + ii), __ldg(G.wtan[1] + jj), __ldg(G.wtan[2] + kk)};
    double2 r[3], rc[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const double2 k = cmul(make_double2(0.0, t[a]), e012);
        r[a] = make_double2(k.x / G.wvox[a], k.y / G.wvox[a]);
    }
    const double n2 = (r[0].x * r[0].x + r[0].y * r[0].y) + (r[1].x * r[1].x + r[1].y * r[1].y) + (r[2].x * r[2].x + r[2].y * r[2].y);
    const double mag = sqrt(n2) + 2.2250738585072014e-308;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        r[a] = make_double2(r[a].x / mag, r[a].y / mag);
        rc[a] = cconj(r[a]);
    }
    const double2 rr = cadd(cadd(cmul(r[0], r[0]), cmul(r[1], r[1])), cmul(r[2], r[2]));
    const double r2 = rr.x * rr.x + rr.y * rr.y;
    const double mu_0 = G.c10, ml = G.c20;
    const double den = mu_0 * (2 * (1 + ml) - r2);
    double2 g[6][6];
    for (int iv = 0; iv < 6; iv++)
        for (int jv = iv; jv < 6; jv++) {
            const int i = vi[iv], j = vj[iv], k = vi[jv], l = vj[jv];
            const double sjk = wr_s(r, k, j, i), sjl = wr_s(r, l, j, i), sik = wr_s(r, k, i, j), sil = wr_s(r, l, i, j);
            const double djk = (j == k), dik = (i == k), djl = (j == l), dil = (i == l);
            const double2 ril = cmul(r[i], rc[l]), rjl = cmul(r[j], rc[l]), rik = cmul(r[i], rc[k]), rjk = cmul(r[j], rc[k]);
            const double2 a1 = cadd(cadd(cadd(cscale(djk, ril), cscale(dik, rjl)), cscale(djl, rik)), cscale(dil, rjk));
            const double2 a2 = cadd(cadd(cadd(cscale(sjk, ril), cscale(sik, rjl)), cscale(sjl, rik)), cscale(sil, rjk));
            const double re = cmul(r[i], rc[j]).x * cmul(r[k], rc[l]).x;
            const double2 a3 = cmul(cmul(cmul(cscale(ml, r[i]), r[j]), rc[k]), rc[l]);
            double2 num = cscale((1 + 2 * ml) * 0.25, a1);
            num = cadd(num, make_double2(0.25 * a2.x - re, 0.25 * a2.y));
            num = csub(num, a3);
            g[iv][jv] = make_double2(num.x / den, num.y / den);
            g[jv][iv] = cconj(g[iv][jv]);
        }
    double2 ey[6];
    for (int iv = 0; iv < 6; iv++) {
        double2 c = make_double2(0, 0);
        for (int j = 3; j < 6; j++) c = cadd(c, cmul(g[iv][j], f[j]));
        c = cscale(2.0, c);
        for (int j = 0; j < 3; j++) c = cadd(c, cmul(g[iv][j], f[j]));
        ey[iv] = c;
    }
#pragma unroll
    for (int j = 0; j < 6; j++) f[j] = cadd(cscale(G.alpha, ey[j]), cscale(G.beta, f[j]));
}

// operator kinds of the fused x pass (GreenArgs::kind)
template <int KIND>
__device__ __forceinline__ void green_apply(const GreenDev& G, int ii, int jj, int kk, double2* f) {
    if (KIND == 1) green_staggered(G, ii, jj, kk, f);
    if (KIND == 2) green_staggered_heat(G, ii, jj, kk, f);
    if (KIND == 3) green_colloc_el(G, ii, jj, kk, f);
    if (KIND == 4) green_colloc_heat(G, ii, jj, kk, f);
    if (KIND == 5) green_colloc_hyper(G, ii, jj, kk, f);
    if (KIND == 6) green_g0div_hyper(G, ii, jj, kk, f);
    if (KIND == 7) green_grad_hyper(G, ii, jj, kk, f);
    if (KIND == 8) green_willot(G, ii, jj, kk, f);
    if (KIND == 11) {
        // G0DivOperatorFourierHyper followed by GradOperatorFourierHyper without leaving Fourier space (the sequence of the
        // reference's "GammaHyper identity" test, fg:24573-24574)
        green_g0div_hyper(G, ii, jj, kk, f);
        green_grad_hyper(G, ii, jj, kk, f);
    }
    if (KIND == 10) {
        // poisson_solve fg:23454-23493: u^ = f^ / (2 sum_a (n_a/L_a)^2 (cos(2 pi i_a/n_a) - 1)); the reference folds the 1/nxyz of its
        // unscaled forward transform into the same divisor, here the forward pass has already applied it
        const double d = 2 * (__ldg(G.pois[0] + ii) + __ldg(G.pois[1] + jj) + __ldg(G.pois[2] + kk));
        f[0] = make_double2(f[0].x / d, f[0].y / d);
    }
    if (KIND == 9) {
        // viscosity, fftTensor(..., zero_trace) fg:18557-18559: component 0 is not transformed, tau^_0 = -(tau^_1 + tau^_2)
        f[0] = make_double2(-(f[1].x + f[2].x), -(f[1].y + f[2].y));
        green_colloc_el(G, ii, jj, kk, f);
    }
}

// ------------------------------------------------------------------------------------------------
// x pass fused with the Green operator: forward x, operator, inverse x -- one HBM round trip
// layout seen by this kernel: element (ii, jj, kk) of component c at base[c*cstride + (jj-jbase)*ostride + ii*estride + kk]
// (single GPU: estride = ny*nzcs, ostride = nzcs; y-slab layout after the transpose: estride = nzcs, ostride = nx*nzcs)
// ------------------------------------------------------------------------------------------------
template <int NC, int KIND>
__global__ void __launch_bounds__(256) k_fft_x_green(double2* __restrict__ base, FftPlanDev P, GreenDev G, long estride,
                                                     int ninner, long ostride, long cstride, int T, int jbase, PencilMap xo, PeerTable pt) {
    extern __shared__ double2 smem[];
    const int n = P.n;
    const int tid = threadIdx.x, nth = blockDim.x;
    const size_t bufsz = (size_t)n * T;
    double2* ptr[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) ptr[c] = smem + c * bufsz;
    double2* fr = smem + NC * bufsz;
    const int inner0 = blockIdx.x * T;
    double2* g = base + (long)blockIdx.y * ostride + inner0;
    const int tmax = min(T, ninner - inner0);
#pragma unroll
    for (int c = 0; c < NC; c++) {
        for (int idx = tid; idx < n * T; idx += nth) {
            const int e = idx / T, t = idx % T;
            ptr[c][idx] = (t < tmax) ? g[c * cstride + (long)e * estride + t] : make_double2(0, 0);
        }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < NC; c++) {
        double2* res = fft_tile<-1>(ptr[c], fr, P, T, T);
        if (res != ptr[c]) { fr = ptr[c]; ptr[c] = res; }
    }
    // Green operator
    for (int idx = tid; idx < n * T; idx += nth) {
        const int ii = idx / T, t = idx % T;
        const int inner = inner0 + t;
        if (t >= tmax) continue;
        const int jj = jbase + blockIdx.y;
        const int kk = inner;
        double2 f[NC];
#pragma unroll
        for (int c = 0; c < NC; c++) f[c] = ptr[c][idx];
        if (ii == 0 && jj == 0 && kk == 0) {
#pragma unroll
            for (int c = 0; c < NC; c++) f[c] = make_double2(G.dc[c], 0.0);
        } else {
            green_apply<KIND>(G, ii, jj, kk, f);
        }
#pragma unroll
        for (int c = 0; c < NC; c++) ptr[c][idx] = f[c];
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < NC; c++) {
        double2* res = fft_tile<+1>(ptr[c], fr, P, T, T);
        if (res != ptr[c]) { fr = ptr[c]; ptr[c] = res; }
    }
#pragma unroll
    for (int c = 0; c < NC; c++) {
        for (int idx = tid; idx < n * T; idx += nth) {
            const int e = idx / T, t = idx % T;
            if (t < tmax) {
                double2* b = pt.n ? pt.p[e / xo.seglen] : base;
                b[c * xo.cstride + (long)blockIdx.y * xo.ostride + inner0 + xo.at(e) + t] = ptr[c][idx];
            }
        }
    }
}

