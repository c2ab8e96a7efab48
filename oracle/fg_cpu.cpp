// fg_cpu.cpp -- OpenMP C++ restatement of the CG iteration of fibergen's Lippmann-Schwinger solver (the CPU baseline).
//
// TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): timed by bench.py's cpu_baseline / --impl reference legs and checked against
// oracle/fg_oracle.py in tests/.  Nothing under fibergen_b200/ may call it.  "fg:N" = /root/reference/src/fibergen.cpp:N.
//
// It keeps the STRUCTURE of the reference's iteration: every stage is its own OpenMP sweep over memory,
//   calcStress fg:18134 (Voigt mixing fg:12752, LinearIsotropic fg:11375) -> divOperatorStaggered fg:18853 -> fftVector fg:18481
//   (3 x r2c + scaling sweep) -> G0OperatorFourierStaggeredGeneral fg:19834 -> fftInvVector fg:18513 -> epsOperatorStaggered fg:18614
//   -> innerProductL2 fg:20871 -> xpay / xpaymz fg:9819/9993 -> innerProductL2 fg:20955 -> xpay,
// in the loop of runCGElasticity fg:23153-23247.  FFTW is replaced by a threaded power-of-two FFT of its own (iterative radix-2
// on rows of contiguous z entries so that the butterflies vectorise; z pass = half-length complex transform per row).
// Scope: linear elasticity, isotropic phases, Voigt mixing, staggered grid, residual estimator, strain BC -- BASELINE configs 2/5.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <omp.h>

typedef std::complex<double> cplx;

namespace {

struct Grid {
    int nx, ny, nz, nzc, nzp;
    size_t plane;          // nx*ny*nzp
    double hx, hy, hz;     // n/L
    double L[3];
};

bool pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

// ---- FFT ------------------------------------------------------------------------------------------------
struct Twiddles {
    std::vector<cplx> w;   // exp(-2 pi i k / n)
    explicit Twiddles(int n) : w(n) {
        for (int k = 0; k < n; k++) {
            const long double a = -2.0L * 3.14159265358979323846264338327950288L * k / n;
            w[k] = cplx((double)cosl(a), (double)sinl(a));
        }
    }
};

inline unsigned bitrev(unsigned x, int bits) {
    unsigned r = 0;
    for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

// in-place radix-2 transform of n "elements", each a row of `len` contiguous complex numbers `stride` apart; sign -1 forward
void fft_rows(cplx* base, int n, size_t stride, int len, const Twiddles& T, int sign) {
    int bits = 0;
    while ((1 << bits) < n) bits++;
    for (int i = 0; i < n; i++) {
        const int j = (int)bitrev((unsigned)i, bits);
        if (j > i) {
            cplx* a = base + (size_t)i * stride;
            cplx* b = base + (size_t)j * stride;
            for (int k = 0; k < len; k++) std::swap(a[k], b[k]);
        }
    }
    for (int half = 1; half < n; half <<= 1) {
        const int step = n / (2 * half);
        for (int blk = 0; blk < n; blk += 2 * half)
            for (int j = 0; j < half; j++) {
                cplx w = T.w[(size_t)j * step];
                if (sign > 0) w = std::conj(w);
                const double wr = w.real(), wi = w.imag();
                double* a = reinterpret_cast<double*>(base + (size_t)(blk + j) * stride);
                double* b = reinterpret_cast<double*>(base + (size_t)(blk + j + half) * stride);
#pragma omp simd
                for (int k = 0; k < len; k++) {
                    const double br = b[2 * k], bi = b[2 * k + 1];
                    const double tr = br * wr - bi * wi, ti = br * wi + bi * wr;
                    const double ar = a[2 * k], ai = a[2 * k + 1];
                    a[2 * k] = ar + tr; a[2 * k + 1] = ai + ti;
                    b[2 * k] = ar - tr; b[2 * k + 1] = ai - ti;
                }
            }
    }
}

struct FFT3 {
    Grid g;
    Twiddles tx, ty, tz, tzh;     // tzh: half length nz/2
    explicit FFT3(const Grid& g_) : g(g_), tx(g_.nx), ty(g_.ny), tz(g_.nz), tzh(std::max(g_.nz / 2, 1)) {}

    // FFT3<double>::forward (fg:7232) on one padded component, followed by the 1/nxyz sweep of fftVector (fg:18501-18506)
    void forward(double* f) const {
        const int nx = g.nx, ny = g.ny, nz = g.nz, nzc = g.nzc, nh = nz / 2;
        cplx* c = reinterpret_cast<cplx*>(f);
#pragma omp parallel
        {
            std::vector<cplx> buf(std::max(nh, 1));
#pragma omp for schedule(static)
            for (long row = 0; row < (long)nx * ny; row++) {
                cplx* r = c + (size_t)row * nzc;
                if (nz == 1) continue;
                // real row of length nz = complex sequence z[m] = x[2m] + i x[2m+1] of length nz/2
                for (int m = 0; m < nh; m++) buf[m] = r[m];
                fft_rows(buf.data(), nh, 1, 1, tzh, -1);
                const cplx z0 = buf[0];
                for (int k = 1; k < nh; k++) {
                    const cplx a = buf[k], b = std::conj(buf[nh - k]);
                    const cplx e = 0.5 * (a + b), o = cplx(0, -0.5) * (a - b);
                    r[k] = e + tz.w[k] * o;
                }
                r[0] = cplx(z0.real() + z0.imag(), 0.0);
                r[nh] = cplx(z0.real() - z0.imag(), 0.0);
            }
#pragma omp for schedule(static)
            for (int i = 0; i < nx; i++) fft_rows(c + (size_t)i * ny * nzc, ny, nzc, nzc, ty, -1);
#pragma omp for schedule(static)
            for (int j = 0; j < ny; j++) fft_rows(c + (size_t)j * nzc, nx, (size_t)ny * nzc, nzc, tx, -1);
        }
        const double scale = 1.0 / ((double)nx * ny * nz);
        const size_t n = (size_t)nx * ny * nzc * 2;
#pragma omp parallel for schedule(static)
        for (size_t q = 0; q < n; q++) f[q] *= scale;
    }

    // FFT3<double>::backward (fg:7239), unnormalised
    void backward(double* f) const {
        const int nx = g.nx, ny = g.ny, nz = g.nz, nzc = g.nzc, nh = nz / 2;
        cplx* c = reinterpret_cast<cplx*>(f);
#pragma omp parallel
        {
            std::vector<cplx> buf(std::max(nh, 1));
#pragma omp for schedule(static)
            for (int j = 0; j < ny; j++) fft_rows(c + (size_t)j * nzc, nx, (size_t)ny * nzc, nzc, tx, +1);
#pragma omp for schedule(static)
            for (int i = 0; i < nx; i++) fft_rows(c + (size_t)i * ny * nzc, ny, nzc, nzc, ty, +1);
#pragma omp for schedule(static)
            for (long row = 0; row < (long)nx * ny; row++) {
                cplx* r = c + (size_t)row * nzc;
                if (nz == 1) continue;
                // Z[k] = (X[k] + conj X[nh-k]) + i conj(W^k) (X[k] - conj X[nh-k]);  Im X[0], Im X[nh] are ignored by c2r
                const double x0 = r[0].real(), xn = r[nh].real();
                buf[0] = cplx(x0 + xn, x0 - xn);
                for (int k = 1; k < nh; k++) {
                    const cplx a = r[k], b = std::conj(r[nh - k]);
                    buf[k] = (a + b) + cplx(0, 1) * std::conj(tz.w[k]) * (a - b);
                }
                fft_rows(buf.data(), nh, 1, 1, tzh, +1);
                for (int m = 0; m < nh; m++) r[m] = buf[m];
            }
        }
    }
};

// ---- field sweeps ----------------------------------------------------------------------------------------------
struct Solver {
    Grid g;
    int nph;
    std::vector<const double*> phi;      // unpadded nx*ny*nz
    std::vector<double> mu, lam;
    double mu0, lambda0;
    FFT3 fft;
    std::vector<double> kpm[3];
    std::vector<cplx> kp[3];

    Solver(const Grid& g_, int nph_, const double* const* phi_, const double* mu_, const double* lam_)
        : g(g_), nph(nph_), phi(phi_, phi_ + nph_), mu(mu_, mu_ + nph_), lam(lam_, lam_ + nph_), mu0(0), lambda0(0), fft(g_) {
        // frequency tables fg:19838-19876
        const int n[3] = {g.nx, g.ny, g.nz};
        for (int a = 0; a < 3; a++) {
            const double h = g.L[a] / (2 * (double)n[a]);
            const double xi_0 = 2 * M_PI * h / g.L[a];
            const int half = (n[a] % 2 == 0) ? (n[a] / 2 - 1) : (n[a] / 2);
            kpm[a].resize(n[a]);
            kp[a].resize(n[a]);
            for (int i = 0; i < n[a]; i++) {
                const double m = (i <= half) ? (double)i : ((double)i - (double)n[a]);
                const double x = xi_0 * m;
                kpm[a][i] = std::sin(x) / h;
                kp[a][i] = kpm[a][i] * std::exp(cplx(0, x));
            }
        }
    }

    size_t idx(int i, int j, int k) const { return ((size_t)i * g.ny + j) * g.nzp + k; }

    // getRefMaterial fg:12153-12236 for Voigt-mixed isotropic phases: the tangent of a voxel is isotropic with mu = sum phi_p mu_p,
    // lambda = sum phi_p lambda_p, eigenvalues {2 mu (x5), 2 mu + 3 lambda}; mu_0 = 1/2 ref_scale 1/2 (lmin + lmax) fg:22283-22313
    void calcRefMaterial() {
        double lmin = INFINITY, lmax = -INFINITY;
        const size_t nv = (size_t)g.nx * g.ny * g.nz;
#pragma omp parallel for schedule(static) reduction(min : lmin) reduction(max : lmax)
        for (size_t v = 0; v < nv; v++) {
            double m = 0, l = 0;
            for (int p = 0; p < nph; p++) {
                const double f = phi[p][v];
                if (f <= 10 * 2.220446049250313e-16) continue;
                m += f * mu[p];
                l += f * lam[p];
            }
            const double e1 = 2 * m, e2 = 2 * m + 3 * l;
            lmin = std::min(lmin, std::min(e1, e2));
            lmax = std::max(lmax, std::max(e1, e2));
        }
        if (lmin < 0) lmin = 0;
        mu0 = 0.5 * (lmin + lmax) * 0.5;
        lambda0 = 0;
    }

    // calcStress fg:18134 with alpha = 1: sigma = P_mix(eps) - 2 mu0 eps - lambda0 tr(eps) I
    void calcStressDiff(const double* eps, double* sig) const {
        const double beta = -2 * mu0, gamma = -lambda0;
#pragma omp parallel for schedule(static) collapse(2)
        for (int i = 0; i < g.nx; i++)
            for (int j = 0; j < g.ny; j++) {
                size_t o = idx(i, j, 0);
                size_t v = ((size_t)i * g.ny + j) * g.nz;
                for (int k = 0; k < g.nz; k++, o++, v++) {
                    double e[6], s[6] = {0, 0, 0, 0, 0, 0};
                    for (int d = 0; d < 6; d++) e[d] = eps[d * g.plane + o];
                    const double tr = e[0] + e[1] + e[2];
                    for (int p = 0; p < nph; p++) {
                        const double f = phi[p][v];
                        if (f <= 10 * 2.220446049250313e-16) continue;
                        const double two_mu = 2 * f * mu[p], ltr = (f * lam[p]) * tr;
                        s[0] += e[0] * two_mu + ltr; s[1] += e[1] * two_mu + ltr; s[2] += e[2] * two_mu + ltr;
                        s[3] += e[3] * two_mu; s[4] += e[4] * two_mu; s[5] += e[5] * two_mu;
                    }
                    for (int d = 0; d < 6; d++) s[d] += beta * e[d];
                    if (gamma != 0) { s[0] += gamma * tr; s[1] += gamma * tr; s[2] += gamma * tr; }
                    for (int d = 0; d < 6; d++) sig[d * g.plane + o] = s[d];
                }
            }
    }

    // divOperatorStaggered fg:18853-18908 (out of place into 3 planes)
    void div(const double* t, double* f) const {
        const size_t P = g.plane;
#pragma omp parallel for schedule(static) collapse(2)
        for (int i = 0; i < g.nx; i++)
            for (int j = 0; j < g.ny; j++) {
                const int im = (i == 0) ? g.nx - 1 : i - 1, ip = (i + 1 == g.nx) ? 0 : i + 1;
                const int jm = (j == 0) ? g.ny - 1 : j - 1, jp = (j + 1 == g.ny) ? 0 : j + 1;
                for (int k = 0; k < g.nz; k++) {
                    const int km = (k == 0) ? g.nz - 1 : k - 1, kq = (k + 1 == g.nz) ? 0 : k + 1;
                    const size_t o = idx(i, j, k);
                    f[o] = (t[o] - t[idx(im, j, k)]) * g.hx + (t[5 * P + idx(i, jp, k)] - t[5 * P + o]) * g.hy + (t[4 * P + idx(i, j, kq)] - t[4 * P + o]) * g.hz;
                    f[P + o] = (t[5 * P + idx(ip, j, k)] - t[5 * P + o]) * g.hx + (t[P + o] - t[P + idx(i, jm, k)]) * g.hy +
                               (t[3 * P + idx(i, j, kq)] - t[3 * P + o]) * g.hz;
                    f[2 * P + o] = (t[4 * P + idx(ip, j, k)] - t[4 * P + o]) * g.hx + (t[3 * P + idx(i, jp, k)] - t[3 * P + o]) * g.hy +
                                   (t[2 * P + o] - t[2 * P + idx(i, j, km)]) * g.hz;
                }
            }
    }

    // G0OperatorFourierStaggeredGeneral fg:19834-19927 on 3 complex planes (alpha = -1)
    void G0(double* f, double alpha) const {
        const double c10 = -alpha / mu0, c20 = -alpha / (mu0 * (1 + mu0 / (lambda0 + mu0)));
        const size_t Pc = g.plane / 2;
        cplx* f0 = reinterpret_cast<cplx*>(f);
        cplx* f1 = f0 + Pc;
        cplx* f2 = f1 + Pc;
#pragma omp parallel for schedule(static) collapse(2)
        for (int i = 0; i < g.nx; i++)
            for (int j = 0; j < g.ny; j++) {
                size_t o = ((size_t)i * g.ny + j) * g.nzc;
                const double s01 = kpm[0][i] * kpm[0][i] + kpm[1][j] * kpm[1][j];
                for (int k = 0; k < g.nzc; k++, o++) {
                    const double norm = s01 + kpm[2][k] * kpm[2][k];
                    const double c1 = c10 / norm, c2 = c20 / (norm * norm);
                    const cplx a = f0[o], b = f1[o], c = f2[o];
                    const cplx fk = c2 * (a * kp[0][i] + b * kp[1][j] + c * kp[2][k]);
                    f0[o] = c1 * a + fk * (-std::conj(kp[0][i]));
                    f1[o] = c1 * b + fk * (-std::conj(kp[1][j]));
                    f2[o] = c1 * c + fk * (-std::conj(kp[2][k]));
                }
            }
        f0[0] = f1[0] = f2[0] = 0;
    }

    // epsOperatorStaggered fg:18614-18692 with E = 0
    void eps(const double* u, double* e) const {
        const size_t P = g.plane;
#pragma omp parallel for schedule(static) collapse(2)
        for (int i = 0; i < g.nx; i++)
            for (int j = 0; j < g.ny; j++) {
                const int im = (i == 0) ? g.nx - 1 : i - 1, ip = (i + 1 == g.nx) ? 0 : i + 1;
                const int jm = (j == 0) ? g.ny - 1 : j - 1, jp = (j + 1 == g.ny) ? 0 : j + 1;
                for (int k = 0; k < g.nz; k++) {
                    const int km = (k == 0) ? g.nz - 1 : k - 1, kq = (k + 1 == g.nz) ? 0 : k + 1;
                    const size_t o = idx(i, j, k);
                    const double u0 = u[o], u1 = u[P + o], u2 = u[2 * P + o];
                    e[o] = (u[idx(ip, j, k)] - u0) * g.hx;
                    e[P + o] = (u[P + idx(i, jp, k)] - u1) * g.hy;
                    e[2 * P + o] = (u[2 * P + idx(i, j, kq)] - u2) * g.hz;
                    e[3 * P + o] = 0.5 * ((u2 - u[2 * P + idx(i, jm, k)]) * g.hy + (u1 - u[P + idx(i, j, km)]) * g.hz);
                    e[4 * P + o] = 0.5 * ((u2 - u[2 * P + idx(im, j, k)]) * g.hx + (u0 - u[idx(i, j, km)]) * g.hz);
                    e[5 * P + o] = 0.5 * ((u1 - u[P + idx(im, j, k)]) * g.hx + (u0 - u[idx(i, jm, k)]) * g.hy);
                }
            }
    }

    // krylovOperator fg:20583: w = -Gamma0 : (C - C0) : p (E = 0)
    void krylov(const double* p, double* w, double* u) const {
        calcStressDiff(p, w);
        div(w, u);
        for (int c = 0; c < 3; c++) fft.forward(u + c * g.plane);
        G0(u, -1.0);
        for (int c = 0; c < 3; c++) fft.backward(u + c * g.plane);
        eps(u, w);
    }

    // innerProductL2 fg:20871 / fg:20955: sum a:(b - c) with Voigt weights, / nxyz
    double inner(const double* a, const double* b, const double* c) const {
        double s = 0;
#pragma omp parallel for schedule(static) collapse(2) reduction(+ : s)
        for (int i = 0; i < g.nx; i++)
            for (int j = 0; j < g.ny; j++) {
                size_t o = idx(i, j, 0);
                for (int k = 0; k < g.nz; k++, o++) {
                    double t = 0;
                    for (int d = 0; d < 6; d++) {
                        const double bv = c ? b[d * g.plane + o] - c[d * g.plane + o] : b[d * g.plane + o];
                        t += ((d >= 3) ? 2.0 : 1.0) * a[d * g.plane + o] * bv;
                    }
                    s += t;
                }
            }
        return s / ((double)g.nx * g.ny * g.nz);
    }

    // r = x + a*(y - z) over all entries (TensorField::xpay / xpaymz fg:9819, fg:9993)
    void xpaymz(double* r, const double* x, double a, const double* y, const double* z) const {
        const size_t n = 6 * g.plane;
#pragma omp parallel for schedule(static)
        for (size_t q = 0; q < n; q++) r[q] = x[q] + a * (z ? y[q] - z[q] : y[q]);
    }
};

}  // namespace

extern "C" {

// number of OpenMP threads for everything in libfgoracle (torchrun exports OMP_NUM_THREADS=1 to its children)
void fgcpu_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }

// Runs warm + steps CG iterations of runCGElasticity (fg:23153-23247) from eps = E and returns the residual history
// sqrt(gamma_k/gamma_0) (ResidualErrorEstimator fg:14397, one entry per iteration), the wall time of the last `steps` iterations,
// the mean stress <P(eps)> of the last iterate and the reference material.  phi: nph unpadded planes (nx*ny*nz doubles, x slowest).
// Returns 0, or -1 if an axis length is not a power of two (the FFT of this restatement is radix-2).
int fgcpu_cg_iterations(int nx, int ny, int nz, const double* L, int nph, const double* const* phi, const double* mu, const double* lam,
                        const double* E, int warm, int steps, double* residuals, double* seconds, int* threads, double* mean_stress,
                        double* mu0_out) {
    if (!pow2(nx) || !pow2(ny) || !pow2(nz) || nz < 2) return -1;
    Grid g;
    g.nx = nx; g.ny = ny; g.nz = nz;
    g.nzc = nz / 2 + 1;
    g.nzp = 2 * g.nzc;
    g.plane = (size_t)nx * ny * g.nzp;
    g.hx = nx / L[0]; g.hy = ny / L[1]; g.hz = nz / L[2];
    for (int a = 0; a < 3; a++) g.L[a] = L[a];
    Solver S(g, nph, phi, mu, lam);
    *threads = omp_get_max_threads();
    const size_t F = 6 * g.plane;
    std::vector<double> eps(F, 0.0), r(F, 0.0), p(F, 0.0), w(F, 0.0), u(3 * g.plane, 0.0);
    const double tiny = 2.2250738585072014e-308;
    S.calcRefMaterial();
    *mu0_out = S.mu0;
    for (int d = 0; d < 6; d++) std::fill(eps.begin() + d * g.plane, eps.begin() + (d + 1) * g.plane, E[d]);
    S.krylov(eps.data(), r.data(), u.data());
    for (int d = 0; d < 6; d++) {          // adjustResidual fg:10012: r += E - eps
        double* rd = r.data() + d * g.plane;
        const double* ed = eps.data() + d * g.plane;
        for (size_t q = 0; q < g.plane; q++) rd[q] += E[d] - ed[q];
    }
    double gamma = S.inner(r.data(), r.data(), nullptr) + tiny;
    const double gamma0 = gamma;
    p = r;
    auto t0 = std::chrono::steady_clock::now();
    for (int it = 0; it < warm + steps; it++) {
        if (it == warm) t0 = std::chrono::steady_clock::now();
        S.krylov(p.data(), w.data(), u.data());
        double alpha = S.inner(p.data(), p.data(), w.data()) + tiny;
        alpha = gamma / alpha;
        S.xpaymz(eps.data(), eps.data(), alpha, p.data(), nullptr);
        residuals[it] = std::sqrt(gamma / gamma0);
        S.xpaymz(r.data(), r.data(), -alpha, p.data(), w.data());
        const double delta = S.inner(r.data(), r.data(), nullptr) + tiny;
        const double beta = delta / gamma;
        gamma = delta;
        S.xpaymz(p.data(), r.data(), beta, p.data(), nullptr);
    }
    *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    // calcMeanStress fg:17793: <P_mix(eps)> = calcStress with mu0 = lambda0 = 0
    {
        const double m0 = S.mu0, l0 = S.lambda0;
        S.mu0 = 0; S.lambda0 = 0;
        S.calcStressDiff(eps.data(), w.data());
        S.mu0 = m0; S.lambda0 = l0;
        for (int d = 0; d < 6; d++) {
            double s = 0;
            for (int i = 0; i < nx; i++)
                for (int j = 0; j < ny; j++)
                    for (int k = 0; k < nz; k++) s += w[d * g.plane + S.idx(i, j, k)];
            mean_stress[d] = s / ((double)nx * ny * nz);
        }
    }
    return 0;
}

}  // extern "C"
