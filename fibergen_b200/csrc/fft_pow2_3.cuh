// Three-pass power-of-two pencil FFT for the long axes (N = R1*R2*R3 = 512, 1024, 2048): three register-resident
// R-point transforms (R in {8,16}) and two shared-memory exchanges.  Compared with the two-pass path (fft_pow2.cuh) a
// thread holds 8..16 complex values instead of 32, which keeps the register count low enough for several CTAs per SM.
//
// forward (natural in, natural out), M = R2*R3:
//   pass 1  thread m < M        : v[n1] = x[M n1 + m]          -> R1-FFT over n1 -> * W_N^(m k1)      -> E1[k1][m]
//   pass 2  thread (k1, m3)     : v[n2] = E1[k1][R3 n2 + m3]   -> R2-FFT over n2 -> * W_M^(m3 k2)     -> E2[k2][k1][m3]
//   pass 3  thread s = k2 R1+k1 : v[m3] = E2[k2][k1][m3]       -> R3-FFT over m3 -> X[k1 + R1 k2 + R1 R2 k3] = X[s + R1 R2 k3]
// inverse mirrored (so that the fused x pass can go forward, apply the Green operator on the registers and come back):
//   pass 1' thread s = k2 R1+k1 : w[k3]                        -> R3-iFFT over k3 -> * conj W_M^(k2 nu3) -> F2[nu3][k1][k2]
//   pass 2' thread nu3 R1 + k1  : v[k2] = F2[nu3][k1][k2]      -> R2-iFFT over k2 -> * conj W_N^(k1 nu), nu = nu3 + R3 nb -> F1[nu][k1]
//   pass 3' thread nu < M       : v[k1] = F1[nu][k1]           -> R1-iFFT over k1 -> x[nu + M na]   (the pass-1 distribution)
// Rows that are read with one thread per row are padded by one element so that consecutive rows start 64 bytes apart modulo
// 128 (conflict-free for 16-byte accesses).
#pragma once
#include "fft_pow2.cuh"

namespace p3 {

template <int A, int B, int C>
struct Max3 {
    static constexpr int v = (A > B ? (A > C ? A : C) : (B > C ? B : C));
};

template <int R1, int R2, int R3, int T>
struct Plan {
    static constexpr int N = R1 * R2 * R3;
    static constexpr int M = R2 * R3;
    static constexpr int TPP = Max3<M, R1 * R3, R1 * R2>::v;      // threads per pencil lane
    static constexpr int E1 = N * T;                               // [k1][m]
    static constexpr int E2 = R1 * R2 * (R3 + 1) * T;              // [k2][k1][m3 (+1)]
    static constexpr int F2 = R3 * R1 * (R2 + 1) * T;              // [nu3][k1][k2 (+1)]
    static constexpr int F1 = M * (R1 + 1) * T;                    // [nu][k1 (+1)]
    static constexpr int BUF1 = (E1 > F1 ? E1 : F1);               // E1 and F1 share a buffer, E2 and F2 the other
    static constexpr int BUF2 = (E2 > F2 ? E2 : F2);
    static constexpr int SMEM_ELEMS = BUF1 + BUF2;

    // v: R1 inputs x[M n1 + s] of thread s < M (garbage elsewhere); out: R3 outputs X[s + R1 R2 k3] of thread s < R1*R2
    template <int DIR>
    static __device__ __forceinline__ void forward(double2* v, double2* out, int s, int t, double2* B1, double2* B2,
                                                   const double2* __restrict__ tw_s) {
        if (s < M) {
            p2::RegFFT<R1, DIR>::run(v);
#pragma unroll
            for (int k1 = 0; k1 < R1; k1++) {
                double2 a = v[k1];
                if (k1) {
                    double2 w = tw_s[s * k1];
                    if (DIR > 0) w.y = -w.y;
                    a = p2::pmul(a, w);
                }
                B1[(k1 * M + s) * T + t] = a;
            }
        }
        __syncthreads();
        if (s < R1 * R3) {
            const int k1 = s / R3, m3 = s % R3;
            double2 u[R2];
#pragma unroll
            for (int n2 = 0; n2 < R2; n2++) u[n2] = B1[(k1 * M + R3 * n2 + m3) * T + t];
            p2::RegFFT<R2, DIR>::run(u);
#pragma unroll
            for (int k2 = 0; k2 < R2; k2++) {
                double2 a = u[k2];
                if (k2) {
                    double2 w = tw_s[R1 * m3 * k2];
                    if (DIR > 0) w.y = -w.y;
                    a = p2::pmul(a, w);
                }
                B2[((k2 * R1 + k1) * (R3 + 1) + m3) * T + t] = a;
            }
        }
        __syncthreads();
        if (s < R1 * R2) {
#pragma unroll
            for (int m3 = 0; m3 < R3; m3++) out[m3] = B2[(s * (R3 + 1) + m3) * T + t];
            p2::RegFFT<R3, DIR>::run(out);
        }
    }

    // forward through ONE exchange buffer (rows of M+1): pass 2 overwrites exactly the elements it read, so no second
    // buffer is needed; used by the single-component strided pass where shared memory bounds the residency.
    static constexpr int INPLACE_ELEMS = R1 * (M + 1) * T;
    template <int DIR>
    static __device__ __forceinline__ void forward_inplace(double2* v, double2* out, int s, int t, double2* B,
                                                           const double2* __restrict__ tw_s) {
        if (s < M) {
            p2::RegFFT<R1, DIR>::run(v);
#pragma unroll
            for (int k1 = 0; k1 < R1; k1++) {
                double2 a = v[k1];
                if (k1) {
                    double2 w = tw_s[s * k1];
                    if (DIR > 0) w.y = -w.y;
                    a = p2::pmul(a, w);
                }
                B[(k1 * (M + 1) + s) * T + t] = a;
            }
        }
        __syncthreads();
        if (s < R1 * R3) {
            const int k1 = s / R3, m3 = s % R3;
            double2* row = B + ((size_t)k1 * (M + 1) + m3) * T + t;
            double2 u[R2];
#pragma unroll
            for (int n2 = 0; n2 < R2; n2++) u[n2] = row[R3 * n2 * T];
            p2::RegFFT<R2, DIR>::run(u);
#pragma unroll
            for (int k2 = 0; k2 < R2; k2++) {
                double2 a = u[k2];
                if (k2) {
                    double2 w = tw_s[R1 * m3 * k2];
                    if (DIR > 0) w.y = -w.y;
                    a = p2::pmul(a, w);
                }
                row[R3 * k2 * T] = a;
            }
        }
        __syncthreads();
        if (s < R1 * R2) {
            const int k2 = s / R1, k1 = s % R1;
            const double2* row = B + ((size_t)k1 * (M + 1) + R3 * k2) * T + t;
#pragma unroll
            for (int m3 = 0; m3 < R3; m3++) out[m3] = row[m3 * T];
            p2::RegFFT<R3, DIR>::run(out);
        }
    }

    // w: R3 spectrum values X[s + R1 R2 k3] of thread s < R1*R2; out: R1 values x[nu + M na] of thread nu = s < M
    static __device__ __forceinline__ void inverse(double2* w, double2* out, int s, int t, double2* B1, double2* B2,
                                                   const double2* __restrict__ tw_s) {
        if (s < R1 * R2) {
            const int k2 = s / R1, k1 = s % R1;
            p2::RegFFT<R3, +1>::run(w);
#pragma unroll
            for (int nu3 = 0; nu3 < R3; nu3++) {
                double2 a = w[nu3];
                if (nu3) {
                    double2 tws = tw_s[R1 * k2 * nu3];
                    tws.y = -tws.y;
                    a = p2::pmul(a, tws);
                }
                B2[((nu3 * R1 + k1) * (R2 + 1) + k2) * T + t] = a;
            }
        }
        __syncthreads();
        if (s < R3 * R1) {
            const int nu3 = s / R1, k1 = s % R1;
            double2 u[R2];
#pragma unroll
            for (int k2 = 0; k2 < R2; k2++) u[k2] = B2[(s * (R2 + 1) + k2) * T + t];
            p2::RegFFT<R2, +1>::run(u);
#pragma unroll
            for (int nb = 0; nb < R2; nb++) {
                const int nu = nu3 + R3 * nb;
                double2 a = u[nb];
                if (k1) {
                    double2 tws = tw_s[k1 * nu];
                    tws.y = -tws.y;
                    a = p2::pmul(a, tws);
                }
                B1[(nu * (R1 + 1) + k1) * T + t] = a;
            }
        }
        __syncthreads();
        if (s < M) {
#pragma unroll
            for (int k1 = 0; k1 < R1; k1++) out[k1] = B1[(s * (R1 + 1) + k1) * T + t];
            p2::RegFFT<R1, +1>::run(out);
        }
    }
};

}  // namespace p3
