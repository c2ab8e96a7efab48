// Power-of-two fast path of the pencil FFT: N = R1*R2 with register-resident R-point transforms
// (R in {4,8,16,32}), one shared-memory exchange between the two register passes and warp shuffles for
// the real<->complex split of the z pass.  A pencil is owned by TPP = max(R1,R2) consecutive threads.
//
//   pass 1: thread n2 < R2 holds x[R2*n1 + n2], n1 = 0..R1-1  -> R1-point FFT over n1 -> times W_N^(n2*k1)
//   exchange through shared memory
//   pass 2: thread k1 < R1 holds a[n2][k1], n2 = 0..R2-1      -> R2-point FFT over n2 -> X[k1 + R1*k2]
#pragma once
#include <cuda_runtime.h>

__constant__ double2 c_w32[32];      // exp(-2 pi i k / 32), filled by fgb_fft_init

namespace p2 {

__device__ __forceinline__ double2 pmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 padd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 psub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }

// a * W_N^e (forward, DIR<0) or its conjugate; e is a compile-time constant after unrolling
template <int N, int DIR>
__device__ __forceinline__ double2 twc(double2 a, int e) {
    e = e % N;
    if (e == 0) return a;
    if (4 * e == N) return (DIR < 0) ? make_double2(a.y, -a.x) : make_double2(-a.y, a.x);
    if (2 * e == N) return make_double2(-a.x, -a.y);
    if (4 * e == 3 * N) return (DIR < 0) ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
    double2 w = c_w32[e * (32 / N)];
    if (DIR > 0) w.y = -w.y;
    return pmul(a, w);
}

template <int N, int DIR>
struct RegFFT;

template <int DIR>
struct RegFFT<1, DIR> {
    static __device__ __forceinline__ void run(double2*) {}
};

template <int DIR>
struct RegFFT<2, DIR> {
    static __device__ __forceinline__ void run(double2* v) {
        const double2 t = v[0];
        v[0] = padd(t, v[1]);
        v[1] = psub(t, v[1]);
    }
};

template <int DIR>
struct RegFFT<4, DIR> {
    static __device__ __forceinline__ void run(double2* v) {
        const double2 a0 = padd(v[0], v[2]), a1 = psub(v[0], v[2]), a2 = padd(v[1], v[3]);
        const double2 d = psub(v[1], v[3]);
        const double2 a3 = (DIR < 0) ? make_double2(d.y, -d.x) : make_double2(-d.y, d.x);
        v[0] = padd(a0, a2);
        v[1] = padd(a1, a3);
        v[2] = psub(a0, a2);
        v[3] = psub(a1, a3);
    }
};

// Cooley-Tukey N = N1*N2 on registers, natural order in and out
template <int N1, int N2, int DIR>
__device__ __forceinline__ void fft_ct(double2* v) {
    constexpr int N = N1 * N2;
    double2 a[N];
#pragma unroll
    for (int n2 = 0; n2 < N2; n2++) {
        double2 t[N1];
#pragma unroll
        for (int n1 = 0; n1 < N1; n1++) t[n1] = v[N2 * n1 + n2];
        RegFFT<N1, DIR>::run(t);
#pragma unroll
        for (int k1 = 0; k1 < N1; k1++) a[n2 * N1 + k1] = twc<N, DIR>(t[k1], n2 * k1);
    }
#pragma unroll
    for (int k1 = 0; k1 < N1; k1++) {
        double2 t[N2];
#pragma unroll
        for (int n2 = 0; n2 < N2; n2++) t[n2] = a[n2 * N1 + k1];
        RegFFT<N2, DIR>::run(t);
#pragma unroll
        for (int k2 = 0; k2 < N2; k2++) v[k1 + N1 * k2] = t[k2];
    }
}

template <int DIR>
struct RegFFT<8, DIR> {
    static __device__ __forceinline__ void run(double2* v) { fft_ct<4, 2, DIR>(v); }
};
template <int DIR>
struct RegFFT<16, DIR> {
    static __device__ __forceinline__ void run(double2* v) { fft_ct<4, 4, DIR>(v); }
};
template <int DIR>
struct RegFFT<32, DIR> {
    static __device__ __forceinline__ void run(double2* v) { fft_ct<4, 8, DIR>(v); }
};

template <int A, int B>
struct Max {
    static constexpr int v = A > B ? A : B;
};

// pass 1 on registers: R1-point FFT then the inter-pass twiddle W_N^(n2*k1) from the shared-memory table
template <int R1, int R2, int DIR>
__device__ __forceinline__ void pass1(double2* v, int n2, const double2* __restrict__ tw_s) {
    RegFFT<R1, DIR>::run(v);
#pragma unroll
    for (int k1 = 1; k1 < R1; k1++) {
        double2 w = tw_s[n2 * k1];
        if (DIR > 0) w.y = -w.y;
        v[k1] = pmul(v[k1], w);
    }
}

}  // namespace p2
