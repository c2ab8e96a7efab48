// Device-side material laws and mixing rules (restated from the maths of fg:10287-13732).
// Stored component order 11,22,33,23,13,12,32,31,21 (fg:9103); as a matrix [[0,5,4],[8,1,3],[7,6,2]].
#pragma once
#include "fgb_internal.h"

#define FGB_VOIGT_THRESHOLD (10 * 2.220446049250313e-16)   // fg:12736

// ---- small tensor helpers (Tensor3x3 fg:9105-9355, SymTensor3x3 fg:9361-9489) -------------------
__device__ __forceinline__ double det9(const double* F) {
    return F[0] * (F[1] * F[2] - F[3] * F[6]) - F[5] * (F[8] * F[2] - F[3] * F[7]) + F[4] * (F[8] * F[6] - F[1] * F[7]);
}

__device__ __forceinline__ void inv9(const double* F, double* R) {
    // M = [[F0,F5,F4],[F8,F1,F3],[F7,F6,F2]]
    const double c00 = F[1] * F[2] - F[3] * F[6];
    const double c01 = F[8] * F[2] - F[3] * F[7];
    const double c02 = F[8] * F[6] - F[1] * F[7];
    const double det = F[0] * c00 - F[5] * c01 + F[4] * c02;
    const double id = 1.0 / det;
    R[0] = c00 * id;                                   // (0,0)
    R[5] = -(F[5] * F[2] - F[4] * F[6]) * id;          // (0,1)
    R[4] = (F[5] * F[3] - F[4] * F[1]) * id;           // (0,2)
    R[8] = -c01 * id;                                  // (1,0)
    R[1] = (F[0] * F[2] - F[4] * F[7]) * id;           // (1,1)
    R[3] = -(F[0] * F[3] - F[4] * F[8]) * id;          // (1,2)
    R[7] = c02 * id;                                   // (2,0)
    R[6] = -(F[0] * F[6] - F[5] * F[7]) * id;          // (2,1)
    R[2] = (F[0] * F[1] - F[5] * F[8]) * id;           // (2,2)
}

// symmetric 3x3 (order 11,22,33,23,13,12) inverse by adjugate
__device__ __forceinline__ void sym_inv6(const double* H, double* R) {
    const double a = H[0], b = H[1], c = H[2], d = H[3], e = H[4], f = H[5];
    const double det = a * (b * c - d * d) - f * (f * c - d * e) + e * (f * d - b * e);
    const double id = 1.0 / det;
    R[0] = (b * c - d * d) * id;
    R[1] = (a * c - e * e) * id;
    R[2] = (a * b - f * f) * id;
    R[3] = (e * f - a * d) * id;
    R[4] = (f * d - e * b) * id;
    R[5] = (e * d - f * c) * id;
}

__device__ __forceinline__ double dot9(const double* a, const double* b) {
    double s = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) s += a[i] * b[i];
    return s;
}

// fix_dim / fix_sym, fg:12115-12138
template <int D>
__device__ __forceinline__ void fix_dim(double* t) {
    if (D == 6) { t[6] = t[3]; t[7] = t[4]; t[8] = t[5]; }
    if (D == 3) { t[3] = t[4] = t[5] = t[6] = t[7] = t[8] = 0; }
}
template <int D>
__device__ __forceinline__ void fix_sym(double* t) {
    if (D == 6) {
        t[6] = t[3] = 0.5 * (t[3] + t[6]);
        t[7] = t[4] = 0.5 * (t[4] + t[7]);
        t[8] = t[5] = 0.5 * (t[5] + t[8]);
    }
    if (D == 3) { t[3] = t[4] = t[5] = t[6] = t[7] = t[8] = 0; }
}

// P = F*S with S symmetric (6) -- the PK1_OP block of fg:11642-11651 / fg:11769-11778
__device__ __forceinline__ void FS(const double* F, const double* S, double* P) {
    P[0] = F[0] * S[0] + F[5] * S[5] + F[4] * S[4];
    P[1] = F[8] * S[5] + F[1] * S[1] + F[3] * S[3];
    P[2] = F[7] * S[4] + F[6] * S[3] + F[2] * S[2];
    P[3] = F[8] * S[4] + F[1] * S[3] + F[3] * S[2];
    P[4] = F[0] * S[4] + F[5] * S[3] + F[4] * S[2];
    P[5] = F[0] * S[5] + F[5] * S[1] + F[4] * S[3];
    P[6] = F[7] * S[5] + F[6] * S[1] + F[2] * S[3];
    P[7] = F[7] * S[0] + F[6] * S[5] + F[2] * S[4];
    P[8] = F[8] * S[0] + F[1] * S[5] + F[3] * S[4];
}

// C = F^T F (symmetric 6)
__device__ __forceinline__ void right_cauchy_green(const double* F, double* C) {
    C[0] = F[0] * F[0] + F[8] * F[8] + F[7] * F[7];
    C[1] = F[5] * F[5] + F[1] * F[1] + F[6] * F[6];
    C[2] = F[4] * F[4] + F[3] * F[3] + F[2] * F[2];
    C[3] = F[5] * F[4] + F[1] * F[3] + F[6] * F[2];
    C[4] = F[0] * F[4] + F[8] * F[3] + F[7] * F[2];
    C[5] = F[0] * F[5] + F[8] * F[1] + F[7] * F[6];
}

// d/dF (F^T F)/2 : W = (W^T F + F^T W)/2 (symmetric 6)
__device__ __forceinline__ void green_strain_deriv(const double* F, const double* W, double* E) {
    E[0] = W[0] * F[0] + W[8] * F[8] + W[7] * F[7];
    E[1] = W[5] * F[5] + W[1] * F[1] + W[6] * F[6];
    E[2] = W[4] * F[4] + W[3] * F[3] + W[2] * F[2];
    E[3] = 0.5 * ((W[5] * F[4] + W[1] * F[3] + W[6] * F[2]) + (F[5] * W[4] + F[1] * W[3] + F[6] * W[2]));
    E[4] = 0.5 * ((W[0] * F[4] + W[8] * F[3] + W[7] * F[2]) + (F[0] * W[4] + F[8] * W[3] + F[7] * W[2]));
    E[5] = 0.5 * ((W[0] * F[5] + W[8] * F[1] + W[7] * F[6]) + (F[0] * W[5] + F[8] * W[1] + F[7] * W[6]));
}

__device__ __forceinline__ void flag_numeric(int* flag) {
    if (flag) atomicExch(flag, 1);
}

__device__ __forceinline__ double checked_log(double x, int* flag) {
    const double y = log(x);
    if (isnan(y)) flag_numeric(flag);      // fg:10293-10301
    return y;
}

// Finv^T W^T and Finv^T W^T Finv^T (fg:11799-11828)
__device__ __forceinline__ void nh_products(const double* Finv, const double* W, double* A, double* B) {
    A[0] = Finv[0] * W[0] + Finv[8] * W[5] + Finv[7] * W[4];
    A[1] = Finv[5] * W[8] + Finv[1] * W[1] + Finv[6] * W[3];
    A[2] = Finv[4] * W[7] + Finv[3] * W[6] + Finv[2] * W[2];
    A[3] = Finv[5] * W[7] + Finv[1] * W[6] + Finv[6] * W[2];
    A[4] = Finv[0] * W[7] + Finv[8] * W[6] + Finv[7] * W[2];
    A[5] = Finv[0] * W[8] + Finv[8] * W[1] + Finv[7] * W[3];
    A[6] = Finv[4] * W[8] + Finv[3] * W[1] + Finv[2] * W[3];
    A[7] = Finv[4] * W[0] + Finv[3] * W[5] + Finv[2] * W[4];
    A[8] = Finv[5] * W[0] + Finv[1] * W[5] + Finv[6] * W[4];
    B[0] = A[0] * Finv[0] + A[5] * Finv[5] + A[4] * Finv[4];
    B[1] = A[8] * Finv[8] + A[1] * Finv[1] + A[3] * Finv[3];
    B[2] = A[7] * Finv[7] + A[6] * Finv[6] + A[2] * Finv[2];
    B[3] = A[8] * Finv[7] + A[1] * Finv[6] + A[3] * Finv[2];
    B[4] = A[0] * Finv[7] + A[5] * Finv[6] + A[4] * Finv[2];
    B[5] = A[0] * Finv[8] + A[5] * Finv[1] + A[4] * Finv[3];
    B[6] = A[7] * Finv[8] + A[6] * Finv[1] + A[2] * Finv[3];
    B[7] = A[7] * Finv[0] + A[6] * Finv[5] + A[2] * Finv[4];
    B[8] = A[8] * Finv[0] + A[1] * Finv[5] + A[3] * Finv[4];
}

#define ACC(dst, val) dst = gamma ? (dst + (val)) : (val)

struct LawCtx {
    const double* orient;   // 3 doubles (transversely isotropic axis at this voxel) or null
    const double* orient0;  // axis at voxel 0: LinearTransverselyIsotropic::dPK1 passes index m = 0 (fg:11582), kept for parity
    int* flag;
};

// ---- P = alpha*P(F) [+ P]   (MaterialLaw::PK1) --------------------------------------------------
template <int D>
__device__ __forceinline__ void law_PK1(const LawDev& L, const LawCtx& lc, const double* F, double alpha, bool gamma, double* P) {
    switch (L.id) {
        case FGB_LAW_ISO: {                                            // fg:11375-11396
            const double two_mu = 2 * alpha * L.p[0];
            const double ltr = alpha * L.p[1] * (F[0] + F[1] + F[2]);
            ACC(P[0], F[0] * two_mu + ltr);
            ACC(P[1], F[1] * two_mu + ltr);
            ACC(P[2], F[2] * two_mu + ltr);
            if (D >= 6) {
                ACC(P[3], F[3] * two_mu);
                ACC(P[4], F[4] * two_mu);
                ACC(P[5], F[5] * two_mu);
            }
        } break;
        case FGB_LAW_GENERAL: {                                        // fg:11254-11275
            if (D >= 6) {
#pragma unroll
                for (int i = 0; i < 6; i++) {
                    const double* C = L.p + 6 * i;
                    ACC(P[i], alpha * (F[0] * C[0] + F[1] * C[1] + F[2] * C[2] + 2.0 * (F[3] * C[3] + F[4] * C[4] + F[5] * C[5])));
                }
            }
        } break;
        case FGB_LAW_TISO: {                                           // fg:11532-11570
            if (D >= 6) {
                const double two_mu = L.p[0], lam = L.p[1], al = L.p[2], be = L.p[3], two_dmu = L.p[4];
                // constant axis (have_a, fg:11516) in p[5..7] if non-zero, else the orientation field
                const bool have_a = (L.p[5] != 0 || L.p[6] != 0 || L.p[7] != 0);
                const double a0 = have_a ? L.p[5] : lc.orient[0], a1 = have_a ? L.p[6] : lc.orient[1], a2 = have_a ? L.p[7] : lc.orient[2];
                const double A[6] = {a0 * a0, a1 * a1, a2 * a2, a1 * a2, a0 * a2, a0 * a1};
                const double tr = F[0] + F[1] + F[2];
                const double aea = A[0] * F[0] + A[1] * F[1] + A[2] * F[2] + 2 * (A[3] * F[3] + A[4] * F[4] + A[5] * F[5]);
                const double cI = lam * tr + al * aea;
                const double cA = al * tr + be * aea;
                // (A e + e A) in symmetric storage
                const double e[3][3] = {{F[0], F[5], F[4]}, {F[5], F[1], F[3]}, {F[4], F[3], F[2]}};
                const double Am[3][3] = {{A[0], A[5], A[4]}, {A[5], A[1], A[3]}, {A[4], A[3], A[2]}};
                const int r[6] = {0, 1, 2, 1, 0, 0}, c[6] = {0, 1, 2, 2, 2, 1};
#pragma unroll
                for (int i = 0; i < 6; i++) {
                    double s = 0;
#pragma unroll
                    for (int m = 0; m < 3; m++) s += Am[r[i]][m] * e[m][c[i]] + e[r[i]][m] * Am[m][c[i]];
                    const double v = two_mu * F[i] + (i < 3 ? cI : 0.0) + cA * A[i] + two_dmu * s;
                    ACC(P[i], alpha * v);
                }
            }
        } break;
        case FGB_LAW_SCALAR: {                                         // fg:11182-11197
            const double am = alpha * L.p[0];
#pragma unroll
            for (int m = 0; m < D; m++) ACC(P[m], F[m] * am);
        } break;
        case FGB_LAW_ANISO3: {                                         // fg:11114-11129
            const double c11 = L.p[0], c22 = L.p[1], c33 = L.p[2], c23 = L.p[3], c13 = L.p[4], c12 = L.p[5];
            ACC(P[0], alpha * (c11 * F[0] + c12 * F[1] + c13 * F[2]));
            ACC(P[1], alpha * (c12 * F[0] + c22 * F[1] + c23 * F[2]));
            ACC(P[2], alpha * (c13 * F[0] + c23 * F[1] + c33 * F[2]));
        } break;
        case FGB_LAW_SVK: {                                            // fg:11621-11661
            if (D == 9) {
                double C[6], S[6], T[9];
                right_cauchy_green(F, C);
                const double E[6] = {0.5 * (C[0] - 1), 0.5 * (C[1] - 1), 0.5 * (C[2] - 1), 0.5 * C[3], 0.5 * C[4], 0.5 * C[5]};
                const double two_mu = 2 * alpha * L.p[0];
                const double ltr = alpha * L.p[1] * (E[0] + E[1] + E[2]);
                S[0] = E[0] * two_mu + ltr; S[1] = E[1] * two_mu + ltr; S[2] = E[2] * two_mu + ltr;
                S[3] = E[3] * two_mu; S[4] = E[4] * two_mu; S[5] = E[5] * two_mu;
                FS(F, S, T);
#pragma unroll
                for (int i = 0; i < 9; i++) ACC(P[i], T[i]);
            }
        } break;
        case FGB_LAW_NH: {                                             // fg:11744-11787
            if (D == 9) {
                double C[6], Ci[6], S[6], T[9];
                right_cauchy_green(F, C);
                sym_inv6(C, Ci);
                const double J = det9(F);
                const double all = alpha * L.p[1] * checked_log(J, lc.flag);
                const double am = alpha * L.p[0];
#pragma unroll
                for (int m = 0; m < 6; m++) S[m] = am * ((m < 3 ? 1.0 : 0.0) - Ci[m]) + all * Ci[m];
                FS(F, S, T);
#pragma unroll
                for (int i = 0; i < 9; i++) ACC(P[i], T[i]);
            }
        } break;
        case FGB_LAW_NH2: {                                            // fg:11891-11919
            if (D == 9) {
                double Fi[9];
                inv9(F, Fi);
                const double trC = dot9(F, F);
                const double J = det9(F);
                const double p23 = pow(J, -2.0 / 3.0);
                if (isnan(p23)) flag_numeric(lc.flag);
                const double muJ23 = alpha * L.p[0] * p23;
                const double Dc = alpha * L.p[1] * J * (J - 1) - muJ23 * (1.0 / 3.0) * trC;
                const int Ti[9] = {0, 1, 2, 6, 7, 8, 3, 4, 5};
#pragma unroll
                for (int i = 0; i < 9; i++) ACC(P[i], muJ23 * F[i] + Dc * Fi[Ti[i]]);
            }
        } break;
    }
}

// ---- dP = alpha*dP/dF(F):W [+ dP]   (MaterialLaw::dPK1) -----------------------------------------
template <int D>
__device__ __forceinline__ void law_dPK1(const LawDev& L, const LawCtx& lc, const double* F, double alpha, bool gamma, const double* W,
                                         double* dP) {
    switch (L.id) {
        case FGB_LAW_TISO: {
            LawCtx l0 = lc;
            l0.orient = lc.orient0;                                    // quirk fg:11582
            law_PK1<D>(L, l0, W, alpha, gamma, dP);
        } break;
        case FGB_LAW_ISO:
        case FGB_LAW_GENERAL:
        case FGB_LAW_SCALAR:
        case FGB_LAW_ANISO3:
            law_PK1<D>(L, lc, W, alpha, gamma, dP);                     // linear laws: tangent action == law
            break;
        case FGB_LAW_SVK: {                                            // fg:11663-11718
            if (D == 9) {
                double C[6], S[6], dE[6], dS[6], T1[9], T2[9];
                right_cauchy_green(F, C);
                const double E[6] = {0.5 * (C[0] - 1), 0.5 * (C[1] - 1), 0.5 * (C[2] - 1), 0.5 * C[3], 0.5 * C[4], 0.5 * C[5]};
                const double two_mu = 2 * alpha * L.p[0];
                const double ltr = alpha * L.p[1] * (E[0] + E[1] + E[2]);
                S[0] = E[0] * two_mu + ltr; S[1] = E[1] * two_mu + ltr; S[2] = E[2] * two_mu + ltr;
                S[3] = E[3] * two_mu; S[4] = E[4] * two_mu; S[5] = E[5] * two_mu;
                green_strain_deriv(F, W, dE);
                const double ltrd = alpha * L.p[1] * (dE[0] + dE[1] + dE[2]);
                dS[0] = dE[0] * two_mu + ltrd; dS[1] = dE[1] * two_mu + ltrd; dS[2] = dE[2] * two_mu + ltrd;
                dS[3] = dE[3] * two_mu; dS[4] = dE[4] * two_mu; dS[5] = dE[5] * two_mu;
                FS(F, dS, T1);
                FS(W, S, T2);
#pragma unroll
                for (int i = 0; i < 9; i++) ACC(dP[i], T1[i] + T2[i]);
            }
        } break;
        case FGB_LAW_NH: {                                             // fg:11789-11856
            if (D == 9) {
                const int Ti[9] = {0, 1, 2, 6, 7, 8, 3, 4, 5};
                double Fi[9], A[9], B[9];
                const double J = det9(F);
                const double c_m = alpha * (L.p[0] - L.p[1] * checked_log(J, lc.flag));
                const double am = alpha * L.p[0];
                inv9(F, Fi);
                nh_products(Fi, W, A, B);
                const double c_tr = alpha * L.p[1] * (A[0] + A[1] + A[2]);
#pragma unroll
                for (int k = 0; k < 9; k++) ACC(dP[k], am * W[k] + c_tr * Fi[Ti[k]] + c_m * B[k]);
            }
        } break;
        case FGB_LAW_NH2: {                                            // fg:11921-11990
            if (D == 9) {
                const int Ti[9] = {0, 1, 2, 6, 7, 8, 3, 4, 5};
                double Fi[9], FiT[9], A[9], B[9];
                inv9(F, Fi);
#pragma unroll
                for (int k = 0; k < 9; k++) FiT[k] = Fi[Ti[k]];
                const double J = det9(F);
                const double trC3 = (1.0 / 3.0) * dot9(F, F);
                const double p23 = pow(J, -2.0 / 3.0);
                if (isnan(p23)) flag_numeric(lc.flag);
                const double a_muJ23 = alpha * L.p[0] * p23;
                const double a_KJ = alpha * L.p[1] * J;
                const double a_KJJ1 = a_KJ * (J - 1);
                const double a_KJJ = a_KJ * (2 * J - 1);
                nh_products(Fi, W, A, B);
                const double tr = A[0] + A[1] + A[2];
                const double FW23 = (2.0 / 3.0) * dot9(F, W);
                const double FiTW = dot9(FiT, W);
#pragma unroll
                for (int k = 0; k < 9; k++)
                    ACC(dP[k], a_muJ23 * (-2.0 / 3.0 * tr * (F[k] - trC3 * Fi[Ti[k]]) + W[k] - FW23 * FiT[k] + trC3 * B[k]) +
                                   a_KJJ * FiTW * Fi[Ti[k]] - a_KJJ1 * B[k]);
            }
        } break;
    }
}

// ---- energies (MaterialLaw::W) ------------------------------------------------------------------
template <int D>
__device__ __forceinline__ double law_W(const LawDev& L, const LawCtx& lc, const double* F) {
    switch (L.id) {
        case FGB_LAW_SVK: {                                            // fg:11613-11619
            double C[6];
            right_cauchy_green(F, C);
            const double E[6] = {0.5 * (C[0] - 1), 0.5 * (C[1] - 1), 0.5 * (C[2] - 1), 0.5 * C[3], 0.5 * C[4], 0.5 * C[5]};
            const double tr = E[0] + E[1] + E[2];
            const double EE = E[0] * E[0] + E[1] * E[1] + E[2] * E[2] + 2 * (E[3] * E[3] + E[4] * E[4] + E[5] * E[5]);
            return 0.5 * L.p[1] * tr * tr + L.p[0] * EE;
        }
        case FGB_LAW_NH: {                                             // fg:11744-11751
            const double trC = dot9(F, F);
            const double logJ = checked_log(det9(F), lc.flag);
            return 0.5 * (L.p[0] * ((trC - 3.0) - 2.0 * logJ) + L.p[1] * logJ * logJ);
        }
        case FGB_LAW_NH2: {                                            // fg:11882-11889
            const double trC = dot9(F, F);
            const double J = det9(F);
            const double J1 = J - 1;
            return 0.5 * (L.p[0] * (pow(J, -2.0 / 3.0) * trC - 3) + L.p[1] * J1 * J1);
        }
        default: {                                                     // linear laws: 0.5 * S.dot(E)
            double S[9];
            law_PK1<D>(L, lc, F, 1.0, false, S);
            double s = 0;
            if (D == 6 && L.id != FGB_LAW_SCALAR) {
                s = S[0] * F[0] + S[1] * F[1] + S[2] * F[2] + 2 * (S[3] * F[3] + S[4] * F[4] + S[5] * F[5]);
            } else {
                s = S[0] * F[0] + S[1] * F[1] + S[2] * F[2];           // Tensor3 dot (fg:11176-11180, fg:11106-11111)
            }
            return 0.5 * s;
        }
    }
}

// ---- laminate jump solve (LaminateMixedMaterialLaw::solve_newton, fg:13157-13454) ----------------
template <int D>
__device__ void laminate_newton(const LawDev& L1, const LawDev& L2, const LawCtx& lc, const LaminateParams& lp, double c1, double c2,
                                const double* n, const double* Fbar, double* F1, double* F2) {
    const int row[9] = {0, 1, 2, 1, 0, 0, 2, 2, 1};
    const int col[9] = {0, 1, 2, 2, 2, 1, 1, 0, 0};
    double Fbarinv[9];
    if (D == 9) inv9(Fbar, Fbarinv);
    double a[3] = {0, 0, 0};
#pragma unroll
    for (int i = 0; i < 9; i++) F1[i] = F2[i] = Fbar[i];
    double W = 0;
    if (D == 9) W = c1 * law_W<D>(L1, lc, F1) + c2 * law_W<D>(L2, lc, F2);
    // dF1/da_k = -c2 e_k (x) n, dF2/da_k = c1 e_k (x) n (symmetrised for D=6, diagonal only for D=3)
    double dF1[3][9], dF2[3][9];
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int i = 0; i < 9; i++) {
            const double v = (row[i] == k) ? n[col[i]] : 0.0;
            dF1[k][i] = -c2 * v;
            dF2[k][i] = c1 * v;
        }
        fix_sym<D>(dF1[k]);
        fix_sym<D>(dF2[k]);
    }
    for (int iter = 0;; iter++) {
        double P1[9], P2[9], g[3], H[6], Hinv[6], Hinvg[3];
        law_PK1<D>(L1, lc, F1, 1.0, false, P1); fix_dim<D>(P1);
        law_PK1<D>(L2, lc, F2, 1.0, false, P2); fix_dim<D>(P2);
#pragma unroll
        for (int k = 0; k < 3; k++) g[k] = c1 * dot9(P1, dF1[k]) + c2 * dot9(P2, dF2[k]);
        const double g_norm = sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
        if (g_norm <= lp.eps_g) break;
#pragma unroll
        for (int i = 0; i < 6; i++) {
            const int k = row[i], l = col[i];
            double dP1[9], dP2[9];
            law_dPK1<D>(L1, lc, F1, 1.0, false, dF1[l], dP1); fix_dim<D>(dP1);
            law_dPK1<D>(L2, lc, F2, 1.0, false, dF2[l], dP2); fix_dim<D>(dP2);
            H[i] = c1 * dot9(dP1, dF1[k]) + c2 * dot9(dP2, dF2[k]);
        }
        sym_inv6(H, Hinv);
        Hinvg[0] = Hinv[0] * g[0] + Hinv[5] * g[1] + Hinv[4] * g[2];
        Hinvg[1] = Hinv[5] * g[0] + Hinv[1] * g[1] + Hinv[3] * g[2];
        Hinvg[2] = Hinv[4] * g[0] + Hinv[3] * g[1] + Hinv[2] * g[2];
        const double gTda = Hinvg[0] * g[0] + Hinvg[1] * g[1] + Hinvg[2] * g[2];
        const double da_norm = sqrt(Hinvg[0] * Hinvg[0] + Hinvg[1] * Hinvg[1] + Hinvg[2] * Hinvg[2]);
        if (da_norm <= lp.eps_a) break;
        double t = 1;
        if (D == 9 && lp.project_t) {
            // w = n . Fbar^-1 Hinvg, x = n . Fbar^-1 a
            const double M[3][3] = {{Fbarinv[0], Fbarinv[5], Fbarinv[4]}, {Fbarinv[8], Fbarinv[1], Fbarinv[3]}, {Fbarinv[7], Fbarinv[6], Fbarinv[2]}};
            double w = 0, x = 0;
#pragma unroll
            for (int r = 0; r < 3; r++) {
                w += (M[r][0] * Hinvg[0] + M[r][1] * Hinvg[1] + M[r][2] * Hinvg[2]) * n[r];
                x += (M[r][0] * a[0] + M[r][1] * a[1] + M[r][2] * a[2]) * n[r];
            }
            if (w > 0) t = fmin(1.0, (x + lp.delta / c1) / w);
            else if (w < 0) t = fmin(1.0, (x - lp.delta / c2) / w);
        }
        double a_next[3], W_next = 0;
        for (;;) {
#pragma unroll
            for (int i = 0; i < 3; i++) a_next[i] = a[i] - t * Hinvg[i];
#pragma unroll
            for (int i = 0; i < 9; i++) {
                F1[i] = Fbar[i] - c2 * a_next[row[i]] * n[col[i]];
                F2[i] = Fbar[i] + c1 * a_next[row[i]] * n[col[i]];
            }
            fix_sym<D>(F1);
            fix_sym<D>(F2);
            if (D != 9) return;                                   // linear: one full step (fg:13367-13370)
            W_next = c1 * law_W<D>(L1, lc, F1) + c2 * law_W<D>(L2, lc, F2);
            if (!lp.backtrack) break;
            if (W_next < W - lp.alpha * t * gTda) break;
            t *= lp.beta;
            if (t <= lp.eps_t) break;
        }
        if (t <= lp.eps_t) break;
        a[0] = a_next[0]; a[1] = a_next[1]; a[2] = a_next[2];
        W = W_next;
        if (iter >= lp.maxiter) break;
    }
}

// inverse of a D x D row-major matrix by Gauss-Jordan with partial pivoting (InvertMatrix fg:1143); A is destroyed
template <int D>
__device__ void invert_dense(double* A, double* R) {
    for (int i = 0; i < D; i++)
        for (int j = 0; j < D; j++) R[i * D + j] = (i == j) ? 1.0 : 0.0;
    for (int c = 0; c < D; c++) {
        int piv = c;
        double best = fabs(A[c * D + c]);
        for (int r = c + 1; r < D; r++) {
            const double v = fabs(A[r * D + c]);
            if (v > best) { best = v; piv = r; }
        }
        if (piv != c) {
            for (int k = 0; k < D; k++) {
                double t = A[c * D + k]; A[c * D + k] = A[piv * D + k]; A[piv * D + k] = t;
                t = R[c * D + k]; R[c * D + k] = R[piv * D + k]; R[piv * D + k] = t;
            }
        }
        const double id = 1.0 / A[c * D + c];
        for (int k = 0; k < D; k++) { A[c * D + k] *= id; R[c * D + k] *= id; }
        for (int r = 0; r < D; r++) {
            if (r == c) continue;
            const double f = A[r * D + c];
            if (f == 0) continue;
            for (int k = 0; k < D; k++) { A[r * D + k] -= f * A[c * D + k]; R[r * D + k] -= f * R[c * D + k]; }
        }
    }
}

// ---- mixed law at voxel offset o -----------------------------------------------------------------
template <int D>
struct Mixed {
    // ReussMixedMaterialLaw (fg:12653-12724): harmonic mean of the phase tangents on mixed voxels; returns the pure phase
    // index (>= 0) if the voxel is pure, else -1 with Ceff filled (row-major, as the reference stores it)
    __device__ static int reuss_matrix(const MaterialDev& M, size_t o, const LawCtx& lc, const double* F, double* Ceff) {
        double Sum[D * D], C[D * D], Ci[D * D];
        for (int i = 0; i < D * D; i++) Sum[i] = 0;
        for (int p = 0; p < M.nphases; p++) {
            const double phi = M.phi[p][o];
            if (phi == 0) continue;
            if (phi == 1) return p;
            for (int m = 0; m < D; m++) {
                double e[9], r[9];
#pragma unroll
                for (int k = 0; k < 9; k++) e[k] = (k == m) ? 1.0 : 0.0;
                law_dPK1<D>(M.law[p], lc, F, 1.0, false, e, r);
                for (int k = 0; k < D; k++) C[m * D + k] = r[k];
            }
            invert_dense<D>(C, Ci);
            for (int i = 0; i < D * D; i++) Sum[i] += phi * Ci[i];
        }
        invert_dense<D>(Sum, Ceff);
        return -1;
    }

    // phase selection of the laminate rule (get_mix fg:13456-13525): returns number of phases (1 or 2)
    __device__ static int laminate_split(const MaterialDev& M, size_t o, const LawCtx& lc, const double* F, int& p1, int& p2, double& c1,
                                         double& c2, double* F1, double* F2) {
        p1 = -1; p2 = -1; c1 = 0; c2 = 0;
        for (int p = 0; p < M.nphases; p++) {
            const double phi = M.phi[p][o];
            if (phi == 0) continue;
            if (phi == 1) { p1 = p; c1 = phi; p2 = -1; break; }
            if (p1 < 0) { p1 = p; c1 = phi; continue; }
            if (p2 < 0) { p2 = p; c2 = phi; continue; }
            flag_numeric(lc.flag);     // more than two phases in a voxel (fg:13473)
        }
        if (p1 < 0) { flag_numeric(lc.flag); p1 = 0; }
#pragma unroll
        for (int i = 0; i < 9; i++) F1[i] = (i < D) ? F[i] : 0.0;
        if (p2 < 0) return 1;
        double n[3] = {M.normals[0][o], M.normals[1][o], M.normals[2][o]};
        double Fbar[9];
#pragma unroll
        for (int i = 0; i < 9; i++) Fbar[i] = (i < D) ? F[i] : 0.0;
        fix_dim<D>(Fbar);
        if (M.lam.fixed_c1 > 0) c1 = M.lam.fixed_c1;
        c2 = 1.0 - c1;
        laminate_newton<D>(M.law[p1], M.law[p2], lc, M.lam, c1, c2, n, Fbar, F1, F2);
        return 2;
    }

    __device__ static LawCtx make_ctx(const MaterialDev& M, size_t o, double* abuf, int* flag) {
        LawCtx lc;
        lc.flag = flag;
        lc.orient = nullptr;
        lc.orient0 = nullptr;
        if (M.orient[0]) {
            abuf[0] = M.orient[0][o]; abuf[1] = M.orient[1][o]; abuf[2] = M.orient[2][o];
            abuf[3] = M.orient[0][0]; abuf[4] = M.orient[1][0]; abuf[5] = M.orient[2][0];
            lc.orient = abuf;
            lc.orient0 = abuf + 3;
        }
        return lc;
    }

    // P = alpha * P_mix(F)
    __device__ static void PK1(const MaterialDev& M, size_t o, const double* F, double alpha, double* P, int* flag) {
        double abuf[6];
        const LawCtx lc = make_ctx(M, o, abuf, flag);
        if (M.mix == FGB_MIX_VOIGT) {                                  // fg:12752-12761
            bool gamma = false;
            for (int p = 0; p < M.nphases; p++) {
                const double phi = M.phi[p][o];
                if (phi <= FGB_VOIGT_THRESHOLD) continue;
                law_PK1<D>(M.law[p], lc, F, phi * alpha, gamma, P);
                gamma = true;
            }
        } else if (M.mix == FGB_MIX_LAMINATE) {                        // fg:13543-13558
            int p1, p2; double c1, c2, F1[9], F2[9];
            const int np = laminate_split(M, o, lc, F, p1, p2, c1, c2, F1, F2);
            law_PK1<D>(M.law[p1], lc, F1, c1 * alpha, false, P);
            if (np == 2) law_PK1<D>(M.law[p2], lc, F2, c2 * alpha, true, P);
        } else if (M.mix == FGB_MIX_REUSS) {                           // fg:12663-12688
            double C[D * D];
            const int pure = reuss_matrix(M, o, lc, F, C);
            if (pure >= 0) { law_PK1<D>(M.law[pure], lc, F, alpha, false, P); return; }
            for (int k = 0; k < D; k++) {
                double s = 0;
                for (int j = 0; j < D; j++) s += alpha * C[k * D + j] * F[j];
                P[k] = s;
            }
        }
    }

    // dP = alpha * dP_mix/dF(F) : W
    __device__ static void dPK1(const MaterialDev& M, size_t o, const double* F, double alpha, const double* W, double* dP, int* flag) {
        double abuf[6];
        const LawCtx lc = make_ctx(M, o, abuf, flag);
        if (M.mix == FGB_MIX_VOIGT) {                                  // fg:12763-12771
            bool gamma = false;
            for (int p = 0; p < M.nphases; p++) {
                const double phi = M.phi[p][o];
                if (phi <= FGB_VOIGT_THRESHOLD) continue;
                law_dPK1<D>(M.law[p], lc, F, phi * alpha, gamma, W, dP);
                gamma = true;
            }
        } else if (M.mix == FGB_MIX_LAMINATE) {                        // fg:13599-13625 (tangent "approx")
            int p1, p2; double c1, c2, F1[9], F2[9];
            const int np = laminate_split(M, o, lc, F, p1, p2, c1, c2, F1, F2);
            law_dPK1<D>(M.law[p1], lc, F1, c1 * alpha, false, W, dP);
            if (np == 2) law_dPK1<D>(M.law[p2], lc, F2, c2 * alpha, true, W, dP);
        } else if (M.mix == FGB_MIX_REUSS) {                           // fg:12690-12718
            double C[D * D];
            const int pure = reuss_matrix(M, o, lc, F, C);
            if (pure >= 0) { law_dPK1<D>(M.law[pure], lc, F, alpha, false, W, dP); return; }
            for (int k = 0; k < D; k++) {
                double s = 0;
                for (int j = 0; j < D; j++) s += alpha * C[k * D + j] * W[j];
                dP[k] = s;
            }
        }
    }

    __device__ static double W(const MaterialDev& M, size_t o, const double* F, int* flag) {
        double abuf[6];
        const LawCtx lc = make_ctx(M, o, abuf, flag);
        double Fx[9];
#pragma unroll
        for (int i = 0; i < 9; i++) Fx[i] = (i < D) ? F[i] : 0.0;
        if (M.mix == FGB_MIX_VOIGT) {                                  // fg:12739-12750
            double w = 0;
            for (int p = 0; p < M.nphases; p++) {
                const double phi = M.phi[p][o];
                if (phi <= FGB_VOIGT_THRESHOLD) continue;
                w += phi * law_W<D>(M.law[p], lc, Fx);
            }
            return w;
        }
        int p1, p2; double c1, c2, F1[9], F2[9];                       // fg:13527-13541
        const int np = laminate_split(M, o, lc, F, p1, p2, c1, c2, F1, F2);
        double w = c1 * law_W<D>(M.law[p1], lc, F1);
        if (np == 2) w += c2 * law_W<D>(M.law[p2], lc, F2);
        return w;
    }
};
