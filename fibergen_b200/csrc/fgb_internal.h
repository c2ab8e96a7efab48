// Internal declarations shared by the CUDA translation units of libfgb200.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <map>
#include "../../include/fgb200.h"

#define FGB_MAX_STAGES 24
#define FGB_CG_RING 8
#define FGB_MAX_LAW_PARAMS 36

struct FftPlanDev {
    int n;
    int nstages;
    int radix[FGB_MAX_STAGES];
    const double2* tw;     // exp(-2 pi i k / n), k = 0..n-1 (device)
};

struct LawDev {
    int id;
    double p[FGB_MAX_LAW_PARAMS];
};

struct LaminateParams {
    double eps_t, eps_a, eps_g, alpha, beta, delta, fixed_c1;
    int maxiter, backtrack, project_t;
};

// everything the material kernels need, passed by value
struct MaterialDev {
    int nphases;
    int mix;
    const double* phi[FGB_MAX_PHASES];
    LawDev law[FGB_MAX_PHASES];
    const double* normals[3];
    const double* orient[3];
    LaminateParams lam;
};

struct GridDev {
    int nx, ny, nz, nzc, nzp;   // global sizes
    int lnx, x0;                // local slab
    size_t plane;               // lnx*ny*nzp (doubles per component)
    int unzcs;                  // complex row stride of the u buffer (nzc rounded up to a multiple of 8: 128-byte aligned rows)
    size_t uplane;              // lnx*ny*2*unzcs (doubles per u component)
    double hx, hy, hz;          // nx/Lx ... (inverse voxel size, fg:18618-18620)
};

struct ProfEntry {
    double ms = 0;
    uint64_t launches = 0;
};

struct fgb_ctx {
    GridDev g;
    double L[3];
    int mode, scheme, dim, udim;
    int device;
    int rank, nranks;
    int sm_count;
    std::map<const void*, int> occupancy;      // resident CTAs per SM of a kernel at its launch block size (fgb_wave_grid)
    int implicit_w_of;            // field id p whose operator result w = sym-grad(u) is held implicitly in ubuf (-1: none)
    size_t smem_optin;
    cudaStream_t stream, own_stream;
    std::string err;

    double* fields[FGB_MAX_FIELDS];
    double* ubuf;               // udim planes (staggered displacement / rhs)
    double* phi[FGB_MAX_PHASES];
    double* normals;            // 3 planes or null
    double* orient;             // 3 planes or null
    int nphases;
    int mix;
    LawDev laws[FGB_MAX_PHASES];
    LaminateParams lam;
    int freq_hack;

    // FFT plans + frequency tables
    FftPlanDev plan[3];         // x, y, z
    double2* tw_dev[3];
    double* kpm_dev[3];         // staggered: sin(xi)/h per index     (fg:19856-19876)
    double2* kp_dev[3];         // staggered: kpm*exp(i xi)
    double* xi_dev[3];          // collocated: m/L per index           (fg:19393-19406)
    double* xi2pi_dev[3];       // 2 pi m/L per index (G0Div/Grad hyper fg:20159, Willot-R fg:19089)
    double* wtan_dev[3];        // Willot-R: 0.25 tan(q/2), q = 2 pi m/n (fg:19153)
    double2* wex_dev[3];        // Willot-R: 1 + exp(i q)                (fg:19132)
    double* pois_dev[3];        // poisson_solve: (n/L)^2 (cos(2 pi i/n) - 1) (fg:23462-23486)

    // reductions
    double* d_partials;         // [nblocks][32]
    double* d_result;           // [32]
    double* h_result;           // pinned [32]
    int red_blocks;
    double* d_scalars;          // device-resident CG scalars: [0] gamma, [1] beta, [2] alpha, [3] <p, p - w>, [4] delta
    bool cg_dev;                // fgb_cgdev_*: kernels take beta / alpha from d_scalars, reductions stay on the device
    bool reduce_on_device;      // fgb_reduce_finish leaves the (rank-gathered) sums on the device instead of returning them
    double* h_ring;             // pinned [FGB_CG_RING][8]: gamma, <p,p-w>, alpha, delta, numeric-error flag of the last iterations
    cudaEvent_t ring_ev[8];
    int* d_flag;                // numeric error flag
    int* h_flag;

    // multi-GPU
    void* nccl_comm;            // ncclComm_t
    void* nccl_lib;             // dlopen handle of libnccl.so.2
    double* sbuf;               // all-to-all staging, [c][q][il][jl][k] complex (same size as the transformed buffer)
    double* xbuf;               // y-slab layout [c][ii][jl][k] complex: the fused x pass runs in place on it
    size_t xbuf_cap;            // doubles sbuf / xbuf can hold (each)
    double* halo;               // [3 lo slots][3 hi slots] of halo_slot doubles (neighbour x planes for the stencils)
    size_t halo_slot;
    double* d_gather;           // rank-ordered reduction staging
    double* iso_halo;           // fused isotropic sweep: [r_lo 3][p_lo 3][r_hi 2][p_hi 2][phi_lo P][phi_hi P] planes of ny*nzp doubles
    bool phi_halo_valid;
    bool p2p;                   // peer buffers mapped: transposes are written by the FFT kernels straight into peer memory
    double* peer_xbuf[8];       // xbuf of every rank (own pointer at [rank])
    double* halo_base;            // one allocation: [halo set 0][halo set 1][iso set 0][iso set 1]; halo / iso_halo point at the current set
    double* peer_halo[8];         // halo_base of every rank (CUDA IPC), for the direct halo push
    size_t iso_set;               // doubles per iso-halo set
    unsigned halo_seq, iso_seq;   // exchange counters (set = seq & 1)
    double* peer_sbuf[8];
    unsigned long long* sync_base;      // barrier / scalar-exchange block of this rank (comm.cu), mapped by every peer
    unsigned long long* peer_sync[8];
    unsigned long long bar_seq, ex_seq;

    // mixed boundary conditions (fgb_set_bc): row-major dim x dim matrices MQ and M:(QC0)
    bool bc_active;
    double bc_relax;
    double bc_MQ[81], bc_MQC0[81];
    double F00[9];
    double* nh_cache;           // fused Neo-Hooke tangent: Finv (9) + 3 coefficients per voxel at the Newton iterate (fused_hyper.cu)
    int nh_cache_of;            // field id the cache was built from (-1: none)
    double nh_cache_mu0;
    double* heatK;              // fused heat path: per-voxel diagonal of the (linear) mixed law, 3 planes (fused_heat.cu)
    bool heatK_valid;
    int heatK_diag;
    double* visc_tmp;           // copy of tau for the viscosity Delta operator (fg:21316-21320)

    // doubly fine grid of the half_staggered / full_staggered schemes (use_dfg fg:14894): the constitutive sweeps run on a grid
    // with twice the resolution; fields are prolongated before and restricted after (fg:14216-14339, fg:18143-18149, fg:18343-18347)
    int dfg;                    // 0 off, 1 half_staggered (fine phases injected from the coarse grid), 2 full_staggered
    GridDev gf;                 // the fine grid (2nx, 2ny, 2nz)
    double* dfg1;               // _temp_dfg_1: dim fine planes
    double* dfg2;               // _temp_dfg_2 (tangent sweeps)
    double* phi_f[FGB_MAX_PHASES];
    double* normals_f;
    double* orient_f;

    uint64_t launches;
    bool profiling;
    std::map<std::string, ProfEntry> prof;
    cudaEvent_t ev0, ev1;
};

// error helpers -----------------------------------------------------------------------------
int fgb_fail(fgb_ctx* ctx, int code, const char* fmt, ...);
#define FGB_CUDA(ctx, call)                                                                  \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess)                                                              \
            return fgb_fail(ctx, FGB_ECUDA, "%s failed: %s (%s:%d)", #call,                  \
                            cudaGetErrorString(e__), __FILE__, __LINE__);                    \
    } while (0)
#define FGB_CHECK_LAUNCH(ctx, name)                                                          \
    do {                                                                                     \
        (ctx)->launches++;                                                                   \
        cudaError_t e__ = cudaGetLastError();                                                \
        if (e__ != cudaSuccess)                                                              \
            return fgb_fail(ctx, FGB_ECUDA, "launch of %s failed: %s", name,                 \
                            cudaGetErrorString(e__));                                        \
    } while (0)

struct ProfScope {
    fgb_ctx* c;
    const char* name;
    ProfScope(fgb_ctx* ctx, const char* n);
    ~ProfScope();
};

// fft.cu -------------------------------------------------------------------------------------
int fgb_fft_init(fgb_ctx* ctx);
void fgb_fft_free(fgb_ctx* ctx);
// row layout of a buffer being transformed: rows of nzcs complex (2*nzcs doubles), components lnx*ny rows apart
struct FftLayout {
    int nzcs;
};
// where element e of a strided pencil lives: segments of `seglen` elements `segstride` apart, `estride` inside a segment.
// seglen == n is the plain strided layout; seglen = ny/nranks addresses the all-to-all send/receive staging layout.
struct PencilMap {
    long estride;
    int seglen;
    long segstride;
    long ostride;      // per blockIdx.y
    long cstride;      // per component
#ifdef __CUDACC__
    __host__ __device__ __forceinline__ long at(int e) const {
        const int q = e / seglen;
        return (long)q * segstride + (long)(e - q * seglen) * estride;
    }
#endif
};
// destination bases of a store that targets peer GPUs: segment q of a pencil (PencilMap::seglen elements) goes to p[q]
// (mapped peer memory over NVLink, cudaIpcOpenMemHandle); n == 0 means "store locally"
struct PeerTable {
    int n;
    double2* p[8];
};
// in-place r2c/c2r along z of `ncomp` components starting at base
int fgb_fft_z_forward(fgb_ctx* ctx, double* base, int ncomp, const FftLayout& lay);
int fgb_fft_z_backward(fgb_ctx* ctx, double* base, int ncomp, const FftLayout& lay);
int fgb_fft_y(fgb_ctx* ctx, double* base, int ncomp, const FftLayout& lay, int dir);
int fgb_fft_strided(fgb_ctx* ctx, int axis, const double* src, double* dst, const PencilMap& mi, const PencilMap& mo, int ninner,
                    int nouter, int ncomp, int dir, const PeerTable* peers = nullptr);
// x pass; green_kind: 0 none (plain forward or backward per dir), otherwise fused fwd-x, Green, inv-x
struct GreenArgs {
    int kind;             // 0 none, 1 staggered elasticity/hyper (general), 2 staggered heat, 3 colloc elasticity, 4 colloc heat, 5 colloc hyper,
                          // 6 G0-div hyper (9 -> 3), 7 grad hyper (3 -> 9), 8 Willot-R, 9 colloc elasticity on the zero-trace representation, 10 Poisson
    double c10, c20;      // staggered coefficients / collocated c10, c20 (Willot-R: mu_0 and mu_0/lambda_0)
    double beta;          // collocated beta
    double alpha;         // Willot-R only (the other kinds fold alpha into c10, c20)
    double dc[9];         // value of the zero frequency
    int freq_hack;
};
int fgb_fft_x(fgb_ctx* ctx, double* base, int ncomp, const FftLayout& lay, int dir, const GreenArgs* ga);
int fgb_fft_x_green_layout(fgb_ctx* ctx, double* base, const GreenArgs* ga, long estride, int nzc_valid, int nouter, long ostride,
                           long cstride, int jbase, const PencilMap* out_map = nullptr, const PeerTable* peers = nullptr);

// stencil.cu ---------------------------------------------------------------------------------
int fgb_k_div(fgb_ctx* ctx, const double* tau, double* u);
int fgb_k_eps(fgb_ctx* ctx, const double* u, double* eta, const double* Econst /*dim, host*/);
int fgb_k_div_vector(fgb_ctx* ctx, const double* u, double* b, double alpha);   // divVector fg:19983: 3-component u buffer -> one component (u layout)
int fgb_k_prolongate(fgb_ctx* ctx, const double* coarse, double* fine);          // prolongate_to_dfg fg:14216
int fgb_k_restrict(fgb_ctx* ctx, const double* fine, double* coarse);            // restrict_from_dfg fg:14273
int fgb_k_inject_phase(fgb_ctx* ctx, const double* coarse, double* fine);        // initFullStageredRawPhases fg:17648-17680
int fgb_k_mxpy(fgb_ctx* ctx, double* r, const double* x, const double* y);        // mxpyTensor fg:20590 on one component plane

// material.cu --------------------------------------------------------------------------------
MaterialDev fgb_material_dev(fgb_ctx* ctx);
int fgb_k_calc_stress(fgb_ctx* ctx, const double* src, double* dst, double mu0, double lambda0, double alpha);
int fgb_k_calc_stress_deriv(fgb_ctx* ctx, const double* F, const double* W, double* dst, double mu0, double lambda0, double alpha);
int fgb_k_calc_polarization(fgb_ctx* ctx, const double* src, double* dst, double mu0, int inv);
int fgb_k_mean_pk1(fgb_ctx* ctx, const double* src, double alpha, double* out);
int fgb_k_mean_energy(fgb_ctx* ctx, const double* src, double* out);
int fgb_k_mean_cauchy(fgb_ctx* ctx, const double* src, double alpha, double* out);
int fgb_k_min_detF(fgb_ctx* ctx, const double* src, double* out);
int fgb_k_ref_material(fgb_ctx* ctx, const double* src, int zero_trace, double* lmin, double* lmax);

// blas.cu ------------------------------------------------------------------------------------
int fgb_k_set_constant(fgb_ctx* ctx, double* f, const double* c, int add);
int fgb_k_copy(fgb_ctx* ctx, const double* src, double* dst, int ncomp);
int fgb_k_xpay(fgb_ctx* ctx, double* r, const double* x, double a, const double* y);
int fgb_k_xpaymz(fgb_ctx* ctx, double* r, const double* x, double a, const double* y, const double* z);
int fgb_k_adjust_residual(fgb_ctx* ctx, double* r, const double* E, const double* z);
int fgb_k_calc_stress_const(fgb_ctx* ctx, const double* src, double* dst, double mu0, double lambda0);
int fgb_k_inner(fgb_ctx* ctx, const double* a, const double* b, const double* c, double* out);
int fgb_k_component_dot(fgb_ctx* ctx, const double* a, const double* b, double* out, int mean_only);
int fgb_k_cg_update(fgb_ctx* ctx, double* x, double* r, const double* p, const double* w, double a, double* delta);
int fgb_k_extrapolate_poly(fgb_ctx* ctx, int n, const double* const* fields, const double* Vinv, const double* tpowers, double* dst);
// finish a block-partial reduction of `nvals` sums (or mins/maxs) and copy to host; op 0 sum, 1 min, 2 max
int fgb_reduce_finish(fgb_ctx* ctx, int nblocks, int nvals, int op, double* host_out);
// grid of a grid-stride kernel: enough blocks for n items, at most `cap` blocks, rounded down to whole waves of the kernel's resident
// CTAs (148 SMs x occupancy) so that no partial last wave runs at reduced occupancy
unsigned fgb_wave_grid(fgb_ctx* ctx, const void* kernel, int block, size_t n, size_t cap);
int fgb_allreduce_host(fgb_ctx* ctx, double* vals, int n, int op);
int fgb_allgather_dev(fgb_ctx* ctx, int n);       // d_result[0..n) of every rank -> d_gather[rank*n + i], no host synchronisation
// device-resident CG scalars (fgb_cgdev_*): after a sum was left on the device by fgb_reduce_finish,
// mode 0: <p,p-w> = sum/nxyz, alpha = gamma/(<p,p-w> + tiny);  mode 1: delta = sum/nxyz (+tiny), beta = delta/gamma, gamma = delta
int fgb_k_cg_scalars(fgb_ctx* ctx, int mode, int ring_slot);
int fgb_k_xpay_dev(fgb_ctx* ctx, double* r, const double* x, int scal_index, const double* y);   // r = x + d_scalars[i]*y

// fused.cu -----------------------------------------------------------------------------------
int fgb_fused_iso_applicable(const fgb_ctx* ctx);
// p_new = r + cgbeta*p_old (skipped when r == null), f = div (C-C0):p_new -> u buffer
int fgb_k_dir_stress_div_iso(fgb_ctx* ctx, const double* r, double cgbeta, const double* p_old, double* p_new, double mu0, double lambda0,
                             double alpha);
// eta = E + sym-grad u and pAp = <p, p - eta>
int fgb_k_eps_dot(fgb_ctx* ctx, const double* u, double* eta, const double* Econst, const double* p, double* pAp);
int fgb_k_cg_update_implicit(fgb_ctx* ctx, const double* u, const double* Econst, double* x, double* r, const double* p, double a, double* delta);

// fused_heat.cu ------------------------------------------------------------------------------
int fgb_fused_heat_applicable(const fgb_ctx* ctx);
int fgb_heat_tangent(fgb_ctx* ctx, int* diag);
int fgb_k_heat_march(fgb_ctx* ctx, const double* r, double cgbeta, const double* p_old, double* p_new, double mu0, double alpha);
int fgb_k_heat_cg_u(fgb_ctx* ctx, bool dot_only, const double* Econst, double* x, double* r, const double* p, double a, double* out);

// fused_hyper.cu -----------------------------------------------------------------------------
int fgb_fused_nh_applicable(const fgb_ctx* ctx);
int fgb_k_nh_cache(fgb_ctx* ctx, const double* F, double mu0);
int fgb_k_nh_dir_tangent(fgb_ctx* ctx, const double* R, double cgbeta, const double* Q_old, double* Q_new, double* sigma, double lambda0);
int fgb_k_hyper_cg_u(fgb_ctx* ctx, bool dot_only, double* X, double* R, const double* Q, double a, double* out);

// comm.cu ------------------------------------------------------------------------------------
int fgb_comm_halo_heat(fgb_ctx* ctx, const double* r, const double* p_old);   // left neighbour's last plane of r_0, p_0 and K_0
int fgb_comm_free(fgb_ctx* ctx);
// neighbour planes for the fused isotropic sweep: r/p components 0,1,2 of the left neighbour's last plane, components 5,4 of the
// right neighbour's first plane, and (once) the phase fractions of both planes; r may be null
int fgb_comm_halo_iso(fgb_ctx* ctx, const double* r, const double* p_old);
// slab-partitioned x pass: transpose -> fwd x, Green, inv x -> transpose back
int fgb_comm_fft_x(fgb_ctx* ctx, double* base, int ncomp, const FftLayout& lay, const GreenArgs* ga);
int fgb_comm_halo_tau(fgb_ctx* ctx, const double* tau);   // fills ctx->halo for k_div
int fgb_comm_halo_u(fgb_ctx* ctx);                        // fills ctx->halo for k_eps
