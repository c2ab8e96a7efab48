// C ABI of libfgb200 (include/fgb200.h): context, fields, setup and the operator / scheme entry points.
#include "fgb_internal.h"
#include <cstdarg>
#include <cmath>
#include <nvtx3/nvToolsExt.h>

// NVTX range of a kernel scope: named after the reference's Timer of the member(s) the kernel replaces (fg:1643-1737,
// e.g. "calc stress" fg:18136, "fftVector" fg:18483, "G0OperatorFourierStaggered" fg:19836), so that a timeline reads like the
// reference's <print_timings/> table.  Fused kernels carry all the names they cover.
static const char* nvtx_name(const char* scope) {
    static const struct { const char* scope; const char* timer; } map[] = {
        {"calc_stress", "calc stress"}, {"calc_stress_deriv", "calc stress deriv"}, {"calc_stress_const", "calcStressConst"},
        {"calc_polarization", "calc polarization"}, {"div_staggered", "divOperatorStaggered"}, {"eps_staggered", "epsOperatorStaggered"},
        {"div_vector", "divVector"}, {"fft_z_r2c", "fftVector / forward FFT double (z)"}, {"fft_y_fwd", "fftVector / forward FFT double (y)"},
        {"fft_y_fwd_p2p", "fftVector / forward FFT double (y) + slab transpose"}, {"fft_x_fwd", "fftTensor / forward FFT double (x)"},
        {"fft_x_bwd", "fftInvTensor / backward FFT double (x)"},
        {"fft_x_green", "forward FFT (x) + G0OperatorFourierStaggered / GammaOperatorFourierCollocated + backward FFT (x)"},
        {"fft_y_bwd", "fftInvVector / backward FFT double (y)"}, {"fft_z_c2r", "fftInvVector / backward FFT double (z)"},
        {"inner_product", "innerProductL2"}, {"component_dot", "component_dot"}, {"xpay", "xpay"}, {"xpaymz", "xpaymz"},
        {"set_constant", "setConstant"}, {"adjust_residual", "adjustResidual"}, {"cg_update", "xpay + xpaymz + innerProductL2"},
        {"cg_direction_stress_div", "xpay + calc stress + divOperatorStaggered"}, {"stress_div", "calc stress + divOperatorStaggered"},
        {"eps_dot", "epsOperatorStaggered + innerProductL2"}, {"eps_dot_implicit", "epsOperatorStaggered + innerProductL2 (implicit)"},
        {"cg_update_implicit", "epsOperatorStaggered + xpay + xpaymz + innerProductL2"},
        {"heat_dir_flux_div", "xpay + calc stress + divOperatorStaggeredHeat"}, {"heat_flux_div", "calc stress + divOperatorStaggeredHeat"},
        {"nh_dir_tangent", "xpay + calc stress deriv"}, {"nh_tangent", "calc stress deriv"}, {"nh_tangent_cache", "calc stress deriv (F-dependent part)"},
        {"mean_pk1", "meanPK1"}, {"mean_energy", "meanW"}, {"mean_cauchy", "meanCauchy"}, {"ref_material", "getRefMaterial"},
        {"prolongate_to_dfg", "prolongate_to_dfg"}, {"restrict_from_dfg", "restrict_from_dfg"}, {"init_phase", "phase initialization"},
        {"halo_exchange", "halo exchange (slab partition)"}, {"heat_tangent", "laminate solve_newton (once per phase change)"},
    };
    for (const auto& m : map)
        if (!strcmp(m.scope, scope)) return m.timer;
    return scope;
}

static std::string g_create_error;

int fgb_fail(fgb_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    else g_create_error = buf;
    return code;
}

// ---- profiling: event pairs recorded on the launching stream, resolved lazily ---------------------
struct ProfPending {
    const char* name;
    cudaEvent_t e0, e1;
};
static std::map<fgb_ctx*, std::vector<ProfPending>> g_pending;
static std::map<fgb_ctx*, std::vector<cudaEvent_t>> g_event_pool;

static cudaEvent_t get_event(fgb_ctx* c) {
    auto& pool = g_event_pool[c];
    if (!pool.empty()) {
        cudaEvent_t e = pool.back();
        pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

ProfScope::ProfScope(fgb_ctx* ctx, const char* n) : c(ctx), name(n) {
    nvtxRangePushA(nvtx_name(n));
    if (!c->profiling) return;
    ProfPending p;
    p.name = n;
    p.e0 = get_event(c);
    p.e1 = get_event(c);
    cudaEventRecord(p.e0, c->stream);
    g_pending[c].push_back(p);
}
ProfScope::~ProfScope() {
    nvtxRangePop();
    if (!c->profiling) return;
    cudaEventRecord(g_pending[c].back().e1, c->stream);
}

static void prof_resolve(fgb_ctx* c) {
    auto it = g_pending.find(c);
    if (it == g_pending.end()) return;
    cudaStreamSynchronize(c->stream);
    for (auto& p : it->second) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) {
            ProfEntry& e = c->prof[p.name];
            e.ms += ms;
            e.launches++;
        }
        g_event_pool[c].push_back(p.e0);
        g_event_pool[c].push_back(p.e1);
    }
    it->second.clear();
}

// ---- context --------------------------------------------------------------------------------------
extern "C" const char* fgb_version(void) { return "fgb200 0.1 (sm_100a)"; }

extern "C" const char* fgb_last_error(const fgb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int fgb_create(fgb_ctx** out, int nx, int ny, int nz, double Lx, double Ly, double Lz, int mode, int gamma_scheme,
                          int device, int rank, int nranks) {
    if (!out) return FGB_EINVAL;
    *out = nullptr;
    if (nx < 1 || ny < 1 || nz < 1 || !(Lx > 0) || !(Ly > 0) || !(Lz > 0))
        return fgb_fail(nullptr, FGB_EINVAL, "invalid grid %dx%dx%d / box %g x %g x %g", nx, ny, nz, Lx, Ly, Lz);
    if (mode < FGB_MODE_ELASTICITY || mode > FGB_MODE_POROUS) return fgb_fail(nullptr, FGB_EINVAL, "unknown mode %d", mode);
    if (gamma_scheme != FGB_GAMMA_COLLOCATED && gamma_scheme != FGB_GAMMA_STAGGERED && gamma_scheme != FGB_GAMMA_WILLOT)
        return fgb_fail(nullptr, FGB_EUNSUPPORTED, "gamma scheme %d not supported (collocated, staggered and willot only)", gamma_scheme);
    if (gamma_scheme == FGB_GAMMA_WILLOT && mode != FGB_MODE_ELASTICITY && mode != FGB_MODE_VISCOSITY)
        return fgb_fail(nullptr, FGB_EINVAL, "Unknown gamma scheme 'willot' for this mode (fg:20488-20531: elasticity and viscosity only)");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fgb_fail(nullptr, FGB_EINVAL, "bad rank %d of %d", rank, nranks);
    if (nranks > 1 && (nx % nranks || ny % nranks))
        return fgb_fail(nullptr, FGB_EUNSUPPORTED, "slab partition needs nx and ny divisible by the number of ranks");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fgb_fail(nullptr, FGB_ENODEV, "no CUDA device available (this library has no CPU fallback)");
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) return fgb_fail(nullptr, FGB_ENODEV, "cudaGetDevice failed");
    }
    if (device >= ndev) return fgb_fail(nullptr, FGB_ENODEV, "device %d out of range (%d devices)", device, ndev);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fgb_fail(nullptr, FGB_ENODEV, "cannot query device %d", device);
    if (prop.major != 10)
        return fgb_fail(nullptr, FGB_ENODEV, "device %d is sm_%d%d; this library is built for sm_100a only (no fallback)", device,
                        prop.major, prop.minor);
    if (cudaSetDevice(device) != cudaSuccess) return fgb_fail(nullptr, FGB_ENODEV, "cudaSetDevice(%d) failed", device);

    fgb_ctx* c = new fgb_ctx();
    c->g.nx = nx; c->g.ny = ny; c->g.nz = nz;
    c->g.nzc = nz / 2 + 1;
    c->g.nzp = 2 * c->g.nzc;
    c->g.lnx = nx / nranks;
    c->g.x0 = rank * c->g.lnx;
    c->g.plane = (size_t)c->g.lnx * ny * c->g.nzp;
    c->g.unzcs = (c->g.nzc > 8) ? ((c->g.nzc + 7) / 8) * 8 : c->g.nzc;
    c->g.uplane = (size_t)c->g.lnx * ny * 2 * c->g.unzcs;
    c->g.hx = nx / Lx; c->g.hy = ny / Ly; c->g.hz = nz / Lz;
    c->L[0] = Lx; c->L[1] = Ly; c->L[2] = Lz;
    c->mode = mode;
    c->scheme = gamma_scheme;
    c->dim = (mode == FGB_MODE_HYPERELASTICITY) ? 9 : ((mode == FGB_MODE_HEAT || mode == FGB_MODE_POROUS) ? 3 : 6);
    c->udim = (c->dim == 3) ? 1 : 3;
    c->device = device;
    c->rank = rank; c->nranks = nranks;
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    c->own_stream = nullptr;
    c->stream = nullptr;
    for (int i = 0; i < FGB_MAX_FIELDS; i++) c->fields[i] = nullptr;
    for (int i = 0; i < FGB_MAX_PHASES; i++) { c->phi[i] = nullptr; c->laws[i].id = -1; }
    c->ubuf = nullptr; c->normals = nullptr; c->orient = nullptr;
    c->nphases = 0; c->mix = FGB_MIX_VOIGT; c->freq_hack = 0;
    const double eps = 2.220446049250313e-16;
    c->lam.eps_t = 4 * eps;
    c->lam.eps_a = pow(eps, 2.0 / 3.0);
    c->lam.eps_g = eps;
    c->lam.alpha = 0.001; c->lam.beta = 0.1;
    c->lam.delta = 1 - 1024 * eps;
    c->lam.maxiter = 32; c->lam.backtrack = 1; c->lam.project_t = 1; c->lam.fixed_c1 = -1.0;
    for (int a = 0; a < 3; a++) {
        c->tw_dev[a] = nullptr; c->kpm_dev[a] = nullptr; c->kp_dev[a] = nullptr; c->xi_dev[a] = nullptr;
        c->xi2pi_dev[a] = nullptr; c->wtan_dev[a] = nullptr; c->wex_dev[a] = nullptr; c->pois_dev[a] = nullptr;
    }
    c->d_partials = nullptr; c->d_result = nullptr; c->h_result = nullptr; c->d_scalars = nullptr; c->d_flag = nullptr; c->h_flag = nullptr;
    c->cg_dev = false; c->reduce_on_device = false; c->h_ring = nullptr;
    for (int i = 0; i < FGB_CG_RING; i++) c->ring_ev[i] = nullptr;
    c->nccl_comm = nullptr; c->nccl_lib = nullptr; c->sbuf = nullptr; c->xbuf = nullptr; c->xbuf_cap = 0;
    c->halo = nullptr; c->halo_slot = 0; c->d_gather = nullptr; c->visc_tmp = nullptr; c->p2p = false; c->iso_halo = nullptr; c->phi_halo_valid = false;
    c->heatK = nullptr; c->heatK_valid = false; c->heatK_diag = 0;
    c->nh_cache = nullptr; c->nh_cache_of = -1; c->nh_cache_mu0 = 0;
    c->dfg = 0; c->dfg1 = c->dfg2 = nullptr; c->normals_f = c->orient_f = nullptr;
    for (int i = 0; i < FGB_MAX_PHASES; i++) c->phi_f[i] = nullptr;
    c->halo_base = nullptr; c->iso_set = 0; c->halo_seq = c->iso_seq = 0;
    c->sync_base = nullptr; c->bar_seq = c->ex_seq = 0;
    for (int q = 0; q < 8; q++) c->peer_sync[q] = nullptr;
    for (int q = 0; q < 8; q++) c->peer_halo[q] = nullptr;
    c->launches = 0; c->profiling = false;
    c->bc_active = false; c->bc_relax = 1.0;
    for (int i = 0; i < 81; i++) c->bc_MQ[i] = c->bc_MQC0[i] = 0;
    for (int i = 0; i < 9; i++) c->F00[i] = 0;

#define CREATE_CUDA(call)                                                                              \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess) {                                                                      \
            int code = (e__ == cudaErrorMemoryAllocation) ? FGB_ENOMEM : FGB_ECUDA;                    \
            fgb_fail(nullptr, code, "%s failed: %s", #call, cudaGetErrorString(e__));                  \
            fgb_destroy(c);                                                                            \
            return code;                                                                               \
        }                                                                                              \
    } while (0)
    CREATE_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    c->red_blocks = c->sm_count * 8;
    CREATE_CUDA(cudaMalloc(&c->d_partials, sizeof(double) * 32 * (size_t)c->red_blocks));
    c->implicit_w_of = -1;
    CREATE_CUDA(cudaMalloc(&c->d_result, sizeof(double) * 64));
    CREATE_CUDA(cudaMallocHost(&c->h_result, sizeof(double) * 64));
    CREATE_CUDA(cudaMalloc(&c->d_scalars, sizeof(double) * 32));
    CREATE_CUDA(cudaMemset(c->d_scalars, 0, sizeof(double) * 32));
    CREATE_CUDA(cudaMallocHost(&c->h_ring, sizeof(double) * 8 * FGB_CG_RING));
    for (int i = 0; i < FGB_CG_RING; i++) CREATE_CUDA(cudaEventCreateWithFlags(&c->ring_ev[i], cudaEventDisableTiming));
    CREATE_CUDA(cudaMalloc(&c->d_flag, sizeof(int)));
    CREATE_CUDA(cudaMemset(c->d_flag, 0, sizeof(int)));
    CREATE_CUDA(cudaMallocHost(&c->h_flag, sizeof(int)));
    if (c->scheme == FGB_GAMMA_STAGGERED) CREATE_CUDA(cudaMalloc(&c->ubuf, sizeof(double) * c->g.uplane * c->udim));
    int rc = fgb_fft_init(c);
    if (rc) {
        g_create_error = c->err;
        fgb_destroy(c);
        return rc;
    }
    *out = c;
    return FGB_OK;
}

extern "C" void fgb_destroy(fgb_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    fgb_comm_free(c);
    for (int i = 0; i < FGB_MAX_FIELDS; i++) if (c->fields[i]) cudaFree(c->fields[i]);
    for (int i = 0; i < FGB_MAX_PHASES; i++) if (c->phi[i]) cudaFree(c->phi[i]);
    if (c->ubuf) cudaFree(c->ubuf);
    if (c->visc_tmp) cudaFree(c->visc_tmp);
    if (c->heatK) cudaFree(c->heatK);
    if (c->nh_cache) cudaFree(c->nh_cache);
    if (c->normals) cudaFree(c->normals);
    if (c->orient) cudaFree(c->orient);
    for (int i = 0; i < FGB_MAX_PHASES; i++) if (c->phi_f[i]) cudaFree(c->phi_f[i]);
    if (c->dfg1) cudaFree(c->dfg1);
    if (c->dfg2) cudaFree(c->dfg2);
    if (c->normals_f) cudaFree(c->normals_f);
    if (c->orient_f) cudaFree(c->orient_f);
    fgb_fft_free(c);
    if (c->d_partials) cudaFree(c->d_partials);
    if (c->d_result) cudaFree(c->d_result);
    if (c->h_result) cudaFreeHost(c->h_result);
    if (c->d_flag) cudaFree(c->d_flag);
    if (c->h_flag) cudaFreeHost(c->h_flag);
    if (c->d_scalars) cudaFree(c->d_scalars);
    if (c->h_ring) cudaFreeHost(c->h_ring);
    for (int i = 0; i < FGB_CG_RING; i++) if (c->ring_ev[i]) cudaEventDestroy(c->ring_ev[i]);
    for (auto& p : g_pending[c]) { cudaEventDestroy(p.e0); cudaEventDestroy(p.e1); }
    for (auto& e : g_event_pool[c]) cudaEventDestroy(e);
    g_pending.erase(c);
    g_event_pool.erase(c);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

extern "C" int fgb_set_stream(fgb_ctx* c, void* s) {
    if (!c) return FGB_EINVAL;
    FGB_CUDA(c, cudaStreamSynchronize(c->stream));
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return FGB_OK;
}

extern "C" int fgb_synchronize(fgb_ctx* c) {
    if (!c) return FGB_EINVAL;
    FGB_CUDA(c, cudaStreamSynchronize(c->stream));
    return FGB_OK;
}

extern "C" int fgb_local_nx(const fgb_ctx* c) { return c ? c->g.lnx : 0; }
extern "C" int fgb_local_x0(const fgb_ctx* c) { return c ? c->g.x0 : 0; }
extern "C" size_t fgb_plane_elems(const fgb_ctx* c) { return c ? c->g.plane : 0; }
extern "C" int fgb_dim(const fgb_ctx* c) { return c ? c->dim : 0; }

// ---- fields -----------------------------------------------------------------------------------------
#define CHECK_CTX(c) do { if (!(c)) return FGB_EINVAL; cudaSetDevice((c)->device); } while (0)
#define CHECK_FIELD(c, f)                                                                           \
    do {                                                                                            \
        if ((f) < 0 || (f) >= FGB_MAX_FIELDS || !(c)->fields[f])                                    \
            return fgb_fail(c, FGB_EINVAL, "invalid field id %d", (int)(f));                        \
    } while (0)

extern "C" int fgb_field_alloc(fgb_ctx* c) {
    CHECK_CTX(c);
    for (int i = 0; i < FGB_MAX_FIELDS; i++) {
        if (!c->fields[i]) {
            cudaError_t e = cudaMalloc(&c->fields[i], sizeof(double) * c->g.plane * c->dim);
            if (e != cudaSuccess) {
                c->fields[i] = nullptr;
                return fgb_fail(c, FGB_ENOMEM, "cannot allocate a %d-component field of %zu doubles per component: %s", c->dim,
                                c->g.plane, cudaGetErrorString(e));
            }
            FGB_CUDA(c, cudaMemsetAsync(c->fields[i], 0, sizeof(double) * c->g.plane * c->dim, c->stream));
            return i;
        }
    }
    return fgb_fail(c, FGB_ENOMEM, "all %d field slots in use", FGB_MAX_FIELDS);
}

extern "C" int fgb_field_free(fgb_ctx* c, int f) {
    CHECK_CTX(c);
    CHECK_FIELD(c, f);
    FGB_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->fields[f]);
    c->fields[f] = nullptr;
    return FGB_OK;
}

extern "C" int fgb_field_upload(fgb_ctx* c, int f, const double* const* comps) {
    CHECK_CTX(c);
    CHECK_FIELD(c, f);
    for (int d = 0; d < c->dim; d++)
        FGB_CUDA(c, cudaMemcpyAsync(c->fields[f] + (size_t)d * c->g.plane, comps[d], sizeof(double) * c->g.plane, cudaMemcpyHostToDevice, c->stream));
    FGB_CUDA(c, cudaStreamSynchronize(c->stream));
    return FGB_OK;
}

extern "C" int fgb_field_download(fgb_ctx* c, int f, double* const* comps) {
    CHECK_CTX(c);
    CHECK_FIELD(c, f);
    for (int d = 0; d < c->dim; d++)
        FGB_CUDA(c, cudaMemcpyAsync(comps[d], c->fields[f] + (size_t)d * c->g.plane, sizeof(double) * c->g.plane, cudaMemcpyDeviceToHost, c->stream));
    FGB_CUDA(c, cudaStreamSynchronize(c->stream));
    return FGB_OK;
}

extern "C" void* fgb_field_device_ptr(fgb_ctx* c, int f, int comp) {
    if (!c || f < 0 || f >= FGB_MAX_FIELDS || !c->fields[f] || comp < 0 || comp >= c->dim) return nullptr;
    return c->fields[f] + (size_t)comp * c->g.plane;
}

// ---- setup ------------------------------------------------------------------------------------------
extern "C" int fgb_set_num_phases(fgb_ctx* c, int n) {
    CHECK_CTX(c);
    if (n < 1 || n > FGB_MAX_PHASES) return fgb_fail(c, FGB_EINVAL, "number of phases %d out of range 1..%d", n, FGB_MAX_PHASES);
    c->nphases = n;
    return FGB_OK;
}

static int upload_planes(fgb_ctx* c, double** dst, const double* const* comps, int n, size_t plane) {
    if (!*dst) {
        cudaError_t e = cudaMalloc(dst, sizeof(double) * plane * n);
        if (e != cudaSuccess) { *dst = nullptr; return fgb_fail(c, FGB_ENOMEM, "device allocation failed: %s", cudaGetErrorString(e)); }
    }
    for (int d = 0; d < n; d++)
        FGB_CUDA(c, cudaMemcpyAsync(*dst + (size_t)d * plane, comps[d], sizeof(double) * plane, cudaMemcpyHostToDevice, c->stream));
    FGB_CUDA(c, cudaStreamSynchronize(c->stream));
    return FGB_OK;
}

// use_dfg (fg:14894-14897): the half_staggered / full_staggered schemes evaluate the material on a doubly fine grid
extern "C" int fgb_set_dfg(fgb_ctx* c, int mode) {
    CHECK_CTX(c);
    if (mode < 0 || mode > 2) return fgb_fail(c, FGB_EINVAL, "dfg mode %d (0 off, 1 half_staggered, 2 full_staggered)", mode);
    if (mode && c->scheme != FGB_GAMMA_STAGGERED) return fgb_fail(c, FGB_EINVAL, "the doubly fine grid belongs to the staggered scheme");
    if (mode && c->nranks > 1) return fgb_fail(c, FGB_EUNSUPPORTED, "half_staggered / full_staggered are single-GPU in this build");
    c->dfg = mode;
    if (!mode) return FGB_OK;
    GridDev& f = c->gf;
    f = c->g;
    f.nx = 2 * c->g.nx; f.ny = 2 * c->g.ny; f.nz = 2 * c->g.nz;
    f.lnx = 2 * c->g.lnx; f.x0 = 2 * c->g.x0;
    f.nzc = f.nz / 2 + 1;
    f.nzp = 2 * f.nzc;
    f.plane = (size_t)f.lnx * f.ny * f.nzp;
    f.hx = f.nx / c->L[0]; f.hy = f.ny / c->L[1]; f.hz = f.nz / c->L[2];
    if (!c->dfg1) {
        cudaError_t e = cudaMalloc(&c->dfg1, sizeof(double) * f.plane * c->dim);                     // _temp_dfg_1 fg:15145-15147
        if (e != cudaSuccess) { c->dfg1 = nullptr; return fgb_fail(c, FGB_ENOMEM, "cannot allocate the doubly fine grid (%zu doubles per component)", f.plane); }
    }
    return FGB_OK;
}

extern "C" int fgb_set_phase(fgb_ctx* c, int p, const double* phi) {
    CHECK_CTX(c);
    if (p < 0 || p >= c->nphases) return fgb_fail(c, FGB_EINVAL, "phase index %d out of range", p);
    const double* comps[1] = {phi};
    c->phi_halo_valid = false;
    c->heatK_valid = false;
    if (c->dfg == 2) return upload_planes(c, &c->phi_f[p], comps, 1, c->gf.plane);                    // full_staggered: phases live on the fine grid (fg:17154-17156)
    int rc = upload_planes(c, &c->phi[p], comps, 1, c->g.plane);
    if (rc || c->dfg != 1) return rc;
    // half_staggered: the fine-grid phase is the piecewise constant continuation of the coarse one (initFullStageredRawPhases fg:17648)
    if (!c->phi_f[p]) {
        cudaError_t e = cudaMalloc(&c->phi_f[p], sizeof(double) * c->gf.plane);
        if (e != cudaSuccess) { c->phi_f[p] = nullptr; return fgb_fail(c, FGB_ENOMEM, "device allocation failed: %s", cudaGetErrorString(e)); }
        FGB_CUDA(c, cudaMemsetAsync(c->phi_f[p], 0, sizeof(double) * c->gf.plane, c->stream));
    }
    return fgb_k_inject_phase(c, c->phi[p], c->phi_f[p]);
}

// prolongate_to_dfg fg:14216 / restrict_from_dfg fg:14273 between a field and _temp_dfg_1 (reference self-test fg:24491-24515)
extern "C" int fgb_dfg_prolongate(fgb_ctx* c, int f) {
    CHECK_CTX(c); CHECK_FIELD(c, f);
    if (!c->dfg) return fgb_fail(c, FGB_EINVAL, "the doubly fine grid is not enabled (fgb_set_dfg)");
    return fgb_k_prolongate(c, c->fields[f], c->dfg1);
}
extern "C" int fgb_dfg_restrict(fgb_ctx* c, int f) {
    CHECK_CTX(c); CHECK_FIELD(c, f);
    if (!c->dfg) return fgb_fail(c, FGB_EINVAL, "the doubly fine grid is not enabled (fgb_set_dfg)");
    return fgb_k_restrict(c, c->dfg1, c->fields[f]);
}

extern "C" int fgb_set_law(fgb_ctx* c, int p, int law_id, const double* params, int nparams) {
    CHECK_CTX(c);
    if (p < 0 || p >= c->nphases) return fgb_fail(c, FGB_EINVAL, "phase index %d out of range", p);
    static const int need[8] = {2, 36, 5, 1, 6, 2, 2, 2};
    static const int ldim[8] = {6, 6, 6, 0, 3, 9, 9, 9};
    if (law_id < 0 || law_id > FGB_LAW_NH2) return fgb_fail(c, FGB_EUNSUPPORTED, "unknown material law id %d (fg:15285)", law_id);
    if (nparams != need[law_id] && !(law_id == FGB_LAW_TISO && nparams == 8))
        return fgb_fail(c, FGB_EINVAL, "law %d needs %d parameters, got %d", law_id, need[law_id], nparams);
    if (ldim[law_id] && ldim[law_id] != c->dim)
        return fgb_fail(c, FGB_EINVAL, "law %d acts on %d components but mode has %d (fg:15211-15294)", law_id, ldim[law_id], c->dim);
    if (law_id == FGB_LAW_SCALAR && c->dim == 9) return fgb_fail(c, FGB_EINVAL, "scalar law is not defined for hyperelasticity");
    c->heatK_valid = false;
    c->laws[p].id = law_id;
    for (int i = 0; i < FGB_MAX_LAW_PARAMS; i++) c->laws[p].p[i] = i < nparams ? params[i] : 0.0;
    return FGB_OK;
}

// with the doubly fine grid the normals / orientation live on the fine grid (get_normals / get_orientation fg:14911-14937)
extern "C" int fgb_set_normals(fgb_ctx* c, const double* const* comps3) {
    CHECK_CTX(c);
    c->heatK_valid = false;
    if (c->dfg) return upload_planes(c, &c->normals_f, comps3, 3, c->gf.plane);
    return upload_planes(c, &c->normals, comps3, 3, c->g.plane);
}
extern "C" int fgb_set_orientation(fgb_ctx* c, const double* const* comps3) {
    CHECK_CTX(c);
    if (c->dfg) return upload_planes(c, &c->orient_f, comps3, 3, c->gf.plane);
    return upload_planes(c, &c->orient, comps3, 3, c->g.plane);
}

extern "C" int fgb_set_mixing(fgb_ctx* c, int rule, const double* lp, int n) {
    CHECK_CTX(c);
    if (rule != FGB_MIX_VOIGT && rule != FGB_MIX_LAMINATE && rule != FGB_MIX_REUSS)
        return fgb_fail(c, FGB_EUNSUPPORTED, "Unknown material mixing rule %d (voigt, reuss and laminate only)", rule);
    c->mix = rule;
    c->heatK_valid = false;
    if (lp) {
        if (n != 10) return fgb_fail(c, FGB_EINVAL, "laminate parameter vector must have 10 entries");
        c->lam.eps_t = lp[0]; c->lam.eps_a = lp[1]; c->lam.eps_g = lp[2]; c->lam.alpha = lp[3]; c->lam.beta = lp[4];
        c->lam.delta = lp[5]; c->lam.maxiter = (int)lp[6]; c->lam.backtrack = lp[7] != 0; c->lam.project_t = lp[8] != 0;
        c->lam.fixed_c1 = lp[9];
    }
    return FGB_OK;
}

extern "C" int fgb_set_freq_hack(fgb_ctx* c, int on) {
    CHECK_CTX(c);
    c->freq_hack = on ? 1 : 0;
    return FGB_OK;
}

extern "C" int fgb_set_bc(fgb_ctx* c, const double* MQ, const double* M_QC0, double bc_relax) {
    CHECK_CTX(c);
    const int d = c->dim;
    c->bc_active = false;
    c->bc_relax = bc_relax;
    double nrm = 0;
    for (int i = 0; i < d * d; i++) {
        c->bc_MQ[i] = MQ ? MQ[i] : 0.0;
        c->bc_MQC0[i] = M_QC0 ? M_QC0[i] : 0.0;
        nrm += c->bc_MQ[i] * c->bc_MQ[i];
    }
    // initBCProjector skips the mean when ||MQ||_F < eps (fg:20233)
    c->bc_active = std::sqrt(nrm) >= 2.220446049250313e-16;
    return FGB_OK;
}

// ---- BLAS-1 / reductions ------------------------------------------------------------------------------
extern "C" int fgb_set_constant(fgb_ctx* c, int f, const double* v) { CHECK_CTX(c); CHECK_FIELD(c, f); return fgb_k_set_constant(c, c->fields[f], v, 0); }
extern "C" int fgb_add_constant(fgb_ctx* c, int f, const double* v) { CHECK_CTX(c); CHECK_FIELD(c, f); return fgb_k_set_constant(c, c->fields[f], v, 1); }
extern "C" int fgb_copy(fgb_ctx* c, int s, int d) { CHECK_CTX(c); CHECK_FIELD(c, s); CHECK_FIELD(c, d); return fgb_k_copy(c, c->fields[s], c->fields[d], c->dim); }
extern "C" int fgb_xpay(fgb_ctx* c, int r, int x, double a, int y) {
    CHECK_CTX(c); CHECK_FIELD(c, r); CHECK_FIELD(c, x); CHECK_FIELD(c, y);
    return fgb_k_xpay(c, c->fields[r], c->fields[x], a, c->fields[y]);
}
extern "C" int fgb_xpaymz(fgb_ctx* c, int r, int x, double a, int y, int z) {
    CHECK_CTX(c); CHECK_FIELD(c, r); CHECK_FIELD(c, x); CHECK_FIELD(c, y); CHECK_FIELD(c, z);
    return fgb_k_xpaymz(c, c->fields[r], c->fields[x], a, c->fields[y], c->fields[z]);
}
extern "C" int fgb_adjust_residual(fgb_ctx* c, int r, const double* E, int z) {
    CHECK_CTX(c); CHECK_FIELD(c, r); CHECK_FIELD(c, z);
    return fgb_k_adjust_residual(c, c->fields[r], E, c->fields[z]);
}
extern "C" int fgb_extrapolate_polynomial(fgb_ctx* c, int n, const int* fields, const double* Vinv, const double* tpowers, int dst) {
    CHECK_CTX(c); CHECK_FIELD(c, dst);
    if (n < 1 || n > 8 || !fields || !Vinv || !tpowers) return fgb_fail(c, FGB_EINVAL, "polynomial extrapolation needs 1..8 fields");
    const double* f[8];
    for (int i = 0; i < n; i++) { CHECK_FIELD(c, fields[i]); f[i] = c->fields[fields[i]]; }
    return fgb_k_extrapolate_poly(c, n, f, Vinv, tpowers, c->fields[dst]);
}
extern "C" int fgb_inner(fgb_ctx* c, int a, int b, int cc, double* out) {
    CHECK_CTX(c); CHECK_FIELD(c, a); CHECK_FIELD(c, b);
    if (cc >= 0) CHECK_FIELD(c, cc);
    return fgb_k_inner(c, c->fields[a], c->fields[b], cc >= 0 ? c->fields[cc] : nullptr, out);
}
extern "C" int fgb_average(fgb_ctx* c, int f, double* out) { CHECK_CTX(c); CHECK_FIELD(c, f); return fgb_k_component_dot(c, c->fields[f], nullptr, out, 1); }
extern "C" int fgb_component_dot(fgb_ctx* c, int a, int b, double* out) {
    CHECK_CTX(c); CHECK_FIELD(c, a); CHECK_FIELD(c, b);
    return fgb_k_component_dot(c, c->fields[a], c->fields[b], out, 0);
}

static int poll_flag(fgb_ctx* c) {
    FGB_CUDA(c, cudaMemcpyAsync(c->h_flag, c->d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    FGB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (*c->h_flag & 4) {
        FGB_CUDA(c, cudaMemsetAsync(c->d_flag, 0, sizeof(int), c->stream));
        return fgb_fail(c, FGB_ECOMM, "peer synchronisation timed out (a rank of the slab partition did not reach the barrier within 4 s)");
    }
    if (*c->h_flag) {
        FGB_CUDA(c, cudaMemsetAsync(c->d_flag, 0, sizeof(int), c->stream));
        return fgb_fail(c, FGB_ENUMERIC, "material law domain error on the device (log/pow of a non-positive det F, or >2 phases in a laminate voxel)");
    }
    return FGB_OK;
}

extern "C" int fgb_mean_pk1(fgb_ctx* c, int f, double alpha, double* out) {
    CHECK_CTX(c); CHECK_FIELD(c, f);
    int rc = fgb_k_mean_pk1(c, c->fields[f], alpha, out);
    return rc ? rc : poll_flag(c);
}
extern "C" int fgb_mean_energy(fgb_ctx* c, int f, double* out) {
    CHECK_CTX(c); CHECK_FIELD(c, f);
    if (c->mix == FGB_MIX_REUSS) return fgb_fail(c, FGB_EUNSUPPORTED, "Reuss MaterialLaw energy not implemented");   // fg:12660
    int rc = fgb_k_mean_energy(c, c->fields[f], out);
    return rc ? rc : poll_flag(c);
}
extern "C" int fgb_mean_cauchy(fgb_ctx* c, int f, double alpha, double* out) {
    CHECK_CTX(c); CHECK_FIELD(c, f);
    int rc = fgb_k_mean_cauchy(c, c->fields[f], alpha, out);
    return rc ? rc : poll_flag(c);
}
extern "C" int fgb_min_detF(fgb_ctx* c, int f, double* out) { CHECK_CTX(c); CHECK_FIELD(c, f); return fgb_k_min_detF(c, c->fields[f], out); }
extern "C" int fgb_ref_material(fgb_ctx* c, int f, int zt, double* lmin, double* lmax) {
    CHECK_CTX(c); CHECK_FIELD(c, f);
    int rc = fgb_k_ref_material(c, c->fields[f], zt, lmin, lmax);
    return rc ? rc : poll_flag(c);
}

// ---- constitutive sweeps ---------------------------------------------------------------------------------
extern "C" int fgb_calc_stress(fgb_ctx* c, int s, int d, double mu0, double lambda0, double alpha) {
    CHECK_CTX(c); CHECK_FIELD(c, s); CHECK_FIELD(c, d);
    return fgb_k_calc_stress(c, c->fields[s], c->fields[d], mu0, lambda0, alpha);
}
extern "C" int fgb_calc_stress_deriv(fgb_ctx* c, int F, int W, int d, double mu0, double lambda0, double alpha) {
    CHECK_CTX(c); CHECK_FIELD(c, F); CHECK_FIELD(c, W); CHECK_FIELD(c, d);
    return fgb_k_calc_stress_deriv(c, c->fields[F], c->fields[W], c->fields[d], mu0, lambda0, alpha);
}
extern "C" int fgb_calc_stress_const(fgb_ctx* c, int s, int d, double mu0, double lambda0) {
    CHECK_CTX(c); CHECK_FIELD(c, s); CHECK_FIELD(c, d);
    return fgb_k_calc_stress_const(c, c->fields[s], c->fields[d], mu0, lambda0);
}
extern "C" int fgb_calc_polarization(fgb_ctx* c, int s, int d, double mu0, int inv) {
    CHECK_CTX(c); CHECK_FIELD(c, s); CHECK_FIELD(c, d);
    return fgb_k_calc_polarization(c, c->fields[s], c->fields[d], mu0, inv);
}

// ---- Green operator -----------------------------------------------------------------------------------------
// Voigt product M:v (fg:563-575): entries 3..5 of v doubled for dim 6
static void dyad4_mv(int d, const double* M, const double* v, double* out) {
    for (int r = 0; r < d; r++) {
        double s = 0;
        for (int k = 0; k < d; k++) s += M[r * d + k] * v[k] * ((d == 6 && k >= 3) ? 2.0 : 1.0);
        out[r] = s;
    }
}

// R = bc_relax*MQ:F0 - (1-bc_relax)*M:(QC0:F00)   (calcBCProjector fg:20258)
static int bc_term(fgb_ctx* c, const double* tau, double* R) {
    const int d = c->dim;
    for (int i = 0; i < d; i++) R[i] = 0;
    if (!c->bc_active && c->bc_relax == 1.0) return FGB_OK;
    double F0[9] = {0}, a[9], b[9];
    if (c->bc_active) {
        int rc = fgb_k_component_dot(c, tau, nullptr, F0, 1);       // initBCProjector fg:20220/20228
        if (rc) return rc;
    }
    dyad4_mv(d, c->bc_MQ, F0, a);
    dyad4_mv(d, c->bc_MQC0, c->F00, b);
    for (int i = 0; i < d; i++) R[i] = c->bc_relax * a[i] - (1 - c->bc_relax) * b[i];
    return FGB_OK;
}

static int green_args(fgb_ctx* c, GreenArgs& ga, double mu0, double lambda0, double alpha, double beta, bool staggered) {
    memset(&ga, 0, sizeof(ga));
    ga.beta = beta;
    ga.freq_hack = c->freq_hack;
    if (staggered) {
        if (c->dim == 3) { ga.kind = 2; ga.c10 = -alpha / (2 * mu0); }                                              // fg:19758-19763
        else if (c->dim == 6) { ga.kind = 1; ga.c10 = -alpha / mu0; ga.c20 = -alpha / (mu0 * (1 + mu0 / (lambda0 + mu0))); }   // fg:19749-19755
        else { ga.kind = 1; ga.c10 = -alpha / (2 * mu0); ga.c20 = -alpha / (2 * mu0 * (1 + 2 * mu0 / lambda0)); }  // fg:19768-19774
    } else if (c->scheme == FGB_GAMMA_WILLOT) {
        ga.kind = 8; ga.c10 = mu0; ga.c20 = mu0 / lambda0; ga.alpha = alpha;                                        // fg:19091
    } else {
        if (c->dim == 3) { ga.kind = 4; ga.c10 = alpha / (2 * mu0); }                                               // fg:19309
        else if (c->dim == 6) { ga.kind = 3; ga.c10 = alpha / (4 * mu0); ga.c20 = -alpha / (mu0 * (1 + mu0 / (lambda0 + mu0))); }  // fg:19387-19388
        else { ga.kind = 5; ga.c10 = alpha / (2 * mu0); ga.c20 = -alpha / (2 * mu0 * (1 + 2 * mu0 / lambda0)); }   // fg:19626-19627
    }
    return FGB_OK;
}

// the displacement buffer exists from the start in staggered contexts; collocated contexts get it on first use of a
// staggered-grid operator (get_raw_field('u') always uses them, fg:15517-15557)
static int ensure_ubuf(fgb_ctx* c) {
    if (c->ubuf) return FGB_OK;
    FGB_CUDA(c, cudaMalloc(&c->ubuf, sizeof(double) * c->g.uplane * c->udim));
    return FGB_OK;
}

// G0OperatorStaggered* (fg:20101-20153) on the u buffer
static int g0_staggered(fgb_ctx* c, double mu0, double lambda0, double alpha) {
    GreenArgs ga;
    green_args(c, ga, mu0, lambda0, alpha, 0.0, true);
    int rc;
    const FftLayout lay = {c->g.unzcs};
    c->implicit_w_of = -1;
    if ((rc = fgb_fft_z_forward(c, c->ubuf, c->udim, lay))) return rc;
    if (c->nranks > 1) {
        // slab partition: the y passes are part of the transposed x pass (they read/write the all-to-all staging layout)
        if ((rc = fgb_comm_fft_x(c, c->ubuf, c->udim, lay, &ga))) return rc;
    } else {
        if ((rc = fgb_fft_y(c, c->ubuf, c->udim, lay, -1))) return rc;
        if ((rc = fgb_fft_x(c, c->ubuf, c->udim, lay, 0, &ga))) return rc;
        if ((rc = fgb_fft_y(c, c->ubuf, c->udim, lay, +1))) return rc;
    }
    return fgb_fft_z_backward(c, c->ubuf, c->udim, lay);
}

static int gamma_impl(fgb_ctx* c, double* field, const double* E, double mu0, double lambda0, double alpha, double beta) {
    double R[9];
    int rc = bc_term(c, field, R);
    if (rc) return rc;
    double Ec[9];
    for (int i = 0; i < c->dim; i++) Ec[i] = E[i] + alpha * R[i];                  // applyBCProjector fg:20263-20279
    if (c->scheme == FGB_GAMMA_STAGGERED) {                                         // GammaOperatorStaggered* fg:20288-20378
        if (c->nranks > 1 && (rc = fgb_comm_halo_tau(c, field))) return rc;
        if ((rc = fgb_k_div(c, field, c->ubuf))) return rc;
        if ((rc = g0_staggered(c, mu0, lambda0, alpha))) return rc;
        if (c->nranks > 1 && (rc = fgb_comm_halo_u(c))) return rc;
        return fgb_k_eps(c, c->ubuf, field, Ec);
    }
    GreenArgs ga;                                                                   // GammaOperatorCollocated* fg:20302-20340
    green_args(c, ga, mu0, lambda0, alpha, beta, false);
    for (int i = 0; i < c->dim; i++) ga.dc[i] = Ec[i];
    const FftLayout lay = {c->g.nzc};
    if ((rc = fgb_fft_z_forward(c, field, c->dim, lay))) return rc;
    if (c->nranks > 1) {
        if ((rc = fgb_comm_fft_x(c, field, c->dim, lay, &ga))) return rc;
    } else {
        if ((rc = fgb_fft_y(c, field, c->dim, lay, -1))) return rc;
        if ((rc = fgb_fft_x(c, field, c->dim, lay, 0, &ga))) return rc;
        if ((rc = fgb_fft_y(c, field, c->dim, lay, +1))) return rc;
    }
    return fgb_fft_z_backward(c, field, c->dim, lay);
}

// DeltaOperatorStaggered (fg:20422-20460): viscosity dual formulation.  `copy` holds tau (input), field is overwritten.
static int delta_collocated(fgb_ctx* c, double* field, const double* E, double mu0, double alpha);
static int delta_impl(fgb_ctx* c, double* field, const double* tau_copy, const double* E, double mu0, double alpha) {
    if (c->scheme == FGB_GAMMA_COLLOCATED) return delta_collocated(c, field, E, mu0, alpha);
    // DeltaOperatorStaggered fg:20422-20460 and DeltaOperatorWillotR fg:20380-20418 share this form
    const double m = 1 / (4 * mu0);
    double mean[9], adj[9];
    int rc = fgb_k_component_dot(c, tau_copy, nullptr, mean, 1);
    if (rc) return rc;
    for (int i = 0; i < 6; i++) adj[i] = E[i] - 2 * alpha * m * mean[i];
    if ((rc = gamma_impl(c, field, adj, -1.0 / (4 * m), INFINITY, alpha, 0.0))) return rc;
    return fgb_k_xpay(c, field, field, 2 * alpha * m, tau_copy);
}

// DeltaOperatorCollocated fg:20462-20471: fftTensor(zero_trace) -> applyDeltaFourier fg:19075 -> fftInvTensor(zero_trace).
// Component 0 is never transformed: tau^_0 := -(tau^_1 + tau^_2) in Fourier space (inside the fused x pass, kind 9) and
// eta_0 := -(eta_1 + eta_2) in real space afterwards (mxpyTensor fg:20590).
static int delta_collocated(fgb_ctx* c, double* field, const double* E, double mu0, double alpha) {
    const double m = 1 / (4 * mu0);
    int rc;
    double R[9] = {0};
    if (c->bc_active || c->bc_relax != 1.0) {
        // initBCProjector(tau_hat) fg:20220 reads the zero frequency, whose component 0 is -(tau^_1 + tau^_2)
        double F0[9] = {0}, a[9], b[9];
        if (c->bc_active) {
            if ((rc = fgb_k_component_dot(c, field, nullptr, F0, 1))) return rc;
            F0[0] = -(F0[1] + F0[2]);
        }
        dyad4_mv(6, c->bc_MQ, F0, a);
        dyad4_mv(6, c->bc_MQC0, c->F00, b);
        for (int i = 0; i < 6; i++) R[i] = c->bc_relax * a[i] - (1 - c->bc_relax) * b[i];
    }
    GreenArgs ga;
    green_args(c, ga, -1.0 / (4 * m), INFINITY, alpha, 2 * alpha * m, false);      // fg:19078
    ga.kind = 9;
    for (int i = 0; i < 6; i++) ga.dc[i] = E[i] + alpha * R[i];
    const FftLayout lay = {c->g.nzc};
    double* f1 = field + c->g.plane;
    if ((rc = fgb_fft_z_forward(c, f1, 5, lay))) return rc;
    if (c->nranks > 1) {
        if ((rc = fgb_comm_fft_x(c, field, 6, lay, &ga))) return rc;               // (the y passes of component 0 act on scratch data)
    } else {
        if ((rc = fgb_fft_y(c, f1, 5, lay, -1))) return rc;
        if ((rc = fgb_fft_x(c, field, 6, lay, 0, &ga))) return rc;
        if ((rc = fgb_fft_y(c, f1, 5, lay, +1))) return rc;
    }
    if ((rc = fgb_fft_z_backward(c, f1, 5, lay))) return rc;
    return fgb_k_mxpy(c, field, field + c->g.plane, field + 2 * c->g.plane);
}

// Fourier-space operator on all `dim` components of a field in the reference layout: forward transform, operator `kind` fused into
// the x pass, inverse transform
static int fourier_operator(fgb_ctx* c, double* field, const GreenArgs& ga) {
    int rc;
    const FftLayout lay = {c->g.nzc};
    if ((rc = fgb_fft_z_forward(c, field, c->dim, lay))) return rc;
    if (c->nranks > 1) {
        if ((rc = fgb_comm_fft_x(c, field, c->dim, lay, &ga))) return rc;
    } else {
        if ((rc = fgb_fft_y(c, field, c->dim, lay, -1))) return rc;
        if ((rc = fgb_fft_x(c, field, c->dim, lay, 0, &ga))) return rc;
        if ((rc = fgb_fft_y(c, field, c->dim, lay, +1))) return rc;
    }
    return fgb_fft_z_backward(c, field, c->dim, lay);
}

// G0DivOperatorHyper fg:20281-20286 (fftTensor, G0DivOperatorFourierHyper fg:20155, fftInvVector): components 0..2 of the field
// become u = alpha * G0 Div tau (collocated Fourier discretisation, xi = 2 pi m / L); components 3..8 are left undefined
extern "C" int fgb_g0div_hyper(fgb_ctx* c, int f, double mu0, double lambda0, double alpha) {
    CHECK_CTX(c); CHECK_FIELD(c, f);
    if (c->dim != 9) return fgb_fail(c, FGB_EINVAL, "fgb_g0div_hyper needs the 9-component hyperelasticity layout");
    GreenArgs ga;
    memset(&ga, 0, sizeof(ga));
    ga.kind = 6;
    ga.c10 = -alpha / (2 * mu0);                                   // fg:20162-20163
    ga.c20 = alpha / (2 * mu0 * (1 + 2 * mu0 / lambda0));
    return fourier_operator(c, c->fields[f], ga);
}

// fftTensor, G0DivOperatorFourierHyper, GradOperatorFourierHyper, fftInvTensor (fg:24572-24575): grad G0 Div applied in Fourier space
extern "C" int fgb_grad_g0div_hyper(fgb_ctx* c, int f, double mu0, double lambda0, double alpha) {
    CHECK_CTX(c); CHECK_FIELD(c, f);
    if (c->dim != 9) return fgb_fail(c, FGB_EINVAL, "fgb_grad_g0div_hyper needs the 9-component hyperelasticity layout");
    GreenArgs ga;
    memset(&ga, 0, sizeof(ga));
    ga.kind = 11;
    ga.c10 = -alpha / (2 * mu0);
    ga.c20 = alpha / (2 * mu0 * (1 + 2 * mu0 / lambda0));
    return fourier_operator(c, c->fields[f], ga);
}

// fftTensor, GradOperatorFourierHyper fg:22069-22116, fftInvTensor (the sequence of fg:24531-24533): components 0..2 of the field
// hold a vector field q on entry, all 9 components hold grad q on return
extern "C" int fgb_grad_hyper(fgb_ctx* c, int f) {
    CHECK_CTX(c); CHECK_FIELD(c, f);
    if (c->dim != 9) return fgb_fail(c, FGB_EINVAL, "fgb_grad_hyper needs the 9-component hyperelasticity layout");
    GreenArgs ga;
    memset(&ga, 0, sizeof(ga));
    ga.kind = 7;
    return fourier_operator(c, c->fields[f], ga);
}

extern "C" int fgb_gamma(fgb_ctx* c, int f, const double* E, double mu0, double lambda0, double alpha, double beta) {
    CHECK_CTX(c); CHECK_FIELD(c, f);
    if (c->mode == FGB_MODE_VISCOSITY) return fgb_fail(c, FGB_EUNSUPPORTED, "use fgb_basic_step / fgb_cg_apply in viscosity mode");
    return gamma_impl(c, c->fields[f], E, mu0, lambda0, alpha, beta);
}

extern "C" int fgb_div_staggered(fgb_ctx* c, int f) {
    CHECK_CTX(c); CHECK_FIELD(c, f);
    if (int rcu = ensure_ubuf(c)) return rcu;
    int rc;
    if (c->nranks > 1 && (rc = fgb_comm_halo_tau(c, c->fields[f]))) return rc;
    return fgb_k_div(c, c->fields[f], c->ubuf);
}
extern "C" int fgb_g0_staggered(fgb_ctx* c, double mu0, double lambda0, double alpha) {
    CHECK_CTX(c);
    if (int rcu = ensure_ubuf(c)) return rcu;
    return g0_staggered(c, mu0, lambda0, alpha);
}
// get_raw_field('u') fg:15517-15557: u = G0 div_h tau(eps) into the u buffer (read it with fgb_u_download)
extern "C" int fgb_calc_displacement(fgb_ctx* c, int eps, int tmp, double mu0, double lambda0) {
    CHECK_CTX(c); CHECK_FIELD(c, eps); CHECK_FIELD(c, tmp);
    if (eps == tmp) return fgb_fail(c, FGB_EINVAL, "fgb_calc_displacement needs a scratch field different from the strain field");
    int rc = ensure_ubuf(c);
    if (rc) return rc;
    double m = mu0, l = lambda0, a = 1.0;
    if (c->mode == FGB_MODE_VISCOSITY) {
        if ((rc = fgb_k_calc_stress(c, c->fields[eps], c->fields[tmp], mu0, lambda0, 1.0))) return rc;      // calcStressDiff fg:18030
        m = 1 / (4 * mu0); l = INFINITY; a = 1 / (2 * mu0);                                                // fg:15535
    } else if (c->mode == FGB_MODE_HYPERELASTICITY) {
        // fg:15524-15527: calcStressDiff, then G0DivOperatorHyper (collocated Fourier div and G0, not the staggered operators)
        if ((rc = fgb_k_calc_stress(c, c->fields[eps], c->fields[tmp], mu0, lambda0, 1.0))) return rc;
        if ((rc = poll_flag(c))) return rc;
        if ((rc = fgb_g0div_hyper(c, tmp, mu0, lambda0, 1.0))) return rc;
        c->implicit_w_of = -1;
        for (int d = 0; d < 3; d++)
            FGB_CUDA(c, cudaMemcpy2DAsync(c->ubuf + (size_t)d * c->g.uplane, sizeof(double) * 2 * c->g.unzcs, c->fields[tmp] + (size_t)d * c->g.plane,
                                          sizeof(double) * c->g.nzp, sizeof(double) * c->g.nzp, (size_t)c->g.lnx * c->g.ny, cudaMemcpyDeviceToDevice, c->stream));
        return FGB_OK;
    } else {
        if ((rc = fgb_k_calc_stress_const(c, c->fields[eps], c->fields[tmp], mu0, lambda0))) return rc;    // fg:17973
    }
    if ((rc = poll_flag(c))) return rc;
    if (c->nranks > 1 && (rc = fgb_comm_halo_tau(c, c->fields[tmp]))) return rc;
    if ((rc = fgb_k_div(c, c->fields[tmp], c->ubuf))) return rc;
    return g0_staggered(c, m, l, a);
}
// get_raw_field("p") fg:15559-15573: pressure of the viscosity formulation.  calcStressDiff, divOperatorStaggered, divVector with
// alpha = 1/(2 mu0) (fg:19983), poisson_solve (fg:23454).  tmp is a scratch field; the result is left in component 0 of the u buffer
// (fgb_u_download with ncomp = 1).
extern "C" int fgb_calc_pressure(fgb_ctx* c, int eps, int tmp, double mu0, double lambda0) {
    CHECK_CTX(c); CHECK_FIELD(c, eps); CHECK_FIELD(c, tmp);
    if (eps == tmp) return fgb_fail(c, FGB_EINVAL, "fgb_calc_pressure needs a scratch field different from the strain field");
    if (c->dim < 6) return fgb_fail(c, FGB_EINVAL, "field 'p' needs a mode with at least 6 tensor components (fg:15564-15565)");
    int rc = ensure_ubuf(c);
    if (rc) return rc;
    if ((rc = fgb_k_calc_stress(c, c->fields[eps], c->fields[tmp], mu0, lambda0, 1.0))) return rc;
    if ((rc = poll_flag(c))) return rc;
    if (c->nranks > 1 && (rc = fgb_comm_halo_tau(c, c->fields[tmp]))) return rc;
    if ((rc = fgb_k_div(c, c->fields[tmp], c->ubuf))) return rc;
    if (c->nranks > 1 && (rc = fgb_comm_halo_u(c))) return rc;
    double* b = c->fields[tmp];                                    // one component in the u layout fits into the scratch field
    if ((rc = fgb_k_div_vector(c, c->ubuf, b, 1 / (2 * mu0)))) return rc;
    GreenArgs ga;
    memset(&ga, 0, sizeof(ga));
    ga.kind = 10;
    const FftLayout lay = {c->g.unzcs};
    if ((rc = fgb_fft_z_forward(c, b, 1, lay))) return rc;
    if (c->nranks > 1) {
        if ((rc = fgb_comm_fft_x(c, b, 1, lay, &ga))) return rc;
    } else {
        if ((rc = fgb_fft_y(c, b, 1, lay, -1))) return rc;
        if ((rc = fgb_fft_x(c, b, 1, lay, 0, &ga))) return rc;
        if ((rc = fgb_fft_y(c, b, 1, lay, +1))) return rc;
    }
    if ((rc = fgb_fft_z_backward(c, b, 1, lay))) return rc;
    c->implicit_w_of = -1;
    FGB_CUDA(c, cudaMemcpyAsync(c->ubuf, b, sizeof(double) * c->g.uplane, cudaMemcpyDeviceToDevice, c->stream));
    return FGB_OK;
}
extern "C" int fgb_eps_staggered(fgb_ctx* c, int f, const double* E) {
    CHECK_CTX(c); CHECK_FIELD(c, f);
    if (int rcu = ensure_ubuf(c)) return rcu;
    int rc;
    if (c->nranks > 1 && (rc = fgb_comm_halo_u(c))) return rc;
    return fgb_k_eps(c, c->ubuf, c->fields[f], E);
}
extern "C" int fgb_u_upload(fgb_ctx* c, const double* const* comps, int n) {
    CHECK_CTX(c);
    if (int rcu = ensure_ubuf(c)) return rcu;
    if (n != c->udim) return fgb_fail(c, FGB_EINVAL, "u buffer has %d components", c->udim);
    c->implicit_w_of = -1;
    for (int d = 0; d < n; d++)
        FGB_CUDA(c, cudaMemcpy2DAsync(c->ubuf + (size_t)d * c->g.uplane, sizeof(double) * 2 * c->g.unzcs, comps[d], sizeof(double) * c->g.nzp,
                                      sizeof(double) * c->g.nzp, (size_t)c->g.lnx * c->g.ny, cudaMemcpyHostToDevice, c->stream));
    FGB_CUDA(c, cudaStreamSynchronize(c->stream));
    return FGB_OK;
}
extern "C" int fgb_u_download(fgb_ctx* c, double* const* comps, int n) {
    CHECK_CTX(c);
    if (int rcu = ensure_ubuf(c)) return rcu;
    if (n < 1 || n > c->udim) return fgb_fail(c, FGB_EINVAL, "u buffer has %d components", c->udim);
    for (int d = 0; d < n; d++)
        FGB_CUDA(c, cudaMemcpy2DAsync(comps[d], sizeof(double) * c->g.nzp, c->ubuf + (size_t)d * c->g.uplane, sizeof(double) * 2 * c->g.unzcs,
                                      sizeof(double) * c->g.nzp, (size_t)c->g.lnx * c->g.ny, cudaMemcpyDeviceToHost, c->stream));
    FGB_CUDA(c, cudaStreamSynchronize(c->stream));
    return FGB_OK;
}

extern "C" int fgb_fft_forward(fgb_ctx* c, int f) {
    CHECK_CTX(c); CHECK_FIELD(c, f);
    int rc;
    const FftLayout lay = {c->g.nzc};
    if ((rc = fgb_fft_z_forward(c, c->fields[f], c->dim, lay))) return rc;
    if ((rc = fgb_fft_y(c, c->fields[f], c->dim, lay, -1))) return rc;
    if (c->nranks > 1) return fgb_fail(c, FGB_EUNSUPPORTED, "plain 3-D transform is single-GPU only (the slab path always fuses the Green operator)");
    return fgb_fft_x(c, c->fields[f], c->dim, lay, -1, nullptr);
}
extern "C" int fgb_fft_backward(fgb_ctx* c, int f) {
    CHECK_CTX(c); CHECK_FIELD(c, f);
    int rc;
    if (c->nranks > 1) return fgb_fail(c, FGB_EUNSUPPORTED, "plain 3-D transform is single-GPU only (the slab path always fuses the Green operator)");
    const FftLayout lay = {c->g.nzc};
    if ((rc = fgb_fft_x(c, c->fields[f], c->dim, lay, +1, nullptr))) return rc;
    if ((rc = fgb_fft_y(c, c->fields[f], c->dim, lay, +1))) return rc;
    return fgb_fft_z_backward(c, c->fields[f], c->dim, lay);
}

// ---- scheme-level ----------------------------------------------------------------------------------------------
static int fused_kind(fgb_ctx* c, double lambda0);
static int scratch_field(fgb_ctx* c, double** out) {
    if (!c->visc_tmp) {
        cudaError_t e = cudaMalloc(&c->visc_tmp, sizeof(double) * c->g.plane * c->dim);
        if (e != cudaSuccess) { c->visc_tmp = nullptr; return fgb_fail(c, FGB_ENOMEM, "cannot allocate the viscosity scratch field"); }
    }
    *out = c->visc_tmp;
    return FGB_OK;
}

// basicScheme fg:20558-20577
extern "C" int fgb_basic_step(fgb_ctx* c, int src, int dst, const double* E, double mu0, double lambda0) {
    CHECK_CTX(c); CHECK_FIELD(c, src); CHECK_FIELD(c, dst);
    int rc;
    if (c->bc_relax != 1.0 && (rc = fgb_k_component_dot(c, c->fields[src], nullptr, c->F00, 1))) return rc;   // fg:20563-20565
    if (fused_kind(c, lambda0) == 2) {
        // heat / porous: (K - 2 mu0) grad and div_h in one sweep (calcStressDiff fg:18030 + divOperatorStaggeredHeat fg:18914)
        if (c->nranks > 1 && (rc = fgb_comm_halo_heat(c, nullptr, c->fields[src]))) return rc;
        if ((rc = fgb_k_heat_march(c, nullptr, 0.0, c->fields[src], nullptr, mu0, 1.0))) return rc;
        if ((rc = g0_staggered(c, mu0, lambda0, -1.0))) return rc;
        if (c->nranks > 1 && (rc = fgb_comm_halo_u(c))) return rc;
        return fgb_k_eps(c, c->ubuf, c->fields[dst], E);
    }
    if (fgb_fused_iso_applicable(c) && !c->bc_active && c->bc_relax == 1.0) {
        // fused: (C-C0):eps and div_h in one sweep, tau never written (calcStressDiff fg:18030 + divOperatorStaggered fg:18853)
        if (c->nranks > 1 && (rc = fgb_comm_halo_iso(c, nullptr, c->fields[src]))) return rc;
        if ((rc = fgb_k_dir_stress_div_iso(c, nullptr, 0.0, c->fields[src], nullptr, mu0, lambda0, 1.0))) return rc;
        if ((rc = g0_staggered(c, mu0, lambda0, -1.0))) return rc;
        if (c->nranks > 1 && (rc = fgb_comm_halo_u(c))) return rc;
        return fgb_k_eps(c, c->ubuf, c->fields[dst], E);
    }
    if ((rc = fgb_k_calc_stress(c, c->fields[src], c->fields[dst], mu0, lambda0, 1.0))) return rc;            // calcStressDiff fg:18030
    if (c->mode == FGB_MODE_VISCOSITY) {
        double* tmp = nullptr;
        if (c->scheme != FGB_GAMMA_COLLOCATED) {
            if ((rc = scratch_field(c, &tmp))) return rc;
            if ((rc = fgb_k_copy(c, c->fields[dst], tmp, c->dim))) return rc;
        }
        return delta_impl(c, c->fields[dst], tmp, E, mu0, -1.0);
    }
    return gamma_impl(c, c->fields[dst], E, mu0, lambda0, -1.0, 0.0);
}

// polarizationScheme fg:20536-20553
extern "C" int fgb_polarization_step(fgb_ctx* c, int src, int dst, const double* P0, double mu0, double lambda0) {
    CHECK_CTX(c); CHECK_FIELD(c, src); CHECK_FIELD(c, dst);
    int rc;
    if (c->bc_relax != 1.0 && (rc = fgb_k_component_dot(c, c->fields[src], nullptr, c->F00, 1))) return rc;
    if ((rc = fgb_k_calc_polarization(c, c->fields[src], c->fields[dst], mu0, 0))) return rc;
    double P00[9], E[9];
    if ((rc = fgb_k_component_dot(c, c->fields[dst], nullptr, P00, 1))) return rc;
    for (int i = 0; i < c->dim; i++) E[i] = P00[i] + P0[i];
    return gamma_impl(c, c->fields[dst], E, mu0, lambda0, -4 * mu0, 1.0);
}

// krylovOperator fg:20583 / ApplyOperator fg:23132 followed by <p, p - w> (fg:23211)
extern "C" int fgb_cg_apply(fgb_ctx* c, int F, int p, int w, double mu0, double lambda0, double* pAp) {
    CHECK_CTX(c); CHECK_FIELD(c, p); CHECK_FIELD(c, w);
    if (p == w) return fgb_fail(c, FGB_EINVAL, "krylovOperator cannot work in place (fg:20581)");
    int rc;
    double zero[9] = {0};
    if (F >= 0) {
        CHECK_FIELD(c, F);
        if ((rc = fgb_k_calc_stress_deriv(c, c->fields[F], c->fields[p], c->fields[w], mu0, lambda0, 1.0))) return rc;
        if ((rc = gamma_impl(c, c->fields[w], zero, mu0, lambda0, -1.0, 0.0))) return rc;
    } else {
        if (c->bc_relax != 1.0 && (rc = fgb_k_component_dot(c, c->fields[p], nullptr, c->F00, 1))) return rc;
        if ((rc = fgb_k_calc_stress(c, c->fields[p], c->fields[w], mu0, lambda0, 1.0))) return rc;
        if (c->mode == FGB_MODE_VISCOSITY) {
            double* tmp = nullptr;
            if (c->scheme != FGB_GAMMA_COLLOCATED) {
                if ((rc = scratch_field(c, &tmp))) return rc;
                if ((rc = fgb_k_copy(c, c->fields[w], tmp, c->dim))) return rc;
            }
            if ((rc = delta_impl(c, c->fields[w], tmp, zero, mu0, -1.0))) return rc;
        } else if ((rc = gamma_impl(c, c->fields[w], zero, mu0, lambda0, -1.0, 0.0))) return rc;
    }
    if (pAp) {
        c->reduce_on_device = c->cg_dev;
        rc = fgb_k_inner(c, c->fields[p], c->fields[p], c->fields[w], pAp);
        c->reduce_on_device = false;
        return rc;
    }
    return FGB_OK;
}

// One fused CG operator application: p_new = r + beta*p_old (skipped when r < 0, then p_new must equal p_old),
// w = -Gamma0:(C-C0):p_new (or the tangent operator at F), pAp = <p_new, p_new - w>.
// 1 if fgb_cg_step / fgb_cg_update accept w = FGB_W_IMPLICIT on this context (fused linear-elastic staggered path, no BC projector)
// 0: generic kernels; 1: linear elasticity, isotropic phases, Voigt mixing (fused.cu); 2: heat / porous, scalar phases, any mixing rule
// (fused_heat.cu; builds the per-voxel conductivity on first use)
static int fused_kind(fgb_ctx* c, double lambda0) {
    if (c->bc_active || c->bc_relax != 1.0) return 0;
    if (fgb_fused_iso_applicable(c)) return 1;
    if (lambda0 == 0.0 && fgb_fused_heat_applicable(c)) {
        int diag = 0;
        if (fgb_heat_tangent(c, &diag) == FGB_OK && diag) return 2;
    }
    return 0;
}
extern "C" int fgb_cg_implicit_w_supported(const fgb_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    return fused_kind(const_cast<fgb_ctx*>(c), 0.0) ? 1 : 0;
}

extern "C" int fgb_cg_step(fgb_ctx* c, int F, int r, double beta, int p_old, int p_new, int w, double mu0, double lambda0, double* pAp) {
    CHECK_CTX(c); CHECK_FIELD(c, p_old); CHECK_FIELD(c, p_new);
    const bool implicit_w = (w == FGB_W_IMPLICIT);
    if (!implicit_w) CHECK_FIELD(c, w);
    if (r >= 0) CHECK_FIELD(c, r);
    if (r < 0 && p_old != p_new) return fgb_fail(c, FGB_EINVAL, "fgb_cg_step without a direction update needs p_new == p_old");
    if (p_new == w) return fgb_fail(c, FGB_EINVAL, "krylovOperator cannot work in place (fg:20581)");
    if (F >= 0 && c->nh_cache_of == F && c->nh_cache_mu0 == mu0 && r >= 0 && fgb_fused_nh_applicable(c)) {
        // Neo-Hooke tangent from the per-voxel cache of this Newton iterate (fgb_cg_tangent_prepare): Q = R + beta Q and the tangent
        // stress in one elementwise sweep, then div_h, G0, and either the explicit or the implicit gradient
        CHECK_FIELD(c, F);
        double* sigma = nullptr;
        double zero[9] = {0};
        int rc;
        if (implicit_w) { if ((rc = scratch_field(c, &sigma))) return rc; }
        else sigma = c->fields[w];
        if ((rc = fgb_k_nh_dir_tangent(c, c->fields[r], beta, c->fields[p_old], c->fields[p_new], sigma, lambda0))) return rc;
        if (c->nranks > 1 && (rc = fgb_comm_halo_tau(c, sigma))) return rc;
        if ((rc = fgb_k_div(c, sigma, c->ubuf))) return rc;
        if ((rc = g0_staggered(c, mu0, lambda0, -1.0))) return rc;
        if (c->nranks > 1 && (rc = fgb_comm_halo_u(c))) return rc;
        if (implicit_w) {
            if (!pAp) return fgb_fail(c, FGB_EINVAL, "w = FGB_W_IMPLICIT needs pAp");
            c->reduce_on_device = c->cg_dev;
            rc = fgb_k_hyper_cg_u(c, true, nullptr, nullptr, c->fields[p_new], 0.0, pAp);
            c->reduce_on_device = false;
            if (rc) return rc;
            c->implicit_w_of = p_new;
            return FGB_OK;
        }
        if ((rc = fgb_k_eps(c, c->ubuf, c->fields[w], zero))) return rc;
        if (pAp) {
            c->reduce_on_device = c->cg_dev;
            rc = fgb_k_inner(c, c->fields[p_new], c->fields[p_new], c->fields[w], pAp);
            c->reduce_on_device = false;
            return rc;
        }
        return FGB_OK;
    }
    const int fk = (F < 0) ? fused_kind(c, lambda0) : 0;
    if (implicit_w && !(fk && r >= 0 && p_old != p_new && pAp))
        return fgb_fail(c, FGB_EUNSUPPORTED, "w = FGB_W_IMPLICIT needs the fused linear CG step (see fgb_cg_implicit_w_supported)");
    int rc;
    if (fk == 2 && (r < 0 || p_old != p_new)) {
        double zero[9] = {0};
        if (c->nranks > 1 && (rc = fgb_comm_halo_heat(c, r >= 0 ? c->fields[r] : nullptr, c->fields[p_old]))) return rc;
        if ((rc = fgb_k_heat_march(c, r >= 0 ? c->fields[r] : nullptr, beta, c->fields[p_old], c->fields[p_new], mu0, 1.0))) return rc;
        if ((rc = g0_staggered(c, mu0, lambda0, -1.0))) return rc;
        if (c->nranks > 1 && (rc = fgb_comm_halo_u(c))) return rc;
        if (implicit_w) {
            c->reduce_on_device = c->cg_dev;
            rc = fgb_k_heat_cg_u(c, true, zero, nullptr, nullptr, c->fields[p_new], 0.0, pAp);
            c->reduce_on_device = false;
            if (rc) return rc;
            c->implicit_w_of = p_new;
            return FGB_OK;
        }
        if ((rc = fgb_k_eps(c, c->ubuf, c->fields[w], zero))) return rc;
        if (pAp) {
            c->reduce_on_device = c->cg_dev;
            rc = fgb_k_inner(c, c->fields[p_new], c->fields[p_new], c->fields[w], pAp);
            c->reduce_on_device = false;
            return rc;
        }
        return FGB_OK;
    }
    if (fk == 1 && (r < 0 || p_old != p_new)) {
        double zero[9] = {0};
        if (c->nranks > 1 && (rc = fgb_comm_halo_iso(c, r >= 0 ? c->fields[r] : nullptr, c->fields[p_old]))) return rc;
        if ((rc = fgb_k_dir_stress_div_iso(c, r >= 0 ? c->fields[r] : nullptr, beta, c->fields[p_old], c->fields[p_new], mu0, lambda0, 1.0))) return rc;
        if ((rc = g0_staggered(c, mu0, lambda0, -1.0))) return rc;
        if (c->nranks > 1 && (rc = fgb_comm_halo_u(c))) return rc;
        if (implicit_w) {
            // w = sym-grad(u) is not written out: the sum is taken on the fly and fgb_cg_update re-evaluates w from u
            c->reduce_on_device = c->cg_dev;
            rc = fgb_k_eps_dot(c, c->ubuf, nullptr, zero, c->fields[p_new], pAp);
            c->reduce_on_device = false;
            if (rc) return rc;
            c->implicit_w_of = p_new;
            return FGB_OK;
        }
        if (pAp) {
            c->reduce_on_device = c->cg_dev;
            rc = fgb_k_eps_dot(c, c->ubuf, c->fields[w], zero, c->fields[p_new], pAp);
            c->reduce_on_device = false;
            return rc;
        }
        return fgb_k_eps(c, c->ubuf, c->fields[w], zero);
    }
    if (r >= 0) {                                                                                                 // p = r + beta*p fg:23245
        if (c->cg_dev) rc = fgb_k_xpay_dev(c, c->fields[p_new], c->fields[r], 1, c->fields[p_old]);
        else rc = fgb_k_xpay(c, c->fields[p_new], c->fields[r], beta, c->fields[p_old]);
        if (rc) return rc;
    }
    return fgb_cg_apply(c, F, p_new, w, mu0, lambda0, pAp);
}

extern "C" int fgb_cg_update(fgb_ctx* c, int x, int r, int p, int w, double a, double* delta) {
    CHECK_CTX(c); CHECK_FIELD(c, x); CHECK_FIELD(c, r); CHECK_FIELD(c, p);
    if (w == FGB_W_IMPLICIT) {
        if (c->implicit_w_of != p || x == p || r == p || x == r)
            return fgb_fail(c, FGB_EINVAL, "fgb_cg_update: no implicit operator result for field %d (call fgb_cg_step with w = FGB_W_IMPLICIT first)", p);
        const double zero[9] = {0};
        c->reduce_on_device = c->cg_dev;
        const int rc = (c->dim == 3)   ? fgb_k_heat_cg_u(c, false, zero, c->fields[x], c->fields[r], c->fields[p], a, delta)
                       : (c->dim == 9) ? fgb_k_hyper_cg_u(c, false, c->fields[x], c->fields[r], c->fields[p], a, delta)
                                       : fgb_k_cg_update_implicit(c, c->ubuf, zero, c->fields[x], c->fields[r], c->fields[p], a, delta);
        c->reduce_on_device = false;
        return rc;
    }
    CHECK_FIELD(c, w);
    c->reduce_on_device = c->cg_dev;
    const int rc = fgb_k_cg_update(c, c->fields[x], c->fields[r], c->fields[p], c->fields[w], a, delta);
    c->reduce_on_device = false;
    return rc;
}

// ---- CG with device-resident scalars -------------------------------------------------------------------------------------------
// The same iteration as fgb_cg_step / fgb_cg_update (runCGElasticity fg:23206-23246), but gamma, beta and alpha never leave the
// device: a step and the following update are enqueued without any host synchronisation, and the host learns delta through a
// pinned ring buffer when it asks for it (fgb_cgdev_wait).  The decision of iteration k depends on gamma_k only (fg:23223-23226),
// so a host loop can enqueue the operator application of iteration k+1 before it waits for delta_k.
extern "C" int fgb_cgdev_begin(fgb_ctx* c, double gamma) {
    CHECK_CTX(c);
    const double init[5] = {gamma, 0.0, 0.0, 0.0, 0.0};            // gamma, beta = 0 (first direction p = r), alpha, <p,p-w>, delta
    FGB_CUDA(c, cudaMemcpyAsync(c->d_scalars, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
    FGB_CUDA(c, cudaStreamSynchronize(c->stream));
    return FGB_OK;
}
extern "C" int fgb_cgdev_step(fgb_ctx* c, int F, int r, int p_old, int p_new, int w, double mu0, double lambda0) {
    CHECK_CTX(c);
    if (r < 0) return fgb_fail(c, FGB_EINVAL, "fgb_cgdev_step needs the residual field");
    double dummy = 0;
    c->cg_dev = true;
    int rc = fgb_cg_step(c, F, r, 0.0, p_old, p_new, w, mu0, lambda0, &dummy);
    if (!rc) rc = fgb_k_cg_scalars(c, 0, 0);
    c->cg_dev = false;
    return rc;
}
extern "C" int fgb_cgdev_update(fgb_ctx* c, int x, int r, int p, int w, int slot) {
    CHECK_CTX(c);
    if (slot < 0 || slot >= FGB_CG_RING) return fgb_fail(c, FGB_EINVAL, "ring slot %d out of range 0..%d", slot, FGB_CG_RING - 1);
    double dummy = 0;
    c->cg_dev = true;
    int rc = fgb_cg_update(c, x, r, p, w, 0.0, &dummy);
    if (!rc) rc = fgb_k_cg_scalars(c, 1, slot);
    c->cg_dev = false;
    return rc;
}
extern "C" int fgb_cgdev_wait(fgb_ctx* c, int slot, double* out4) {
    CHECK_CTX(c);
    if (slot < 0 || slot >= FGB_CG_RING) return fgb_fail(c, FGB_EINVAL, "ring slot %d out of range 0..%d", slot, FGB_CG_RING - 1);
    FGB_CUDA(c, cudaEventSynchronize(c->ring_ev[slot]));
    for (int i = 0; i < 4; i++) out4[i] = c->h_ring[8 * slot + i];
    if (c->h_ring[8 * slot + 4] != 0.0) return poll_flag(c);          // a material law flagged a domain error (fg:10293) / a peer timed out
    return FGB_OK;
}

// Newton-CG (runCGHyper fg:22761-22790): F is fixed during the inner CG solve, so everything of the tangent that depends on F only is
// evaluated once here.  Returns 1 if the fused Neo-Hooke tangent path is now active for inner iterations that pass this F (staggered
// grid, Voigt mixing, all phases Neo-Hooke, no BC projector; then fgb_cg_step also accepts w = FGB_W_IMPLICIT and p_new == p_old),
// 0 if the generic tangent sweep will be used, < 0 on error.
extern "C" int fgb_cg_tangent_prepare(fgb_ctx* c, int F, double mu0, double lambda0) {
    CHECK_CTX(c); CHECK_FIELD(c, F);
    (void)lambda0;
    c->nh_cache_of = -1;
    if (!fgb_fused_nh_applicable(c)) return 0;
    int rc = fgb_k_nh_cache(c, c->fields[F], mu0);
    if (rc) return rc;
    if ((rc = poll_flag(c))) return rc;
    c->nh_cache_of = F;
    c->nh_cache_mu0 = mu0;
    return 1;
}

extern "C" int fgb_cg_direction(fgb_ctx* c, int p, int r, double beta) {
    CHECK_CTX(c); CHECK_FIELD(c, p); CHECK_FIELD(c, r);
    return fgb_k_xpay(c, c->fields[p], c->fields[r], beta, c->fields[p]);                      // p = r + beta*p fg:23245
}

extern "C" int fgb_check_numeric(fgb_ctx* c) { CHECK_CTX(c); return poll_flag(c); }

// ---- instrumentation ---------------------------------------------------------------------------------------------
extern "C" uint64_t fgb_launch_count(const fgb_ctx* c) { return c ? c->launches : 0; }
extern "C" void fgb_launch_count_reset(fgb_ctx* c) { if (c) c->launches = 0; }
extern "C" int fgb_profile_enable(fgb_ctx* c, int on) {
    CHECK_CTX(c);
    prof_resolve(c);
    c->profiling = on != 0;
    if (on) c->prof.clear();
    return FGB_OK;
}
extern "C" int fgb_profile_get(fgb_ctx* c, const char* kernel, double* total_ms, uint64_t* launches) {
    CHECK_CTX(c);
    prof_resolve(c);
    auto it = c->prof.find(kernel);
    if (it == c->prof.end()) { if (total_ms) *total_ms = 0; if (launches) *launches = 0; return FGB_OK; }
    if (total_ms) *total_ms = it->second.ms;
    if (launches) *launches = it->second.launches;
    return FGB_OK;
}
extern "C" int fgb_profile_names(fgb_ctx* c, char* buf, int buflen) {
    CHECK_CTX(c);
    prof_resolve(c);
    std::string s;
    for (auto& kv : c->prof) { if (!s.empty()) s += ","; s += kv.first; }
    snprintf(buf, buflen, "%s", s.c_str());
    return FGB_OK;
}
