/*
 * fgb200.h -- C ABI of the B200-native Lippmann-Schwinger solve loop of fibergen.
 *
 * Drop-in boundary for the hot path of fospald/fibergen.  The reference has
 * no plugin/FFI seam for this path: LSSolver<T,P,DIM> (fibergen.cpp, "fg") calls its own members.
 * This header is the seam a maintainer inserts between the scheme drivers
 * (runBasic fg:21716, runPolarization fg:21808, runCGElasticity fg:23153, runCGHyper fg:22699)
 * and everything they call.  Every entry point names the reference member(s) it replaces.
 *
 * Conventions
 *   - plain C, no exceptions cross the boundary: every call returns 0 or a negative FGB_E* code,
 *     the message is available from fgb_last_error() (reference: set_exception, fg:396).
 *   - all host pointers are caller-owned and only read/written during the call.
 *   - fields live on the device in the reference layout (fg:9549-9579, fg:227-232): `dim` planes of
 *     nx*ny*nzp doubles, nzp = 2*(nz/2+1), index (i*ny + j)*nzp + k; the complex shadow
 *     (fg:9707) is nx*ny*nzc complex numbers, nzc = nz/2+1, aliasing the same memory.
 *   - component order 11,22,33,23,13,12,32,31,21 (fg:9103); dim = 3 (heat/porous), 6 (elasticity,
 *     viscosity), 9 (hyperelasticity) (fg:14980-14997).
 *   - there is NO CPU fallback: fgb_create fails without an sm_100 device.
 */
#ifndef FGB200_H
#define FGB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fgb_ctx fgb_ctx;

/* error codes */
#define FGB_OK            0
#define FGB_EINVAL       -1   /* bad argument / unsupported combination            */
#define FGB_ENODEV       -2   /* no CUDA device of compute capability 10.x         */
#define FGB_ENOMEM       -3   /* device allocation failed                          */
#define FGB_ECUDA        -4   /* CUDA runtime error (message has the detail)       */
#define FGB_EUNSUPPORTED -5   /* feature the reference has but this path refuses   */
#define FGB_ENUMERIC     -6   /* NaN / domain error flagged on the device (fg:10293, fg:21202) */
#define FGB_ECOMM        -7   /* multi-GPU communicator error                      */

/* LSSolver::_mode (fg:14698) */
#define FGB_MODE_ELASTICITY      0
#define FGB_MODE_HYPERELASTICITY 1
#define FGB_MODE_VISCOSITY       2
#define FGB_MODE_HEAT            3
#define FGB_MODE_POROUS          4

/* LSSolver::_gamma_scheme (fg:14697); half_staggered / full_staggered are handled above this layer (dfg transfer, fgb_dfg_*) */
#define FGB_GAMMA_COLLOCATED 0
#define FGB_GAMMA_STAGGERED  1
#define FGB_GAMMA_WILLOT     2   /* GammaOperatorWillotR fg:20322 (elasticity, viscosity); needs a reference material with lambda_0 != 0 */

/* material laws (fg:15211-15294) with their parameter vectors */
#define FGB_LAW_ISO      0  /* LinearIsotropicMaterialLaw fg:11354           params: mu, lambda            */
#define FGB_LAW_GENERAL  1  /* LinearGeneralMaterialLaw fg:11233             params: C[36] row-major       */
#define FGB_LAW_TISO     2  /* LinearTransverselyIsotropic fg:11479          params: 2mu, lambda, alpha, beta, 2dmu [, ax, ay, az] */
#define FGB_LAW_SCALAR   3  /* ScalarLinearIsotropicMaterialLaw fg:11161     params: mu                    */
#define FGB_LAW_ANISO3   4  /* MatrixLinearAnisotropicMaterialLaw fg:11089   params: c11,c22,c33,c23,c13,c12 */
#define FGB_LAW_SVK      5  /* SaintVenantKirchhoffMaterialLaw fg:11598      params: mu, lambda            */
#define FGB_LAW_NH       6  /* NeoHookeMaterialLaw fg:11729                  params: mu, lambda            */
#define FGB_LAW_NH2      7  /* NeoHooke2MaterialLaw fg:11867                 params: mu, K                 */

/* mixing rules (fg:14975-15042) */
#define FGB_MIX_VOIGT    0  /* VoigtMixedMaterialLaw fg:12729    */
#define FGB_MIX_REUSS    1  /* ReussMixedMaterialLaw fg:12653    */
#define FGB_MIX_LAMINATE 2  /* LaminateMixedMaterialLaw fg:13086 (tangent="approx") */

#define FGB_MAX_PHASES 8
#define FGB_MAX_FIELDS 12

/* ---- context ------------------------------------------------------------------------------ */

/* One context per LSSolver (created where _epsilon is allocated, fg:15153).
 * rank/nranks describe the x-slab partition: this process owns
 * i in [rank*nx/nranks, (rank+1)*nx/nranks).  nranks == 1 for a single GPU.
 * device < 0 selects the current CUDA device. */
int  fgb_create(fgb_ctx** out, int nx, int ny, int nz, double Lx, double Ly, double Lz,
                int mode, int gamma_scheme, int device, int rank, int nranks);
void fgb_destroy(fgb_ctx* ctx);
const char* fgb_last_error(const fgb_ctx* ctx);     /* ctx may be NULL: last creation error */
const char* fgb_version(void);

/* Use an externally owned cudaStream_t for all launches (0 = the context's own stream). */
int  fgb_set_stream(fgb_ctx* ctx, void* cuda_stream);
int  fgb_synchronize(fgb_ctx* ctx);

/* Multi-GPU: the library talks NCCL itself.  Rank 0 calls fgb_comm_unique_id, the host side
 * broadcasts the 128 bytes (torch.distributed / MPI / a file), every rank calls fgb_comm_init. */
int  fgb_comm_unique_id(void* id128);
int  fgb_comm_init(fgb_ctx* ctx, const void* id128);

/* geometry of the local slab */
int  fgb_local_nx(const fgb_ctx* ctx);        /* number of local x-planes            */
int  fgb_local_x0(const fgb_ctx* ctx);        /* first global x index of this rank   */
size_t fgb_plane_elems(const fgb_ctx* ctx);   /* local_nx*ny*nzp doubles per component */
int  fgb_dim(const fgb_ctx* ctx);             /* tensor components: 3, 6 or 9 (fg:14980-14997) */

/* ---- fields ------------------------------------------------------------------------------- */

/* TensorField(nx,ny,nz,dim) (fg:9584): returns a field id >= 0 or an error code. */
int  fgb_field_alloc(fgb_ctx* ctx);
int  fgb_field_free(fgb_ctx* ctx, int field);
/* host <-> device in the padded plane layout; comps[c] points to local_nx*ny*nzp doubles (fg:15396 get_raw_field) */
int  fgb_field_upload(fgb_ctx* ctx, int field, const double* const* comps);
int  fgb_field_download(fgb_ctx* ctx, int field, double* const* comps);
void* fgb_field_device_ptr(fgb_ctx* ctx, int field, int comp);   /* for zero-copy interop */

/* ---- setup (LSSolver::readSettings fg:15044, initPhi fg:17489) ------------------------------ */

int  fgb_set_num_phases(fgb_ctx* ctx, int nphases);
/* half_staggered / full_staggered (use_dfg fg:14894-14897): the constitutive sweeps (calcStress, calcStressDeriv, calcPolarization, the
 * means, the reference-material scan) run on a doubly fine grid (2nx, 2ny, 2nz): prolongate_to_dfg fg:14216 before, restrict_from_dfg
 * fg:14273 after (fg:18143-18149, fg:18343-18347).  mode 0 off, 1 half_staggered: fgb_set_phase takes coarse planes and the fine
 * phases are their piecewise constant continuation (fg:17648); 2 full_staggered: fgb_set_phase / fgb_get_phase / fgb_set_normals /
 * fgb_set_orientation take planes of the fine grid, 2*local_nx * 2*ny * 2*(nz+1) doubles (fg:17154-17156, fg:14911-14937).
 * Call it right after fgb_create, before any phase data is set.  Staggered scheme, single GPU. */
int  fgb_set_dfg(fgb_ctx* ctx, int mode);
/* the transfer operators themselves between a field and the internal fine field _temp_dfg_1 (reference self-test fg:24491-24515) */
int  fgb_dfg_prolongate(fgb_ctx* ctx, int field);
int  fgb_dfg_restrict(fgb_ctx* ctx, int field);
/* Phase::phi (fg:12010): one padded plane per phase */
int  fgb_set_phase(fgb_ctx* ctx, int phase, const double* phi_plane);
int  fgb_set_law(fgb_ctx* ctx, int phase, int law_id, const double* params, int nparams);
/* get_normals()/get_orientation() (fg:14911-14937): 3 padded planes each */
int  fgb_set_normals(fgb_ctx* ctx, const double* const* comps3);
int  fgb_set_orientation(fgb_ctx* ctx, const double* const* comps3);
/* mixing rule + LaminateMixedMaterialLaw settings (fg:13110-13145):
 * lam_params = {eps_t, eps_a, eps_g, alpha, beta, delta, maxiter, backtrack, project_t, fixed_c1} or NULL for defaults */
int  fgb_set_mixing(fgb_ctx* ctx, int rule_id, const double* lam_params, int nparams);
/* Phase initialisation on the device: LSSolver::initPhi fg:17489-17581 for capsule fibres as the <place_fiber> action creates them
 * (CapsuleFiber fg:5254: centre, axis, total length L0, radius R; L0 <= 4R/3 is a sphere).  Every non-matrix phase gets the
 * composite-voxel volume fraction of integratePhiVoxel fg:16622 (smooth_levels < 0: adaptive with tolerance smooth_tol, the
 * reference defaults are -1 and 1e-3, fg:14842-14843), then normalizePhi fg:17588 (the last material has the highest priority).
 * Periodic images are separate entries of the list (the reference's ghost fibres).  x0 = cell origin (NULL: 0).  Optionally the
 * normals (gradient of the distance to the closest fibre, fg:6905-6924) and the orientation (its axis, fg:6885-6903) are
 * written for every voxel that has a fibre within reach (about 8 voxels); they are only read at interface voxels. */
typedef struct {
    double c[3];      /* centre */
    double a[3];      /* axis (need not be normalised) */
    double L0, R;     /* total length, radius */
    int material;     /* phase index */
} fgb_capsule;
int  fgb_init_phase_capsules(fgb_ctx* ctx, int nfib, const fgb_capsule* fibers, int matrix_mat, int smooth_levels, double smooth_tol,
                             const double* x0, int with_normals, int with_orientation);
/* Phase::phi back to the host: one padded plane (writeRawPhase fg:17004) */
int  fgb_get_phase(fgb_ctx* ctx, int phase, double* phi_plane);
/* freq_hack (fg:19392-19394) for the collocated elasticity operator */
int  fgb_set_freq_hack(fgb_ctx* ctx, int on);

/* ---- operator-level entry points (parity tests, post-processing) ---------------------------- */

/* TensorField BLAS-1 (fg:9799-10066) */
int  fgb_set_constant(fgb_ctx* ctx, int field, const double* c);                 /* setConstant fg:10047 */
int  fgb_add_constant(fgb_ctx* ctx, int field, const double* c);                 /* add fg:9841          */
int  fgb_copy(fgb_ctx* ctx, int src, int dst);                                   /* copyTo fg:9669       */
int  fgb_xpay(fgb_ctx* ctx, int r, int x, double a, int y);                      /* r = x + a*y      fg:9819 */
int  fgb_xpaymz(fgb_ctx* ctx, int r, int x, double a, int y, int z);             /* r = x + a*(y-z)  fg:9993 */
int  fgb_adjust_residual(fgb_ctx* ctx, int r, const double* E, int z);           /* r += E - z       fg:10012 */
/* extrapolateLoadstepPolynomial fg:21468-21513: per voxel and component p = Vinv * (f_0 .. f_{n-1}), dst = tpowers . p with
 * f_i the value of fields[i]; Vinv is n x n row-major, n <= 8.  dst may be one of the fields. */
int  fgb_extrapolate_polynomial(fgb_ctx* ctx, int n, const int* fields, const double* Vinv, const double* tpowers, int dst);

/* reductions; results are global over all ranks (fixed combine order) */
int  fgb_inner(fgb_ctx* ctx, int a, int b, int c_or_neg, double* out);           /* innerProductL2 fg:20871/20955 */
int  fgb_average(fgb_ctx* ctx, int field, double* out_dim);                      /* average fg:10171        */
int  fgb_component_dot(fgb_ctx* ctx, int a, int b, double* out_dim);             /* component_dot fg:10088  */
int  fgb_mean_pk1(fgb_ctx* ctx, int field, double alpha, double* out_dim);       /* meanPK1 fg:12312        */
int  fgb_mean_energy(fgb_ctx* ctx, int field, double* out);                      /* meanW fg:12239          */
int  fgb_min_detF(fgb_ctx* ctx, int field, double* out);                         /* calcMinDetF fg:17871    */
/* mean Cauchy stress <P(F) F^T / det F> * alpha, 9 components (hyperelasticity): meanCauchy fg:12268, Cauchy fg:10326 */
int  fgb_mean_cauchy(fgb_ctx* ctx, int field, double alpha, double* out9);
/* getRefMaterial (fg:12153-12236): min/max eigenvalue of the tangent over all voxels */
int  fgb_ref_material(fgb_ctx* ctx, int field, int zero_trace, double* lmin, double* lmax);

/* constitutive sweeps */
int  fgb_calc_stress(fgb_ctx* ctx, int src, int dst, double mu0, double lambda0, double alpha);        /* calcStress fg:18134      */
int  fgb_calc_stress_deriv(fgb_ctx* ctx, int F, int W, int dst, double mu0, double lambda0, double alpha); /* calcStressDeriv fg:18425 */
int  fgb_calc_stress_const(fgb_ctx* ctx, int src, int dst, double mu0, double lambda0);               /* calcStressConst fg:17973 */
int  fgb_calc_polarization(fgb_ctx* ctx, int src, int dst, double mu0, int inv);                       /* calcPolarization fg:18044 */

/* Mixed boundary conditions (setBCProjector fg:20599-20665): the host keeps the small-matrix algebra and hands
 * over the two dim x dim row-major matrices the operator needs, MQ = M:Q and M_QC0 = M:(Q:C0), plus bc_relax.
 * Every Green-operator application then adds alpha*R, R = bc_relax*MQ:<tau> - (1-bc_relax)*M_QC0:<eps_in>
 * (initBCProjector fg:20220/20228, calcBCProjector fg:20258, applyBCProjector fg:20263/20272).
 * NULL matrices (or ||MQ|| < eps, fg:20233) disable the term and its mean pass. */
int  fgb_set_bc(fgb_ctx* ctx, const double* MQ, const double* M_QC0, double bc_relax);

/* Green operator, in place on `field` (GammaOperator fg:20488-20531):
 *   collocated: eta^ = alpha*Gamma0^:tau^ + beta*tau^, eta^(0) = E + alpha*R
 *   staggered : eta = E + grad_h G0 div_h tau + alpha*R (beta ignored, as in the reference) */
int  fgb_gamma(fgb_ctx* ctx, int field, const double* E, double mu0, double lambda0, double alpha, double beta);
/* the individual stages, exposed for the operator-identity tests of fibergen --test (fg:23946-24583) */
int  fgb_div_staggered(fgb_ctx* ctx, int field);                                 /* divOperatorStaggered* fg:18853-19071: field -> u buffer */
int  fgb_g0_staggered(fgb_ctx* ctx, double mu0, double lambda0, double alpha);   /* G0OperatorStaggered* fg:20101-20153 on the u buffer     */
int  fgb_eps_staggered(fgb_ctx* ctx, int field, const double* E);                /* epsOperatorStaggered* fg:18614-18846: u buffer -> field */
/* displacement fluctuation of a strain field, get_raw_field("u") fg:15517-15557: u = G0 div_h tau(eps) with tau = C0:eps
 * (elasticity, heat: staggered-grid div_h and G0), the viscosity dual form (staggered), or in hyperelasticity tau = (P - C0):F with
 * the collocated G0DivOperatorHyper (fg:15524-15527); alpha = 1.
 * tmp is a scratch field (overwritten); the result is left in the u buffer (fgb_u_download: 3 components, 1 for heat). */
int  fgb_calc_displacement(fgb_ctx* ctx, int eps, int tmp, double mu0, double lambda0);
/* G0DivOperatorHyper fg:20281-20286 (fftTensor, G0DivOperatorFourierHyper fg:20155-20218, fftInvVector), in place: components 0..2 of
 * the 9-component field become alpha * G0 Div tau, collocated Fourier discretisation with xi = 2 pi m / L; 3..8 are left undefined. */
int  fgb_g0div_hyper(fgb_ctx* ctx, int field, double mu0, double lambda0, double alpha);
/* fftTensor, GradOperatorFourierHyper fg:22069-22116, fftInvTensor: components 0..2 hold a vector field q, all 9 receive grad q */
int  fgb_grad_hyper(fgb_ctx* ctx, int field);
/* fftTensor, G0DivOperatorFourierHyper, GradOperatorFourierHyper, fftInvTensor (fg:24572-24575): field <- alpha * grad G0 Div field,
 * both operators applied in Fourier space (no real-space round trip in between, which would drop the imaginary Nyquist parts) */
int  fgb_grad_g0div_hyper(fgb_ctx* ctx, int field, double mu0, double lambda0, double alpha);
/* get_raw_field("p") fg:15559-15573 (calcStressDiff, divOperatorStaggered, divVector fg:19983, poisson_solve fg:23454): the pressure
 * is left in component 0 of the u buffer (fgb_u_download with ncomp = 1); tmp is a scratch field */
int  fgb_calc_pressure(fgb_ctx* ctx, int eps, int tmp, double mu0, double lambda0);
int  fgb_u_upload(fgb_ctx* ctx, const double* const* comps, int ncomp);          /* test access to the displacement buffer */
int  fgb_u_download(fgb_ctx* ctx, double* const* comps, int ncomp);
/* fftTensor / fftInvTensor (fg:18531-18584) on all components of a field, in place (forward scaled by 1/nxyz) */
int  fgb_fft_forward(fgb_ctx* ctx, int field);
int  fgb_fft_backward(fgb_ctx* ctx, int field);

/* ---- scheme-level entry points: one iteration on the device ---------------------------------- */

/* basicScheme (fg:20558): dst = E - Gamma0:(C - C0):src (+ BC terms, fgb_set_bc); src == dst allowed (fg:21786).
 * In viscosity mode this is the Delta operator (DeltaOperatorStaggered fg:20422). */
int  fgb_basic_step(fgb_ctx* ctx, int src, int dst, const double* E, double mu0, double lambda0);
/* polarizationScheme (fg:20536): dst = (-4mu0*Gamma0 + 1):Q(src) with mean <Q> + P0; src == dst allowed */
int  fgb_polarization_step(fgb_ctx* ctx, int src, int dst, const double* P0, double mu0, double lambda0);

/* CG (runCGElasticity fg:23153-23247, inner loop of runCGHyper fg:22844-23088).
 * fgb_cg_apply: w = -Gamma0:(C-C0):p  (krylovOperator fg:20583) or, with F >= 0,
 *               w = -Gamma0:(dP/dF(F)-C0):p (ApplyOperator fg:23132); pAp (may be NULL) receives <p, p - w>.
 * fgb_cg_update: x += a*p ; r -= a*(p - w) ; returns delta = <r, r>          (fg:23221, fg:23237-23240)
 * fgb_cg_direction: p = r + beta*p                                          (fg:23245)            */
int  fgb_cg_apply(fgb_ctx* ctx, int F_or_neg, int p, int w, double mu0, double lambda0, double* pAp);
/* fused form of "fgb_cg_direction then fgb_cg_apply": p_new = r + beta*p_old (r < 0: no update, p_new == p_old), then the
 * operator on p_new.  With p_new != p_old the direction update, the constitutive law and div_h run as ONE sweep and
 * sym-grad_h + <p, p - w> as another (staggered scheme, isotropic phases, Voigt mixing); otherwise it composes the calls above. */
int  fgb_cg_step(fgb_ctx* ctx, int F_or_neg, int r_or_neg, double beta, int p_old, int p_new, int w, double mu0, double lambda0, double* pAp);
int  fgb_cg_update(fgb_ctx* ctx, int x, int r, int p, int w, double a, double* delta);
/* w = FGB_W_IMPLICIT in fgb_cg_step: the operator result is not written to a field; it stays implicit as w = sym-grad_h(u) in the
 * context's displacement buffer, <p, p - w> is summed on the fly, and the following fgb_cg_update(x, r, p_new, FGB_W_IMPLICIT, ...)
 * re-evaluates w while it updates x and r (72 B per voxel less HBM traffic per iteration).  Only on the fused path of fgb_cg_step
 * (fgb_cg_implicit_w_supported() == 1: staggered linear elasticity, isotropic phases, Voigt mixing, no BC projector) and with
 * r >= 0, p_new != p_old, pAp != NULL; the implicit result is invalidated by the next call that overwrites the u buffer. */
#define FGB_W_IMPLICIT (-2)
int  fgb_cg_implicit_w_supported(const fgb_ctx* ctx);
int  fgb_cg_direction(fgb_ctx* ctx, int p, int r, double beta);
/* Newton-CG (runCGHyper fg:22761-22790): the deformation gradient F of the outer iteration is fixed during the inner CG solve; this
 * evaluates everything of the tangent that depends on F only (per-voxel inverse, log-determinant, mixed coefficients) once.
 * Returns 1 if fgb_cg_step / fgb_cgdev_step calls that pass this F now take the fused Neo-Hooke sweeps (staggered grid, Voigt mixing,
 * all phases Neo-Hooke, no BC projector; w = FGB_W_IMPLICIT and p_new == p_old are then accepted), 0 if they take the generic
 * calcStressDeriv sweep, < 0 on error.  Any change of F or of the reference material needs a new call. */
int  fgb_cg_tangent_prepare(fgb_ctx* ctx, int F, double mu0, double lambda0);
/* The same iteration with gamma, beta and alpha resident on the device (no host synchronisation inside an iteration):
 *   fgb_cgdev_begin : gamma = <r, r> + tiny of the start residual, beta = 0
 *   fgb_cgdev_step  : p_new = r + beta*p_old ; w = operator(p_new) ; alpha = gamma / (<p_new, p_new - w> + tiny)      (fg:23245, 23209-23218)
 *   fgb_cgdev_update: x += alpha*p ; r -= alpha*(p - w) ; delta = <r, r> + tiny ; beta = delta/gamma ; gamma = delta (fg:23221-23245);
 *                     {gamma before, <p,p-w>, alpha, <r,r>} go to ring slot `slot` (0..7) of a pinned buffer
 *   fgb_cgdev_wait  : blocks until that update has finished and returns the four values.
 * The arithmetic is operation by operation that of the host-scalar form, so both give identical residual histories.  The stop
 * test of iteration k needs gamma_k only, so the caller may enqueue fgb_cgdev_step of iteration k+1 before waiting for delta_k. */
int  fgb_cgdev_begin(fgb_ctx* ctx, double gamma);
int  fgb_cgdev_step(fgb_ctx* ctx, int F_or_neg, int r, int p_old, int p_new, int w, double mu0, double lambda0);
int  fgb_cgdev_update(fgb_ctx* ctx, int x, int r, int p, int w, int slot);
int  fgb_cgdev_wait(fgb_ctx* ctx, int slot, double* out4);
/* returns FGB_ENUMERIC if a material law flagged a domain error since the last call (fg:10293, fg:21202) */
int  fgb_check_numeric(fgb_ctx* ctx);

/* ---- instrumentation ------------------------------------------------------------------------- */
/* number of kernels launched by this context since creation / last reset */
uint64_t fgb_launch_count(const fgb_ctx* ctx);
void     fgb_launch_count_reset(fgb_ctx* ctx);
/* time (ms) and launches of the dominant kernels accumulated with CUDA events when profiling is on */
int  fgb_profile_enable(fgb_ctx* ctx, int on);
int  fgb_profile_get(fgb_ctx* ctx, const char* kernel, double* total_ms, uint64_t* launches);
int  fgb_profile_names(fgb_ctx* ctx, char* buf, int buflen);   /* comma separated kernel names seen so far */

#ifdef __cplusplus
}
#endif
#endif /* FGB200_H */
