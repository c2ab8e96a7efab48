/*
 * fgb200_lssolver.h -- host side of the solve loop above the C ABI of fgb200.h.
 *
 * fgb::LSSolver mirrors the slice of the reference's LSSolver<double,double,3> (fibergen.cpp, "fg") that
 * drives the hot path: same member names, argument meaning and error behaviour
 *   readSettings keys            fg:15044-15094      run                      fg:21247-21399
 *   setBCProjector/calcBCMean    fg:20599-20665      runLoadsteppingSolver    fg:21584-21686
 *   calcRefMaterial              fg:22283-22313      runBasic/runPolarization fg:21716-21851
 *   bc_error/converged           fg:21129-21244      runCGElasticity          fg:23153-23247
 *   error estimators             fg:14344-14637      runCGHyper               fg:22699-23130
 * It keeps the iteration loop, convergence test, residual history, callbacks and the small-matrix BC algebra on
 * the host (so `get_residuals`, `maxiter`, `cancel` and Python callbacks keep working) and calls nothing but the
 * fgb_* entry points for field work.  The flat fgls_* functions at the bottom are the ctypes / cgo / JNI friendly
 * view of the same object (used by the parity tests and bench.py).
 */
#ifndef FGB200_LSSOLVER_H
#define FGB200_LSSOLVER_H

#include "fgb200.h"

#ifdef __cplusplus
#include <string>
#include <utility>
#include <vector>
#include <memory>

namespace fgb {

typedef std::vector<double> Vec;    // length dim
typedef std::vector<double> Mat;    // dim x dim row-major

class ErrorEstimator;

class LSSolver {
public:
    LSSolver(int nx, int ny, int nz, double dx, double dy, double dz, int rank = 0, int nranks = 1, int device = -1);
    ~LSSolver();

    // --- settings: same keys and defaults as LSSolver::readSettings (fg:14800-14865, fg:15044-15094)
    void set(const std::string& key, const std::string& value);
    std::string get(const std::string& key) const;
    // materials (fg:15160-15300): law in {"iso","general","tiso","aniso","nh","nh2","svk"}; "iso" resolves per mode
    int  addMaterial(const std::string& name, const std::string& law, const double* params, int nparams);
    void setReference(double mu, double lambda);                       // <ref> material (fg:15186-15194)
    void init();                                                       // allocate device state (end of readSettings)
    void initComm(const void* nccl_unique_id);                         // slab partition over nranks GPUs

    // --- phase data (initPhi fg:17489, padded planes of local_nx*ny*nzp doubles)
    void setPhase(int material, const double* phi);
    // the same on the device from the fibre list (initPhi fg:17152-17158 with the settings smooth_levels / smooth_tol)
    void initPhase(int nfib, const fgb_capsule* fibers, int matrix_mat = 0, bool normals = false, bool orientation = false);
    void getPhase(int material, double* phi);
    void setNormals(const double* const* comps3);
    void setOrientation(const double* const* comps3);

    // --- loading (fg:20667-20724)
    void setStrain(const Vec& E);
    void setStress(const Vec& S);
    void setBCProjector(const Mat& P);

    // --- solve
    bool run();                                                        // fg:21247; returns true on error/cancel
    void cancel();                                                     // fg:25190
    typedef bool (*ConvergenceCallback)(void* user);
    void setConvergenceCallback(ConvergenceCallback cb, void* user);   // fg:21215

    // --- results
    const std::vector<double>& getResiduals() const { return _residuals; }          // fg:15389
    double getSolveTime() const { return _solve_time; }                              // fg:15391
    Vec    calcMeanStress();                                                         // fg:17793
    Vec    calcMeanStrain();                                                         // _epsilon->average()
    double calcMeanEnergy();                                                         // fg:17765
    Vec    calcMeanCauchyStress();                                                   // fg:17920 (hyperelasticity)
    Mat    calcEffectiveProperties();                                                // fg:26030-26160 (Voigt form)
    void   getField(const std::string& name, double* const* comps);                 // get_raw_field fg:15396: "epsilon", "sigma", "u", "p"
    int    fieldComponents(const std::string& name) const;                          // planes getField(name) writes
    double mu0() const { return _mu_0; }
    double lambda0() const { return _lambda_0; }
    int    dim() const { return _dim; }
    int    localNx() const;
    size_t planeElems() const;
    const std::string& lastError() const { return _error; }
    fgb_ctx* ctx() { return _ctx; }
    int epsilonField() { syncEpsilon(); return _epsilon; }
    unsigned long long launches() const;

    // --- pieces of the reference kept public for the parity tests
    Vec  calcBCMean(const Vec& E, const Vec& S) const;                 // fg:20242
    void calcRefMaterial();                                            // fg:22283
    double bc_error();                                                 // fg:21129
    const Mat& BC_M() const { return _BC_M; }
    const Mat& BC_MQ() const { return _BC_MQ; }

private:
    friend class ErrorEstimator;
    void check(int rc) const;
    void fail(const std::string& msg) const;
    void set_impl(const std::string& key, const std::string& value);
    Vec  expandLoad(const Vec& e, const char* what) const;
    void extrapolateLoadstep(const std::vector<std::pair<double, int>>& last, double t);   // fg:21454-21513
    void pushBC();
    void syncEpsilon();
    bool runLoadsteppingSolver(const Vec& Emax, const Vec& Smax);
    void runSolver(const Vec& E, const Vec& S);
    void runBasic(const Vec& E0, const Vec& S0);
    void runPolarization(const Vec& E0, const Vec& S0);
    void runCGElasticity(const Vec& E0, const Vec& S0);
    void runCGElasticityPipelined(const Vec& E, int r, int p, int p2, int w, double gamma, ErrorEstimator* ee);
    void runCGHyper(const Vec& E0, const Vec& S0);
    bool converged(size_t& iter, double abs_err, double rel_err, bool check_bc = true);
    ErrorEstimator* create_error_estimator(const std::string& name = "");
    int  field(int& slot);                                              // lazily allocated work fields

    int _nx, _ny, _nz;
    double _dx, _dy, _dz;
    int _rank, _nranks, _device;
    fgb_ctx* _ctx;
    int _dim;

    // settings
    double _tol, _abs_tol, _bc_tol, _ref_scale, _newton_relax, _bc_relax;
    size_t _maxiter, _cg_reinit;
    std::string _update_ref, _error_estimator, _outer_error_estimator, _method, _gamma_scheme, _mode, _mixing_rule,
        _cg_inner_product, _G0_solver;
    bool _freq_hack;
    int _smooth_levels;                            // fg:14670-14671
    double _smooth_tol;
    bool _pipelined_cg;                            // CG scalars resident on the device (default); false: host-scalar loop
    std::vector<double> _loadsteps;
    size_t _loadstep_extrapolation_order;          // 0 = none, 1 = linear, ... (fg:14696)
    std::string _loadstep_extrapolation_method;    // "polynomial" (fg:14697; "transformation" is refused)
    long _first_loadstep;                          // < 0: automatic (fg:21591)
    std::vector<double> _laminate_params;

    struct MaterialDef {
        std::string name, law;
        std::vector<double> params;
    };
    std::vector<MaterialDef> _materials;
    bool _reference_set;
    double _mu_0, _lambda_0;

    Vec _E, _S, _current_E, _current_S, _Id;
    Vec _E_raw, _S_raw;                            // loads as given by the caller (expanded to dim by init())
    Mat _BC_P, _BC_Q, _BC_QC0, _BC_M, _BC_MQ;

    int _epsilon, _f1, _f2, _f3, _f4, _f5;    // device field ids (-1 = not allocated)
    bool _eps_stale;                          // Newton-CG: _epsilon still has to be formed as F + newton_relax*X (fg:23049)
    int _eps_F, _eps_X;
    std::vector<double> _residuals;
    double _solve_time;
    bool _cancel;
    ConvergenceCallback _cb;
    void* _cb_user;
    mutable std::string _error;
};

}  // namespace fgb

extern "C" {
#endif /* __cplusplus */

/* ---- flat view of fgb::LSSolver (ctypes / cgo / JNI).  All functions return 0 or a negative FGB_E* code,
 *      the message is in fgls_last_error(). -------------------------------------------------------------- */
typedef struct fgls_solver fgls_solver;
typedef int (*fgls_callback)(void* user);          /* return non-zero to stop (set_convergence_callback fg:27160) */

int  fgls_create(fgls_solver** out, int nx, int ny, int nz, double dx, double dy, double dz, int rank, int nranks, int device);
void fgls_destroy(fgls_solver* s);
const char* fgls_last_error(const fgls_solver* s);
int  fgls_set(fgls_solver* s, const char* key, const char* value);             /* readSettings keys */
int  fgls_add_material(fgls_solver* s, const char* name, const char* law, const double* params, int nparams);
int  fgls_set_reference(fgls_solver* s, double mu, double lambda);
int  fgls_init(fgls_solver* s);
int  fgls_init_comm(fgls_solver* s, const void* id128);
int  fgls_set_phase(fgls_solver* s, int material, const double* phi);
int  fgls_init_phase_capsules(fgls_solver* s, int nfib, const fgb_capsule* fibers, int matrix_mat, int normals, int orientation);
int  fgls_get_phase(fgls_solver* s, int material, double* phi);
int  fgls_set_normals(fgls_solver* s, const double* const* comps3);
int  fgls_set_orientation(fgls_solver* s, const double* const* comps3);
int  fgls_set_strain(fgls_solver* s, const double* E);
int  fgls_set_stress(fgls_solver* s, const double* S);
int  fgls_set_bc_projector(fgls_solver* s, const double* P);                  /* dim x dim row-major */
int  fgls_set_callback(fgls_solver* s, fgls_callback cb, void* user);
int  fgls_run(fgls_solver* s);
int  fgls_cancel(fgls_solver* s);
int  fgls_num_residuals(const fgls_solver* s);
int  fgls_get_residuals(const fgls_solver* s, double* out, int n);
int  fgls_mean_stress(fgls_solver* s, double* out);
int  fgls_mean_strain(fgls_solver* s, double* out);
int  fgls_mean_energy(fgls_solver* s, double* out);
int  fgls_mean_cauchy_stress(fgls_solver* s, double* out9);
int  fgls_field_components(fgls_solver* s, const char* name);      /* planes fgls_get_field writes; < 0: unknown field */
int  fgls_effective_properties(fgls_solver* s, double* Ceff_voigt);            /* dim x dim row-major */
int  fgls_get_field(fgls_solver* s, const char* name, double* const* comps);
int  fgls_ref_material(fgls_solver* s, double* mu0, double* lambda0);
int  fgls_calc_ref_material(fgls_solver* s);
int  fgls_bc_matrices(fgls_solver* s, double* M, double* MQ);
int  fgls_dim(const fgls_solver* s);
int  fgls_local_nx(const fgls_solver* s);
double fgls_solve_time(const fgls_solver* s);
unsigned long long fgls_launches(const fgls_solver* s);
fgb_ctx* fgls_ctx(fgls_solver* s);

#ifdef __cplusplus
}
#endif
#endif /* FGB200_LSSOLVER_H */
