// Host-side mirror of the reference's LSSolver scheme drivers (see include/fgb200_lssolver.h).
// Pure host C++ on top of the C ABI of fgb200.h: loops, convergence logic, estimators, BC algebra.
#include "../../include/fgb200_lssolver.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <sstream>
#include <stdexcept>
#include <nvtx3/nvToolsExt.h>

namespace fgb {

// NVTX range with the name of the reference's Timer for the same scope (fg:1643-1737; "runCGElasticity" fg:23155, "CG loop", ...)
struct NvtxRange {
    explicit NvtxRange(const char* n) { nvtxRangePushA(n); }
    ~NvtxRange() { nvtxRangePop(); }
};

static const double EPS = std::numeric_limits<double>::epsilon();
static const double SMALL = std::numeric_limits<double>::min();   // boost::numeric::bounds<T>::smallest()

// ---- Voigt helpers (fg:494-598) -------------------------------------------------------------------
static Mat Id4(int dim) {
    Mat I(dim * dim, 0.0);
    for (int i = 0; i < dim; i++) I[i * dim + i] = (dim == 6 && i >= 3) ? 0.5 : 1.0;
    return I;
}
static Mat II4(int dim) {
    Mat M(dim * dim, 0.0);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) M[i * dim + j] = 1.0;
    return M;
}
static inline double vw(int dim, int k) { return (dim == 6 && k >= 3) ? 2.0 : 1.0; }
static Vec dyad4(const Mat& M, const Vec& v) {
    const int d = (int)v.size();
    Vec r(d, 0.0);
    for (int i = 0; i < d; i++) {
        double s = 0;
        for (int k = 0; k < d; k++) s += M[i * d + k] * (v[k] * vw(d, k));
        r[i] = s;
    }
    return r;
}
static Mat dyad4(const Mat& A, const Mat& B, int d) {
    Mat C(d * d, 0.0);
    for (int i = 0; i < d; i++)
        for (int j = 0; j < d; j++) {
            double s = 0;
            for (int k = 0; k < d; k++) s += A[i * d + k] * (B[k * d + j] * vw(d, k));
            C[i * d + j] = s;
        }
    return C;
}
static double norm_2(const Vec& v) {
    double s = 0;
    for (double x : v) s += x * x;
    if (v.size() == 6) s += v[3] * v[3] + v[4] * v[4] + v[5] * v[5];
    return std::sqrt(s);
}
static double plain_norm(const Vec& v) {
    double s = 0;
    for (double x : v) s += x * x;
    return std::sqrt(s);
}
static double frob(const Mat& M) { return plain_norm(M); }
static Vec operator+(const Vec& a, const Vec& b) { Vec r(a); for (size_t i = 0; i < r.size(); i++) r[i] += b[i]; return r; }
static Vec operator-(const Vec& a, const Vec& b) { Vec r(a); for (size_t i = 0; i < r.size(); i++) r[i] -= b[i]; return r; }
static Vec operator*(double s, const Vec& a) { Vec r(a); for (double& x : r) x *= s; return r; }
// fix_dim fg:12115-12125
static Vec fix_dim(const Vec& t, int dim) {
    Vec r(9, 0.0);
    for (int i = 0; i < dim; i++) r[i] = t[i];
    if (dim == 6) { r[6] = r[3]; r[7] = r[4]; r[8] = r[5]; }
    return r;
}

// symmetric eigen-decomposition (cyclic Jacobi): A = V diag(w) V^T ; A is n x n row-major
static void jacobi_eig(Mat A, int n, Vec& w, Mat& V) {
    V.assign(n * n, 0.0);
    for (int i = 0; i < n; i++) V[i * n + i] = 1.0;
    for (int sweep = 0; sweep < 100; sweep++) {
        double off = 0;
        for (int p = 0; p < n; p++)
            for (int q = p + 1; q < n; q++) off += A[p * n + q] * A[p * n + q];
        if (off < 1e-300) break;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) {
                const double apq = A[p * n + q];
                if (apq == 0) continue;
                const double theta = (A[q * n + q] - A[p * n + p]) / (2 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
                const double c = 1 / std::sqrt(t * t + 1), s = t * c;
                for (int k = 0; k < n; k++) {
                    const double akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = c * akp - s * akq;
                    A[k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; k++) {
                    const double apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = c * apk - s * aqk;
                    A[q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; k++) {
                    const double vkp = V[k * n + p], vkq = V[k * n + q];
                    V[k * n + p] = c * vkp - s * vkq;
                    V[k * n + q] = s * vkp + c * vkq;
                }
            }
    }
    w.resize(n);
    for (int i = 0; i < n; i++) w[i] = A[i * n + i];
}

// ---- error estimators (fg:14344-14637) ----------------------------------------------------------------
class ErrorEstimator {
public:
    virtual ~ErrorEstimator() {}
    virtual void update() { throw std::runtime_error("Selected error estimator is not compatible with the selected solution method"); }
    virtual void update_cg(double, double) { throw std::runtime_error("Selected error estimator is not compatible with the selected solution method"); }
    double rel_error() const { return _rel; }
    double abs_error() const { return _abs; }
protected:
    double _abs = std::numeric_limits<double>::infinity(), _rel = 1;
};
class NoneEE : public ErrorEstimator {
public:
    NoneEE() { _abs = 1; _rel = 1; }
    void update() override {}
    void update_cg(double, double) override {}
};
class ResidualEE : public ErrorEstimator {
public:
    void update_cg(double gamma, double gamma0) override { _abs = std::sqrt(gamma); _rel = std::sqrt(gamma / gamma0); }
};
class EpsilonEE : public ErrorEstimator {
    LSSolver* s;
    Vec prev;
    Vec mean();
public:
    explicit EpsilonEE(LSSolver* s_) : s(s_) { prev = mean(); }
    void update() override {
        Vec cur = mean();
        _abs = std::fabs(plain_norm(prev) - plain_norm(cur));
        _rel = _abs / (SMALL + plain_norm(cur));
        prev = cur;
    }
    void update_cg(double, double) override { update(); }
};
class SigmaEE : public ErrorEstimator {
    LSSolver* s;
    Vec cur, prev, prev_prev;
    size_t iter = 0;
    Vec mean() { return fix_dim(s->calcMeanStress(), s->dim()); }
public:
    explicit SigmaEE(LSSolver* s_) : s(s_) { cur = mean(); prev = cur; prev_prev = cur; }
    void update() override {
        cur = mean();
        if (iter > 1) _abs = 0.5 * (plain_norm(prev_prev - cur) + plain_norm(prev - cur));
        else _abs = plain_norm(prev - cur);
        _rel = _abs / (SMALL + plain_norm(cur));
        prev_prev = prev;
        prev = cur;
        iter++;
    }
    void update_cg(double, double) override { update(); }
};
class EnergyEE : public ErrorEstimator {
    LSSolver* s;
    double prev;
public:
    explicit EnergyEE(LSSolver* s_) : s(s_) { prev = s->calcMeanEnergy(); }
    void update() override {
        const double cur = s->calcMeanEnergy();
        _abs = std::fabs(prev - cur);
        _rel = _abs / (SMALL + std::fabs(cur));
        prev = cur;
    }
    void update_cg(double, double) override { update(); }
};

// ---- LSSolver ----------------------------------------------------------------------------------------------
LSSolver::LSSolver(int nx, int ny, int nz, double dx, double dy, double dz, int rank, int nranks, int device)
    : _nx(nx), _ny(ny), _nz(nz), _dx(dx), _dy(dy), _dz(dz), _rank(rank), _nranks(nranks), _device(device), _ctx(nullptr), _dim(6) {
    // defaults fg:14800-14865
    _tol = 1e-4;
    _abs_tol = EPS;
    _bc_tol = 1e-3;
    _maxiter = 10000;
    _ref_scale = 1.0;
    _newton_relax = 1.0;
    _bc_relax = 1.0;
    _error_estimator = "epsilon";
    _outer_error_estimator = "epsilon";
    _update_ref = "loadstep";
    _method = "cg";
    _cg_inner_product = "l2";
    _cg_reinit = 0;
    _mode = "elasticity";
    _gamma_scheme = "auto";
    _freq_hack = false;
    _mixing_rule = "voigt";
    _G0_solver = "fft";
    _loadsteps = {0.0, 1.0};
    _loadstep_extrapolation_order = 0;                                                       // fg:14830-14831
    _loadstep_extrapolation_method = "polynomial";
    _first_loadstep = -1;
    _pipelined_cg = true;
    _eps_stale = false;
    _eps_F = _eps_X = -1;
    _smooth_levels = -1;                                                                     // fg:14842-14843
    _smooth_tol = 0.001;
    // the reference sizes the prescribed loads in its constructor (mode is known there); here the mode may still change until
    // init(), so loads set earlier are kept as given and expanded to the tensor dimension by init()
    _E.assign(_dim, 0.0);
    _S.assign(_dim, 0.0);
    _current_E.assign(_dim, 0.0);
    _current_S.assign(_dim, 0.0);
    _Id.assign(_dim, 0.0);
    _Id[0] = _Id[1] = _Id[2] = 1;
    _mu_0 = _lambda_0 = 0;
    _reference_set = false;
    _epsilon = _f1 = _f2 = _f3 = _f4 = _f5 = -1;
    _solve_time = 0;
    _cancel = false;
    _cb = nullptr;
    _cb_user = nullptr;
}

LSSolver::~LSSolver() {
    if (_ctx) fgb_destroy(_ctx);
}

void LSSolver::fail(const std::string& msg) const {
    _error = msg;
    throw std::runtime_error(msg);
}

void LSSolver::check(int rc) const {
    if (rc >= 0) return;
    fail(std::string(fgb_last_error(_ctx)));
}

// the reference's pt_get throws on malformed values (fg:889-946): no silent 0
static double to_double(const std::string& v) {
    char* end = nullptr;
    const double x = std::strtod(v.c_str(), &end);
    while (end && (*end == ' ' || *end == '\t')) end++;
    if (v.empty() || end == v.c_str() || (end && *end != 0)) throw std::runtime_error("Invalid numeric value '" + v + "'");
    return x;
}
static size_t to_size(const std::string& v) {
    char* end = nullptr;
    const long long x = std::strtoll(v.c_str(), &end, 10);
    while (end && (*end == ' ' || *end == '\t')) end++;
    if (v.empty() || end == v.c_str() || (end && *end != 0) || x < 0) throw std::runtime_error("Invalid non-negative integer value '" + v + "'");
    return (size_t)x;
}
static bool to_bool(const std::string& v) {
    if (v == "1" || v == "true" || v == "True" || v == "yes") return true;
    if (v == "0" || v == "false" || v == "False" || v == "no") return false;
    throw std::runtime_error("Invalid boolean value '" + v + "'");
}

void LSSolver::set(const std::string& key, const std::string& value) {
    try {
        set_impl(key, value);
    } catch (const std::exception& e) {
        const std::string msg = e.what();
        if (msg.rfind("Invalid ", 0) == 0) fail("solver setting '" + key + "': " + msg);
        throw;
    }
}

void LSSolver::set_impl(const std::string& key, const std::string& value) {
    if (key == "tol") _tol = to_double(value);
    else if (key == "abs_tol") _abs_tol = to_double(value);
    else if (key == "bc_tol") _bc_tol = to_double(value);
    else if (key == "maxiter") _maxiter = to_size(value);
    else if (key == "update_ref") _update_ref = value;
    else if (key == "ref_scale") _ref_scale = to_double(value);
    else if (key == "newton_relax") _newton_relax = to_double(value);
    else if (key == "error_estimator") _error_estimator = value;
    else if (key == "outer_error_estimator") _outer_error_estimator = value;
    else if (key == "method") _method = value;
    else if (key == "cg_inner_product") _cg_inner_product = value;
    else if (key == "cg_reinit") _cg_reinit = to_size(value);
    else if (key == "loadstep_extrapolation_order") _loadstep_extrapolation_order = to_size(value);          // fg:15090
    else if (key == "loadstep_extrapolation_method") _loadstep_extrapolation_method = value;                 // fg:15091
    else if (key == "first_loadstep") {
        _first_loadstep = (value.size() && value[0] == '-') ? -1 : (long)to_size(value);
    }
    else if (key == "gamma_scheme") _gamma_scheme = value;
    else if (key == "mode") {
        _mode = value;
        // the tensor dimension follows the mode (fg:14980-14997) so that loads given before init() are sized correctly
        if (!_ctx) _dim = (value == "hyperelasticity") ? 9 : ((value == "heat" || value == "porous") ? 3 : 6);
    }
    else if (key == "bc_relax") _bc_relax = to_double(value);
    else if (key == "freq_hack") _freq_hack = to_bool(value);
    else if (key == "G0_solver") _G0_solver = value;
    else if (key == "mixing_rule") _mixing_rule = value;
    else if (key == "smooth_levels") _smooth_levels = (value.size() && value[0] == '-') ? -(int)to_size(value.substr(1)) : (int)to_size(value);   // fg:15055
    else if (key == "smooth_tol") _smooth_tol = to_double(value);                                                                              // fg:15056
    else if (key == "pipelined_cg") _pipelined_cg = to_bool(value);      // not a reference key: host-scalar CG loop when false (A/B parity)
    else if (key == "loadsteps") {
        // uniform_loadsteps(n) (fg:15033) or an explicit comma separated parameter list
        _loadsteps.clear();
        if (value.find(',') == std::string::npos) {
            const size_t n = to_size(value);
            if (n < 1) fail("loadsteps must be at least 1");
            for (size_t i = 0; i <= n; i++) _loadsteps.push_back(i / (double)n);
        } else {
            std::stringstream ss(value);
            std::string tok;
            while (std::getline(ss, tok, ',')) _loadsteps.push_back(to_double(tok));
        }
    } else if (key.rfind("laminate_mixing.", 0) == 0) {
        // eps_t, eps_a, eps_g, alpha, beta, delta, maxiter, backtrack, project_t, fixed_c1 (fg:13130-13145)
        static const char* names[10] = {"eps_t", "eps_a", "eps_g", "alpha", "beta", "delta", "maxiter", "backtrack", "project_t", "fixed_c1"};
        if (_laminate_params.empty()) {
            _laminate_params = {4 * EPS, std::pow(EPS, 2.0 / 3.0), EPS, 0.001, 0.1, 1 - 1024 * EPS, 32, 1, 1, -1.0};
        }
        const std::string sub = key.substr(16);
        bool found = false;
        for (int i = 0; i < 10; i++)
            if (sub == names[i]) { _laminate_params[i] = (i == 7 || i == 8) ? (to_bool(value) ? 1.0 : 0.0) : to_double(value); found = true; }
        if (!found) fail("Unknown laminate_mixing setting '" + sub + "'");
    } else {
        fail("Unknown solver setting '" + key + "'");
    }
}

std::string LSSolver::get(const std::string& key) const {
    std::ostringstream o;
    o.precision(17);
    if (key == "tol") o << _tol;
    else if (key == "method") o << _method;
    else if (key == "mode") o << _mode;
    else if (key == "gamma_scheme") o << _gamma_scheme;
    else if (key == "error_estimator") o << _error_estimator;
    else if (key == "mixing_rule") o << _mixing_rule;
    else if (key == "maxiter") o << _maxiter;
    else fail("Unknown solver setting '" + key + "'");
    return o.str();
}

int LSSolver::addMaterial(const std::string& name, const std::string& law, const double* params, int nparams) {
    MaterialDef m;
    m.name = name;
    m.law = law;
    m.params.assign(params, params + nparams);
    _materials.push_back(m);
    return (int)_materials.size() - 1;
}

void LSSolver::setReference(double mu, double lambda) {
    _mu_0 = mu;
    _lambda_0 = lambda;
    _reference_set = true;
}

void LSSolver::init() {
    // scheme resolution fg:15066-15079
    if (_gamma_scheme == "full-staggered") _gamma_scheme = "full_staggered";
    else if (_gamma_scheme == "half-staggered") _gamma_scheme = "half_staggered";
    else if (_gamma_scheme == "Willot-R") _gamma_scheme = "willot";
    else if (_gamma_scheme == "auto") {
        _gamma_scheme = "staggered";
        if (_method == "polarization") _gamma_scheme = "collocated";
    }
    if (_method == "polarization" && _gamma_scheme != "collocated") _gamma_scheme = "collocated";   // fg:15074-15079
    int mode;
    if (_mode == "elasticity") mode = FGB_MODE_ELASTICITY;
    else if (_mode == "hyperelasticity") mode = FGB_MODE_HYPERELASTICITY;
    else if (_mode == "viscosity") mode = FGB_MODE_VISCOSITY;
    else if (_mode == "heat") mode = FGB_MODE_HEAT;
    else if (_mode == "porous") mode = FGB_MODE_POROUS;
    else { fail("Unknown mode '" + _mode + "'"); return; }
    int scheme;
    if (_gamma_scheme == "collocated") scheme = FGB_GAMMA_COLLOCATED;
    else if (_gamma_scheme == "staggered") scheme = FGB_GAMMA_STAGGERED;
    else if (_gamma_scheme == "willot") scheme = FGB_GAMMA_WILLOT;
    else if (_gamma_scheme == "half_staggered" || _gamma_scheme == "full_staggered") scheme = FGB_GAMMA_STAGGERED;   // fg:20480, fg:20496
    else { fail("Unknown gamma scheme '" + _gamma_scheme + "' (this build provides collocated, staggered, half_staggered, full_staggered and willot)"); return; }
    if (_loadstep_extrapolation_method != "polynomial")
        fail("Unknown loadstep extrapolation method '" + _loadstep_extrapolation_method + "' (this build provides polynomial)");
    if (_G0_solver != "fft") fail("Unknown G0-solver '" + _G0_solver + "' (multigrid is not provided)");
    if (_method != "basic" && _method != "cg" && _method != "polarization") fail("Unknown solver method '" + _method + "'");
    if (_cg_inner_product != "l2") fail("Unknown inner product '" + _cg_inner_product + "'");   // "energy" throws in the reference too (fg:20792)

    if (_ctx) { fgb_destroy(_ctx); _ctx = nullptr; }
    int rc = fgb_create(&_ctx, _nx, _ny, _nz, _dx, _dy, _dz, mode, scheme, _device, _rank, _nranks);
    if (rc) fail(std::string(fgb_last_error(nullptr)));
    _dim = fgb_dim(_ctx);
    _epsilon = _f1 = _f2 = _f3 = _f4 = _f5 = -1;
    if (_gamma_scheme == "half_staggered") check(fgb_set_dfg(_ctx, 1));
    else if (_gamma_scheme == "full_staggered") check(fgb_set_dfg(_ctx, 2));

    if (_materials.empty()) fail("No materials specified");                               // fg:15306
    check(fgb_set_num_phases(_ctx, (int)_materials.size()));
    for (size_t i = 0; i < _materials.size(); i++) {
        const MaterialDef& m = _materials[i];
        int id = -1;
        // law name resolution per mode, fg:15211-15294
        if (_mode == "elasticity" && m.law == "iso") id = FGB_LAW_ISO;
        else if (_mode == "elasticity" && m.law == "general") id = FGB_LAW_GENERAL;
        else if (_mode == "elasticity" && m.law == "tiso") id = FGB_LAW_TISO;
        else if ((_mode == "heat" || _mode == "porous") && m.law == "iso") id = FGB_LAW_SCALAR;
        else if ((_mode == "heat" || _mode == "porous") && m.law == "aniso") id = FGB_LAW_ANISO3;
        else if (_mode == "viscosity" && m.law == "iso") id = FGB_LAW_SCALAR;
        else if (_mode == "hyperelasticity" && m.law == "iso") id = FGB_LAW_SVK;
        else if (_mode == "hyperelasticity" && m.law == "nh") id = FGB_LAW_NH;
        else if (_mode == "hyperelasticity" && m.law == "nh2") id = FGB_LAW_NH2;
        else fail("Unknown material law '" + m.law + "'");
        std::vector<double> p = m.params;
        if (_mode == "viscosity" && id == FGB_LAW_SCALAR) p[0] *= 0.5;                     // fg:15238
        check(fgb_set_law(_ctx, (int)i, id, p.data(), (int)p.size()));
    }
    int mix;
    if (_mixing_rule == "voigt") mix = FGB_MIX_VOIGT;
    else if (_mixing_rule == "reuss") mix = FGB_MIX_REUSS;
    else if (_mixing_rule == "laminate") mix = FGB_MIX_LAMINATE;
    else { fail("Unknown material mixing rule '" + _mixing_rule + "'"); return; }
    check(fgb_set_mixing(_ctx, mix, _laminate_params.empty() ? nullptr : _laminate_params.data(), 10));
    check(fgb_set_freq_hack(_ctx, _freq_hack ? 1 : 0));

    _epsilon = fgb_field_alloc(_ctx);
    check(_epsilon);
    _current_E.assign(_dim, 0.0);
    _current_S.assign(_dim, 0.0);
    _Id.assign(_dim, 0.0);
    _Id[0] = _Id[1] = _Id[2] = 1;
    // loads given before init() (or before a re-init) stay in force
    _E = _E_raw.empty() ? Vec(_dim, 0.0) : expandLoad(_E_raw, "strain");
    _S = _S_raw.empty() ? Vec(_dim, 0.0) : expandLoad(_S_raw, "stress");
    if (!_reference_set) {
        _mu_0 = std::numeric_limits<double>::quiet_NaN();                                  // fg:15340
        _lambda_0 = 0.0;
    }
    _BC_P = Id4(_dim);
    setBCProjector(Id4(_dim));
}

void LSSolver::initComm(const void* id) { check(fgb_comm_init(_ctx, id)); }

void LSSolver::setPhase(int m, const double* phi) { check(fgb_set_phase(_ctx, m, phi)); }
void LSSolver::setNormals(const double* const* c) { check(fgb_set_normals(_ctx, c)); }
void LSSolver::setOrientation(const double* const* c) { check(fgb_set_orientation(_ctx, c)); }

Vec LSSolver::expandLoad(const Vec& e, const char* what) const {
    // fg:20691-20712 (setStrain) / fg:20667-20689 (setStress): 3-, 6- or 9-vectors, shear entries duplicated for dim 9
    Vec r(_dim, 0.0);
    if (e.size() == 3 && _dim == 3) r = e;
    else if (e.size() == 6 && _dim >= 6) {
        for (int i = 0; i < 6; i++) r[i] = e[i];
        if (_dim == 9) { r[6] = e[3]; r[7] = e[4]; r[8] = e[5]; }
    } else if (e.size() == 9 && _dim == 9) r = e;
    else fail(std::string("Invalid size of ") + what + " vector");
    return r;
}

void LSSolver::initPhase(int nfib, const fgb_capsule* fibers, int matrix_mat, bool normals, bool orientation) {
    // initPhi fg:17152-17158 -> fg:17489 with the solver's smooth_levels / smooth_tol
    if (!_ctx) fail("solver not initialised");
    check(fgb_init_phase_capsules(_ctx, nfib, fibers, matrix_mat, _smooth_levels, _smooth_tol, nullptr, normals ? 1 : 0, orientation ? 1 : 0));
}
void LSSolver::getPhase(int m, double* phi) { check(fgb_get_phase(_ctx, m, phi)); }

void LSSolver::setStrain(const Vec& e) {
    if (e.size() != 3 && e.size() != 6 && e.size() != 9) fail("Invalid size of strain vector");
    _E_raw = e;
    if (_ctx) _E = expandLoad(e, "strain");
}

void LSSolver::setStress(const Vec& e) {
    if (e.size() != 3 && e.size() != 6 && e.size() != 9) fail("Invalid size of stress vector");
    _S_raw = e;
    if (_ctx) _S = expandLoad(e, "stress");
}

void LSSolver::setBCProjector(const Mat& P) {
    // fg:20599-20665
    const int dim = _dim;
    const double eps = std::sqrt(EPS);
    if ((int)P.size() != dim * dim) fail("Projector is not symmetric");
    {
        Mat D(P);
        for (int i = 0; i < dim; i++)
            for (int j = 0; j < dim; j++) D[i * dim + j] -= P[j * dim + i];
        if (frob(D) > eps) fail("Projector is not symmetric");
        Mat PP = dyad4(P, P, dim);
        for (int i = 0; i < dim * dim; i++) PP[i] = P[i] - PP[i];
        if (frob(PP) > eps) fail("Specified Projector is not a projector");
    }
    Mat C0 = Id4(dim), II = II4(dim);
    for (int i = 0; i < dim * dim; i++) C0[i] = 2 * _mu_0 * C0[i] + _lambda_0 * II[i];
    _BC_P = P;
    _BC_Q = Id4(dim);
    for (int i = 0; i < dim * dim; i++) _BC_Q[i] -= P[i];
    _BC_QC0 = dyad4(_BC_Q, C0, dim);
    Mat QC0Q = dyad4(_BC_QC0, _BC_Q, dim);
    const int edim = (dim == 6) ? 9 : dim;
    Mat A(edim * edim, 0.0);
    if (dim == 6) {
        for (int i = 0; i < 9; i++)
            for (int j = i; j < 9; j++) A[j * 9 + i] = A[i * 9 + j] = QC0Q[(i < 6 ? i : i - 3) * 6 + (j < 6 ? j : j - 3)];
    } else {
        A = QC0Q;
    }
    // Moore-Penrose pseudo inverse via the SVD of the symmetric matrix (gesvd fg:20642): singular values |w|
    Mat M(edim * edim, 0.0);
    bool finite = true;
    for (double x : A) finite = finite && std::isfinite(x);
    if (finite) {
        Vec w;
        Mat V;
        jacobi_eig(A, edim, w, V);
        double ns = 0;
        for (double x : w) ns += x * x;
        const double cut = std::sqrt(EPS) * std::sqrt(ns);
        for (int k = 0; k < edim; k++) {
            if (std::fabs(w[k]) > cut) {
                const double inv = 1.0 / w[k];
                for (int i = 0; i < edim; i++)
                    for (int j = 0; j < edim; j++) M[i * edim + j] += V[i * edim + k] * inv * V[j * edim + k];
            }
        }
    } else {
        std::fill(M.begin(), M.end(), std::numeric_limits<double>::quiet_NaN());
    }
    if (dim == 6) {
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 6; j++) {
                M[j * 9 + 3 + i] = 0.5 * (M[j * 9 + 3 + i] + M[j * 9 + 6 + i]);
                M[(3 + i) * 9 + j] = 0.5 * (M[(3 + i) * 9 + j] + M[(6 + i) * 9 + j]);
            }
        Mat R(36);
        for (int i = 0; i < 6; i++)
            for (int j = 0; j < 6; j++) R[i * 6 + j] = M[i * 9 + j];
        _BC_M = R;
    } else {
        _BC_M = M;
    }
    _BC_MQ = dyad4(_BC_M, _BC_Q, dim);
    pushBC();
}

void LSSolver::pushBC() {
    if (!_ctx) return;
    // hand MQ and M:(QC0) to the device side; NaN reference (not yet computed) -> inactive
    Mat MQC0 = dyad4(_BC_M, _BC_QC0, _dim);
    bool finite = true;
    for (double x : _BC_MQ) finite = finite && std::isfinite(x);
    for (double x : MQC0) finite = finite && std::isfinite(x);
    if (!finite) check(fgb_set_bc(_ctx, nullptr, nullptr, _bc_relax));
    else check(fgb_set_bc(_ctx, _BC_MQ.data(), MQC0.data(), _bc_relax));
}

Vec LSSolver::calcBCMean(const Vec& E, const Vec& S) const {
    return E + _bc_relax * dyad4(_BC_M, S - dyad4(_BC_QC0, E));                             // fg:20244
}

Vec LSSolver::calcMeanStress() {
    syncEpsilon();
    Vec out(_dim);
    check(fgb_mean_pk1(_ctx, _epsilon, 1.0, out.data()));
    return out;
}
Vec LSSolver::calcMeanStrain() {
    syncEpsilon();
    Vec out(_dim);
    check(fgb_average(_ctx, _epsilon, out.data()));
    return out;
}
Vec LSSolver::calcMeanCauchyStress() {
    syncEpsilon();
    Vec out(9);
    check(fgb_mean_cauchy(_ctx, _epsilon, 1.0, out.data()));
    return out;
}
double LSSolver::calcMeanEnergy() {
    syncEpsilon();
    double w = 0;
    check(fgb_mean_energy(_ctx, _epsilon, &w));
    return w;
}

Vec EpsilonEE::mean() {
    Vec n(s->dim());
    if (fgb_component_dot(s->ctx(), s->epsilonField(), s->epsilonField(), n.data()) < 0) throw std::runtime_error(fgb_last_error(s->ctx()));
    for (double& x : n) x = std::sqrt(x);                                                   // component_norm fg:10127
    return fix_dim(n, s->dim());
}

ErrorEstimator* LSSolver::create_error_estimator(const std::string& name_) {
    const std::string name = name_.empty() ? _error_estimator : name_;
    if (name == "sigma") return new SigmaEE(this);
    if (name == "epsilon") return new EpsilonEE(this);
    if (name == "energy") return new EnergyEE(this);
    if (name == "residual") return new ResidualEE();
    if (name == "none") return new NoneEE();
    fail("Unknown error estimator '" + name + "'");
    return nullptr;
}

void LSSolver::calcRefMaterial() {
    // fg:22283-22313 + getRefMaterial fg:12153-12236
    NvtxRange nv("calc ref material");
    double lmin, lmax;
    check(fgb_ref_material(_ctx, _epsilon, _mode == "viscosity" ? 1 : 0, &lmin, &lmax));
    if (lmin < 0) lmin = 0;                                                                 // fg:12179-12218
    double mu_0 = (_method == "polarization") ? std::sqrt(lmin * lmax) : 0.5 * (lmin + lmax);
    _mu_0 = mu_0 * 0.5 * _ref_scale;
    setBCProjector(_BC_P);
}

double LSSolver::bc_error() {
    // fg:21129-21161
    Vec Emean = calcMeanStrain();
    Vec Smean = calcMeanStress();
    Vec P_Emean = dyad4(_BC_P, Emean);
    Vec Q_Smean = dyad4(_BC_Q, Smean);
    Vec PE = dyad4(_BC_P, _current_E);
    if (_dim == 9) PE = PE - dyad4(_BC_P, _Id);
    const double norm_E = norm_2(PE);
    const double err_F = norm_2(P_Emean - _current_E) / ((norm_E < _bc_tol) ? 1 : norm_E);
    const double norm_S = norm_2(_current_S);
    const double err_S = norm_2(Q_Smean - _current_S) / ((norm_S < _bc_tol) ? 1 : norm_S);
    return std::max(err_F, err_S);
}

bool LSSolver::converged(size_t& iter, double abs_err, double rel_err, bool check_bc) {
    // fg:21177-21244
    if (std::isnan(rel_err)) fail("NaN detected in solution. Aborting.");
    if (_cancel) fail("fibergen canceled");
    _residuals.push_back(rel_err);
    if (_cb && _cb(_cb_user)) return true;
    if (iter >= _maxiter) return true;
    if (rel_err <= _tol || abs_err <= _abs_tol) {
        double bc_err = 0;
        if (check_bc) bc_err = bc_error();
        if (bc_err <= _bc_tol) return true;
    }
    iter++;
    return false;
}

void LSSolver::cancel() { _cancel = true; }
void LSSolver::setConvergenceCallback(ConvergenceCallback cb, void* user) { _cb = cb; _cb_user = user; }

int LSSolver::field(int& slot) {
    if (slot < 0) {
        slot = fgb_field_alloc(_ctx);
        check(slot);
    }
    return slot;
}

bool LSSolver::run() {
    // fg:21247-21399
    _solve_time = 0;
    _residuals.clear();
    _cancel = false;
    _error.clear();
    _eps_stale = false;
    if (!_ctx) fail("solver not initialised");
    try {
        setBCProjector(_BC_P);
        const double eps = std::sqrt(EPS);
        if (plain_norm(dyad4(_BC_P, _S)) > eps * plain_norm(_S)) fail("Incompatible stress boundary condition specified");
        if (plain_norm(dyad4(_BC_Q, _E)) > eps * plain_norm(_E)) fail("Incompatible strain boundary condition specified");
        const auto t0 = std::chrono::steady_clock::now();
        if (_mode == "hyperelasticity") check(fgb_set_constant(_ctx, _epsilon, _Id.data()));
        else {
            Vec z(_dim, 0.0);
            check(fgb_set_constant(_ctx, _epsilon, z.data()));
        }
        const bool ret = runLoadsteppingSolver(_E, _S);
        check(fgb_synchronize(_ctx));
        _solve_time += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return ret;
    } catch (const std::exception& e) {
        if (_error.empty()) _error = e.what();
        return true;
    }
}

// LU with partial pivoting (what lapack::gesv does, fg:21486): X = A^-1, A is n x n row-major
static bool invert_gesv(Mat A, int n, Mat& X) {
    X.assign(n * n, 0.0);
    for (int i = 0; i < n; i++) X[i * n + i] = 1.0;
    for (int k = 0; k < n; k++) {
        int piv = k;
        for (int i = k + 1; i < n; i++)
            if (std::fabs(A[i * n + k]) > std::fabs(A[piv * n + k])) piv = i;
        if (A[piv * n + k] == 0.0) return false;
        if (piv != k)
            for (int j = 0; j < n; j++) { std::swap(A[k * n + j], A[piv * n + j]); std::swap(X[k * n + j], X[piv * n + j]); }
        for (int i = k + 1; i < n; i++) {
            const double l = A[i * n + k] / A[k * n + k];
            for (int j = k; j < n; j++) A[i * n + j] -= l * A[k * n + j];
            for (int j = 0; j < n; j++) X[i * n + j] -= l * X[k * n + j];
        }
    }
    for (int k = n - 1; k >= 0; k--)
        for (int j = 0; j < n; j++) {
            double v = X[k * n + j];
            for (int i = k + 1; i < n; i++) v -= A[k * n + i] * X[i * n + j];
            X[k * n + j] = v / A[k * n + k];
        }
    return true;
}

void LSSolver::extrapolateLoadstep(const std::vector<std::pair<double, int>>& last, double t) {
    // extrapolateLoadstepPolynomial fg:21468-21513: Vandermonde system through the stored load steps, evaluated at t
    const int n = (int)last.size();
    Mat V(n * n), Vinv;
    Vec tpowers(n);
    for (int i = 0; i < n; i++) {
        for (int j = 0; j < n; j++) V[i * n + j] = std::pow(last[i].first, j);
        tpowers[i] = std::pow(t, i);
    }
    if (!invert_gesv(V, n, Vinv)) fail("Error inverting Vandermonde matrix");
    std::vector<int> ids(n);
    for (int i = 0; i < n; i++) ids[i] = last[i].second;
    check(fgb_extrapolate_polynomial(_ctx, n, ids.data(), Vinv.data(), tpowers.data(), _epsilon));
}

bool LSSolver::runLoadsteppingSolver(const Vec& Emax, const Vec& Smax) {
    // fg:21584-21686
    const size_t first = (_first_loadstep >= 0) ? (size_t)_first_loadstep : ((_loadsteps.size() > 2) ? 0 : 1);
    std::vector<std::pair<double, int>> last;          // (load parameter, field holding that step's solution)
    struct Release {
        LSSolver* s;
        std::vector<std::pair<double, int>>* l;
        ~Release() { for (auto& e : *l) fgb_field_free(s->_ctx, e.second); }
    } release{this, &last};
    for (size_t istep = first; istep < _loadsteps.size(); istep++) {
        const double t = _loadsteps[istep];
        Vec E = t * Emax;
        Vec S = t * Smax;
        if (_mode == "hyperelasticity") E = E + (1 - t) * dyad4(_BC_P, _Id);
        if (_loadstep_extrapolation_order > 0 && istep > first) {                          // fg:21634-21650
            while (last.size() > _loadstep_extrapolation_order) {
                check(fgb_field_free(_ctx, last.front().second));
                last.erase(last.begin());
            }
            const int keep = fgb_field_alloc(_ctx);
            check(keep);
            check(fgb_copy(_ctx, _epsilon, keep));
            last.push_back(std::make_pair(_loadsteps[istep - 1], keep));
            if (last.size() >= 2) extrapolateLoadstep(last, t);
        }
        runSolver(E, S);
    }
    return false;
}

void LSSolver::runSolver(const Vec& E, const Vec& S) {
    _current_E = E;
    _current_S = S;
    if (_method == "basic") runBasic(E, S);
    else if (_method == "polarization") runPolarization(E, S);
    else if (_method == "cg") {
        if (_mode == "hyperelasticity") runCGHyper(E, S);
        else runCGElasticity(E, S);
    } else fail("Unknown solver method '" + _method + "'");
}

void LSSolver::runBasic(const Vec& E0, const Vec& S0) {
    // fg:21716-21805
    NvtxRange nv("running solver (basic)");
    size_t iter = 1;
    std::unique_ptr<ErrorEstimator> ee(create_error_estimator());
    bool update_ref = (_update_ref != "never");
    Vec E = calcBCMean(E0, S0);
    for (;;) {
        if (update_ref) {
            calcRefMaterial();
            E = calcBCMean(E0, S0);
            update_ref = false;
        }
        check(fgb_basic_step(_ctx, _epsilon, _epsilon, E.data(), _mu_0, _lambda_0));       // in place, fg:21786
        ee->update();
        if (converged(iter, ee->abs_error(), ee->rel_error())) break;
    }
}

void LSSolver::runPolarization(const Vec& E0, const Vec& S0) {
    // fg:21808-21851
    NvtxRange nv("running solver (polarization)");
    size_t iter = 1;
    std::unique_ptr<ErrorEstimator> ee(create_error_estimator());
    if (_update_ref != "never") calcRefMaterial();
    Vec E = calcBCMean(E0, S0);
    Vec P0 = (4 * _mu_0) * E;
    check(fgb_set_constant(_ctx, _epsilon, P0.data()));
    for (;;) {
        P0 = (4 * _mu_0) * E;
        check(fgb_polarization_step(_ctx, _epsilon, _epsilon, P0.data(), _mu_0, _lambda_0));
        ee->update();
        if (converged(iter, ee->abs_error(), ee->rel_error(), false)) break;
    }
    check(fgb_calc_polarization(_ctx, _epsilon, _epsilon, _mu_0, 1));
}

void LSSolver::runCGElasticity(const Vec& E0, const Vec& S0) {
    // fg:23153-23247
    NvtxRange nv("runCGElasticity");
    if (_update_ref != "never") calcRefMaterial();
    Vec E = calcBCMean(E0, S0);
    std::unique_ptr<ErrorEstimator> ee(create_error_estimator());
    const int r = field(_f1);
    int p = field(_f2), p2 = field(_f4);                                                    // direction vector, ping-pong
    // the operator result w is only consumed by the residual update: on the fused path it is never written to memory
    // (fgb200.h: FGB_W_IMPLICIT); the exact-residual variant and every other configuration keep the explicit field
    // (the fused heat sweeps assume lambda_0 = 0, which holds unless a <ref> material sets it, fg:15186-15194)
    const int w = (_cg_reinit <= 0 && fgb_cg_implicit_w_supported(_ctx) && (_dim != 3 || _lambda_0 == 0.0)) ? FGB_W_IMPLICIT : field(_f3);
    check(fgb_set_constant(_ctx, _epsilon, E.data()));
    check(fgb_cg_step(_ctx, -1, -1, 0.0, _epsilon, _epsilon, r, _mu_0, _lambda_0, nullptr)); // krylovOperator(epsilon -> r)
    check(fgb_adjust_residual(_ctx, r, E.data(), _epsilon));
    double gamma;
    check(fgb_inner(_ctx, r, r, -1, &gamma));
    gamma += SMALL;
    const double gamma0 = gamma;
    Vec zero(_dim, 0.0);
    check(fgb_set_constant(_ctx, p, zero.data()));
    if (_cg_reinit <= 0 && _pipelined_cg) {
        runCGElasticityPipelined(E, r, p, p2, w, gamma, ee.get());
        return;
    }
    double beta = 0.0;                                                                       // p = r  ==  r + 0*p
    size_t iter = 0;
    for (;;) {
        double alpha;
        // p2 = r + beta*p ; w = MinusB(p2) ; alpha = <p2, p2 - w>   (fg:23245, fg:23209, fg:23211)
        check(fgb_cg_step(_ctx, -1, r, beta, p, p2, w, _mu_0, _lambda_0, &alpha));
        std::swap(p, p2);
        alpha += SMALL;
        alpha = gamma / alpha;
        if (_cg_reinit > 0) {
            // exact residual recomputation path keeps the unfused order of the reference (fg:23221-23235)
            check(fgb_xpay(_ctx, _epsilon, _epsilon, alpha, p));
            ee->update_cg(gamma, gamma0);
            if (converged(iter, ee->abs_error(), ee->rel_error())) break;
            if ((iter % _cg_reinit) == 0) {
                check(fgb_cg_step(_ctx, -1, -1, 0.0, _epsilon, _epsilon, r, _mu_0, _lambda_0, nullptr));
                check(fgb_adjust_residual(_ctx, r, E.data(), _epsilon));
            } else {
                check(fgb_xpaymz(_ctx, r, r, -alpha, p, w));
            }
            double delta;
            check(fgb_inner(_ctx, r, r, -1, &delta));
            delta += SMALL;
            beta = delta / gamma;
            gamma = delta;
            continue;
        }
        // fused sweep: epsilon += alpha*p ; r -= alpha*(p - w) ; delta = <r,r>.  The residual update is independent of
        // the convergence decision taken in between in the reference, so doing it first changes nothing observable.
        double delta;
        check(fgb_cg_update(_ctx, _epsilon, r, p, w, alpha, &delta));
        ee->update_cg(gamma, gamma0);
        if (converged(iter, ee->abs_error(), ee->rel_error())) break;
        delta += SMALL;
        beta = delta / gamma;
        gamma = delta;
    }
}

void LSSolver::runCGElasticityPipelined(const Vec& E, int r, int p, int p2, int w, double gamma, ErrorEstimator* ee) {
    // The loop of fg:23206-23246 with gamma, beta, alpha on the device (fgb_cgdev_*).  The stop test of iteration k needs gamma_k
    // only (fg:23223-23226), which the update of iteration k-1 produced: the operator application of iteration k+1 is enqueued
    // before the host waits for delta_k, so the device never idles and an iteration costs one (hidden) host synchronisation.
    (void)E;
    NvtxRange nv("CG loop");
    const double gamma0 = gamma;
    check(fgb_cgdev_begin(_ctx, gamma));
    check(fgb_cgdev_step(_ctx, -1, r, p, p2, w, _mu_0, _lambda_0));
    std::swap(p, p2);
    size_t iter = 0;
    for (size_t k = 0;; k++) {
        const int slot = (int)(k % 8);
        check(fgb_cgdev_update(_ctx, _epsilon, r, p, w, slot));                            // epsilon += alpha*p ; r -= alpha*(p - w)
        ee->update_cg(gamma, gamma0);
        if (converged(iter, ee->abs_error(), ee->rel_error())) break;
        check(fgb_cgdev_step(_ctx, -1, r, p, p2, w, _mu_0, _lambda_0));                    // iteration k+1, before delta_k is known
        std::swap(p, p2);
        double s[4];
        check(fgb_cgdev_wait(_ctx, slot, s));
        gamma = s[3] + SMALL;
    }
}

void LSSolver::syncEpsilon() {
    // inner Newton-CG iterations keep the current iterate as F + newton_relax*X (fg:23049) and form it when somebody looks
    if (!_eps_stale) return;
    _eps_stale = false;
    check(fgb_xpay(_ctx, _epsilon, _eps_F, _newton_relax, _eps_X));
}

void LSSolver::runCGHyper(const Vec& E0, const Vec& S0) {
    // fg:22699-23130
    NvtxRange nv("runCGHyper");
    const int F = field(_f1), X = field(_f2), R = field(_f3), Q = field(_f4);
    Vec dE = E0 - dyad4(_BC_P, calcMeanStrain());
    check(fgb_add_constant(_ctx, _epsilon, dE.data()));
    std::unique_ptr<ErrorEstimator> ee_outer(create_error_estimator(_outer_error_estimator));
    size_t iter_outer = 0;
    double gamma0 = -1;
    Vec zero(_dim, 0.0);
    // the current iterate must be in _epsilon whenever an estimator or a user callback can look at it
    const bool observed = _cb != nullptr || (_error_estimator != "residual" && _error_estimator != "none");
    for (;;) {
        if (gamma0 < 0 || _update_ref == "always") calcRefMaterial();
        check(fgb_copy(_ctx, _epsilon, F));
        check(fgb_calc_stress(_ctx, F, X, 0.0, 0.0, 1.0));                                  // X = P(F)
        Vec X0 = dyad4(_BC_M, S0);
        check(fgb_gamma(_ctx, X, X0.data(), _mu_0, _lambda_0, -1.0, 0.0));                  // X = -Gamma0 X, <X> = X0
        // everything of the tangent that depends on F only, once per Newton iteration; > 0: fused Neo-Hooke sweeps, W stays implicit
        const int fused = _pipelined_cg ? fgb_cg_tangent_prepare(_ctx, F, _mu_0, _lambda_0) : 0;
        check(fused);
        const int W = fused ? FGB_W_IMPLICIT : field(_f5);
        check(fgb_cg_apply(_ctx, F, X, R, _mu_0, _lambda_0, nullptr));                      // R = ApplyOperator(F, X)
        check(fgb_copy(_ctx, R, Q));
        double gamma;
        check(fgb_inner(_ctx, R, R, -1, &gamma));
        gamma += SMALL;
        if (gamma0 < 0) gamma0 = gamma;
        std::unique_ptr<ErrorEstimator> ee(create_error_estimator());
        size_t iter = 0;
        if (_pipelined_cg) {
            // inner CG with gamma, beta, alpha on the device (see runCGElasticityPipelined): the operator application of iteration
            // k+1 is enqueued before the host waits for delta_k
            _eps_F = F; _eps_X = X;
            check(fgb_cgdev_begin(_ctx, gamma));
            check(fgb_cgdev_step(_ctx, F, R, Q, Q, W, _mu_0, _lambda_0));                   // Q = R + 0*Q
            for (size_t k = 0;; k++) {
                const int slot = (int)(k % 8);
                check(fgb_cgdev_update(_ctx, X, R, Q, W, slot));                            // X += alpha*Q ; R -= alpha*(Q - W)
                _eps_stale = true;                                                           // next F = current F + dF (fg:23049)
                if (observed) syncEpsilon();
                ee->update_cg(gamma, gamma0);
                const bool stop = converged(iter, ee->abs_error(), ee->rel_error(), false);
                if (!stop) check(fgb_cgdev_step(_ctx, F, R, Q, Q, W, _mu_0, _lambda_0));
                double s[4];
                check(fgb_cgdev_wait(_ctx, slot, s));                                        // also raises law domain errors (fg:10293)
                if (s[1] + SMALL <= 0) {
                    std::ostringstream o;
                    o << "indefinite operator (alpha=" << (s[1] + SMALL) << ") canceling CG!";
                    fail(o.str());
                }
                if (stop) break;
                gamma = s[3] + SMALL;
            }
            syncEpsilon();
        } else {
            for (;;) {
                double alpha;
                check(fgb_cg_apply(_ctx, F, Q, W, _mu_0, _lambda_0, &alpha));
                alpha += SMALL;
                if (alpha <= 0) {
                    std::ostringstream o;
                    o << "indefinite operator (alpha=" << alpha << ") canceling CG!";
                    fail(o.str());
                }
                alpha = gamma / alpha;
                double delta;
                check(fgb_cg_update(_ctx, X, R, Q, W, alpha, &delta));                      // X += alpha*Q ; R -= alpha*(Q-W)
                check(fgb_xpay(_ctx, _epsilon, F, _newton_relax, X));                       // next F = current F + dF (fg:23049)
                check(fgb_check_numeric(_ctx));
                ee->update_cg(gamma, gamma0);
                if (converged(iter, ee->abs_error(), ee->rel_error(), false)) break;
                delta += SMALL;
                const double beta = delta / gamma;
                gamma = delta;
                check(fgb_cg_direction(_ctx, Q, R, beta));
            }
        }
        ee_outer->update();
        if (converged(iter_outer, ee_outer->abs_error(), ee_outer->rel_error())) break;
    }
}

Mat LSSolver::calcEffectiveProperties() {
    // fg:26030-26160: unit load cases, Ceff = S*E^-1 with E = identity, last 3 columns halved for dim 6
    const int d = _dim;
    Mat C(d * d, 0.0);
    for (int i = 0; i < d; i++) {
        Vec E(d, 0.0);
        E[i] = 1.0;
        _E = E;
        if (run()) throw std::runtime_error(_error);
        Vec S = calcMeanStress();
        for (int r = 0; r < d; r++) C[r * d + i] = S[r];
    }
    if (d == 6)
        for (int r = 0; r < 6; r++)
            for (int c = 3; c < 6; c++) C[r * 6 + c] *= 0.5;
    return C;
}

int LSSolver::fieldComponents(const std::string& name) const {
    if (name == "epsilon" || name == "sigma") return _dim;
    if (name == "u") return _dim == 3 ? 1 : 3;
    if (name == "p") return _dim >= 6 ? 1 : -1;
    return -1;
}

void LSSolver::getField(const std::string& name, double* const* comps) {
    // get_raw_field fg:15396-15557: derived fields are evaluated on the device from the converged strain field
    syncEpsilon();
    if (name == "epsilon") check(fgb_field_download(_ctx, _epsilon, comps));
    else if (name == "sigma") {                                                             // calcStress(epsilon, sigma) fg:15500
        const int t = field(_f5);
        check(fgb_calc_stress(_ctx, _epsilon, t, 0.0, 0.0, 1.0));
        check(fgb_field_download(_ctx, t, comps));
    } else if (name == "u") {                                                               // fg:15517-15557
        check(fgb_calc_displacement(_ctx, _epsilon, field(_f5), _mu_0, _lambda_0));
        check(fgb_u_download(_ctx, comps, fieldComponents(name)));
    } else if (name == "p" && _dim >= 6) {                                                  // fg:15559-15573
        check(fgb_calc_pressure(_ctx, _epsilon, field(_f5), _mu_0, _lambda_0));
        check(fgb_u_download(_ctx, comps, 1));
    } else fail("Unknown field '" + name + "'");
}

int LSSolver::localNx() const { return fgb_local_nx(_ctx); }
size_t LSSolver::planeElems() const { return fgb_plane_elems(_ctx); }
unsigned long long LSSolver::launches() const { return fgb_launch_count(_ctx); }

}  // namespace fgb

// ---- flat C view -----------------------------------------------------------------------------------------------
struct fgls_solver {
    fgb::LSSolver* s;
    std::string err;
};

#define FGLS_TRY(body)                              \
    if (!h) return FGB_EINVAL;                      \
    try {                                           \
        body;                                       \
        return FGB_OK;                              \
    } catch (const std::exception& e) {             \
        h->err = e.what();                          \
        return FGB_EINVAL;                          \
    }

extern "C" {

int fgls_create(fgls_solver** out, int nx, int ny, int nz, double dx, double dy, double dz, int rank, int nranks, int device) {
    if (!out) return FGB_EINVAL;
    fgls_solver* h = new fgls_solver();
    h->s = new fgb::LSSolver(nx, ny, nz, dx, dy, dz, rank, nranks, device);
    *out = h;
    return FGB_OK;
}
void fgls_destroy(fgls_solver* h) {
    if (!h) return;
    delete h->s;
    delete h;
}
const char* fgls_last_error(const fgls_solver* h) { return h ? h->err.c_str() : "null solver"; }
int fgls_set(fgls_solver* h, const char* k, const char* v) { FGLS_TRY(h->s->set(k, v)) }
int fgls_add_material(fgls_solver* h, const char* name, const char* law, const double* p, int n) { FGLS_TRY(h->s->addMaterial(name, law, p, n)) }
int fgls_set_reference(fgls_solver* h, double mu, double lambda) { FGLS_TRY(h->s->setReference(mu, lambda)) }
int fgls_init(fgls_solver* h) { FGLS_TRY(h->s->init()) }
int fgls_init_comm(fgls_solver* h, const void* id) { FGLS_TRY(h->s->initComm(id)) }
int fgls_set_phase(fgls_solver* h, int m, const double* phi) { FGLS_TRY(h->s->setPhase(m, phi)) }
int fgls_init_phase_capsules(fgls_solver* h, int nfib, const fgb_capsule* fibers, int matrix_mat, int normals, int orientation) {
    FGLS_TRY(h->s->initPhase(nfib, fibers, matrix_mat, normals != 0, orientation != 0))
}
int fgls_get_phase(fgls_solver* h, int m, double* phi) { FGLS_TRY(h->s->getPhase(m, phi)) }
int fgls_set_normals(fgls_solver* h, const double* const* c) { FGLS_TRY(h->s->setNormals(c)) }
int fgls_set_orientation(fgls_solver* h, const double* const* c) { FGLS_TRY(h->s->setOrientation(c)) }
int fgls_set_strain(fgls_solver* h, const double* E) { FGLS_TRY(h->s->setStrain(fgb::Vec(E, E + h->s->dim()))) }
int fgls_set_stress(fgls_solver* h, const double* S) { FGLS_TRY(h->s->setStress(fgb::Vec(S, S + h->s->dim()))) }
int fgls_set_bc_projector(fgls_solver* h, const double* P) { FGLS_TRY(h->s->setBCProjector(fgb::Mat(P, P + h->s->dim() * h->s->dim()))) }
int fgls_set_callback(fgls_solver* h, fgls_callback cb, void* user) {
    FGLS_TRY(h->s->setConvergenceCallback(reinterpret_cast<fgb::LSSolver::ConvergenceCallback>(reinterpret_cast<void*>(cb)), user))
}
int fgls_run(fgls_solver* h) {
    if (!h) return FGB_EINVAL;
    try {
        if (h->s->run()) {
            h->err = h->s->lastError();
            return FGB_ENUMERIC;
        }
        return FGB_OK;
    } catch (const std::exception& e) {
        h->err = e.what();
        return FGB_EINVAL;
    }
}
int fgls_cancel(fgls_solver* h) { FGLS_TRY(h->s->cancel()) }
int fgls_num_residuals(const fgls_solver* h) { return h ? (int)h->s->getResiduals().size() : 0; }
int fgls_get_residuals(const fgls_solver* h, double* out, int n) {
    if (!h) return FGB_EINVAL;
    const std::vector<double>& r = h->s->getResiduals();
    for (int i = 0; i < n && i < (int)r.size(); i++) out[i] = r[i];
    return FGB_OK;
}
int fgls_mean_stress(fgls_solver* h, double* out) { FGLS_TRY(fgb::Vec v = h->s->calcMeanStress(); std::copy(v.begin(), v.end(), out)) }
int fgls_mean_strain(fgls_solver* h, double* out) { FGLS_TRY(fgb::Vec v = h->s->calcMeanStrain(); std::copy(v.begin(), v.end(), out)) }
int fgls_mean_energy(fgls_solver* h, double* out) { FGLS_TRY(*out = h->s->calcMeanEnergy()) }
int fgls_mean_cauchy_stress(fgls_solver* h, double* out) { FGLS_TRY(fgb::Vec v = h->s->calcMeanCauchyStress(); std::copy(v.begin(), v.end(), out)) }
int fgls_field_components(fgls_solver* h, const char* name) { return h ? h->s->fieldComponents(name) : -1; }
int fgls_effective_properties(fgls_solver* h, double* C) { FGLS_TRY(fgb::Mat m = h->s->calcEffectiveProperties(); std::copy(m.begin(), m.end(), C)) }
int fgls_get_field(fgls_solver* h, const char* name, double* const* comps) { FGLS_TRY(h->s->getField(name, comps)) }
int fgls_ref_material(fgls_solver* h, double* mu0, double* lambda0) { FGLS_TRY(*mu0 = h->s->mu0(); *lambda0 = h->s->lambda0()) }
int fgls_calc_ref_material(fgls_solver* h) { FGLS_TRY(h->s->calcRefMaterial()) }
int fgls_bc_matrices(fgls_solver* h, double* M, double* MQ) {
    FGLS_TRY(std::copy(h->s->BC_M().begin(), h->s->BC_M().end(), M); std::copy(h->s->BC_MQ().begin(), h->s->BC_MQ().end(), MQ))
}
int fgls_dim(const fgls_solver* h) { return h ? h->s->dim() : 0; }
int fgls_local_nx(const fgls_solver* h) { return h ? h->s->localNx() : 0; }
double fgls_solve_time(const fgls_solver* h) { return h ? h->s->getSolveTime() : 0; }
unsigned long long fgls_launches(const fgls_solver* h) { return h ? h->s->launches() : 0; }
fgb_ctx* fgls_ctx(fgls_solver* h) { return h ? h->s->ctx() : nullptr; }

}  // extern "C"
