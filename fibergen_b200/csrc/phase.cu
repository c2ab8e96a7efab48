// Phase initialisation on the device: LSSolver::initPhi fg:17489-17581 for capsule fibres.
//   per voxel centre: FiberCluster::closestFibers fg:3336 (all fibres of the material with signed distance <= r_voxel),
//   integratePhiVoxel fg:16622-16752 (adaptive subdivision, half-space cuts at the leaves), halfspace_box_cut_volume fg:1385-1575,
//   CapsuleFiber::distanceTo / distanceGrad / curvature fg:5279-5333, fg:5470, normalizePhi fg:17588-17646,
//   normals / orientation of the closest fibre as FiberGenerator::sampleZYSlice writes them (fg:6885-6924).
// One thread per voxel.  The closest-fibre query walks a uniform cell list built on the host (the reference walks a cluster tree,
// fg:3157-3647; both are only filters in front of the exact distance test).  The recursion of integratePhiVoxel becomes an explicit
// stack; sub-lists are bit masks over the voxel's fibre list.  This file is compiled without FMA contraction so that the geometric
// predicates see the same products and sums as the reference's scalar code.
#include "fgb_internal.h"
#include <algorithm>
#include <cfloat>
#include <cmath>

#define PH_MAXINFO 8
#define PH_MAXDEPTH 20

struct CapsuleDev {
    double c1[3], a[3], r[3];
    double R, L;
    int mat;
};

struct PhaseArgs {
    GridDev g;
    double x0[3], dv[3];          // cell origin, voxel size
    int nmat, matrix_mat, smooth_levels;
    double smooth_tol;
    int cs, ncx, ncy, ncz, cx0;   // cell list: cells of cs^3 voxels, ncx x ncy x ncz cells, first local cell row covers global i = cx0*cs
    const int* cell_start;
    const int* cell_fibs;
    const CapsuleDev* fib;
    double* phi[FGB_MAX_PHASES];
    double* normals;              // 3 planes or null
    double* orient;               // 3 planes or null
    double lvlpow[PH_MAXDEPTH + 2];   // pow(2^-l, 2/3) from the host's libm (fg:16666)
    int* flag;
};

__device__ __forceinline__ double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ double norm3(const double* a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

// CapsuleFiber::distanceTo fg:5298-5333
__device__ double capsule_distance(const CapsuleDev& f, const double* p, double* x) {
    double pc[3], q[3];
#pragma unroll
    for (int i = 0; i < 3; i++) pc[i] = p[i] - f.c1[i];
    double t = dot3(pc, f.a);
    t = fmin(fmax(0.0, t), f.L);
#pragma unroll
    for (int i = 0; i < 3; i++) x[i] = f.c1[i] + t * f.a[i];
#pragma unroll
    for (int i = 0; i < 3; i++) q[i] = p[i] - x[i];
    const double d = norm3(q);
    if (d < DBL_EPSILON * f.R) {
#pragma unroll
        for (int i = 0; i < 3; i++) x[i] += f.r[i];
    } else {
#pragma unroll
        for (int i = 0; i < 3; i++) x[i] += q[i] * (f.R / d);
    }
    return d - f.R;
}

// CapsuleFiber::distanceGrad fg:5279-5296
__device__ void capsule_grad(const CapsuleDev& f, const double* p, double* g) {
    double pc[3];
#pragma unroll
    for (int i = 0; i < 3; i++) pc[i] = p[i] - f.c1[i];
    double t = dot3(pc, f.a);
    t = fmin(fmax(0.0, t), f.L);
#pragma unroll
    for (int i = 0; i < 3; i++) g[i] = p[i] - f.c1[i] - t * f.a[i];
    const double n = norm3(g);
    if (n < sqrt(DBL_EPSILON)) {
#pragma unroll
        for (int i = 0; i < 3; i++) g[i] = ((t < 0.5 * f.L) ? -1 : 1) * f.a[i];
    } else {
#pragma unroll
        for (int i = 0; i < 3; i++) g[i] /= n;
    }
}

__constant__ int c_edges[12][2] = {{0, 1}, {2, 4}, {3, 6}, {5, 7}, {0, 2}, {1, 4}, {3, 5}, {6, 7}, {0, 3}, {1, 6}, {2, 5}, {4, 7}};
__constant__ int c_faces[6][4] = {{8, 6, -10, -4}, {9, 7, -11, -5}, {0, 9, -2, -8}, {1, 11, -3, -10}, {0, 5, -1, -4}, {2, 7, -3, -6}};
__constant__ int c_crossp[3][2] = {{1, 2}, {2, 0}, {0, 1}};

// halfspace_box_cut_volume fg:1385-1575 (Gauss divergence theorem over the six faces)
__device__ double halfspace_box_cut_volume(const double* x, const double* n, const double* x0, double dx, double dy, double dz) {
    double v[8][3], dist[6], xi[3], pts[5][3], V = 0;
    int iedge[12], nint = 0, any = -1, num_inside = 0;
    unsigned inside = 0;
    for (int i = 0; i < 8; i++) { v[i][0] = x0[0]; v[i][1] = x0[1]; v[i][2] = x0[2]; }
    v[1][0] += dx;
    v[2][1] += dy;
    v[3][2] += dz;
    v[4][0] = v[1][0]; v[4][1] = v[1][1] + dy; v[4][2] = v[1][2];
    v[5][0] = v[2][0]; v[5][1] = v[2][1]; v[5][2] = v[2][2] + dz;
    v[6][0] = v[3][0] + dx; v[6][1] = v[3][1]; v[6][2] = v[3][2];
    v[7][0] = v[6][0]; v[7][1] = v[6][1] + dy; v[7][2] = v[6][2];
    for (int i = 0; i < 8; i++) {
        const double q[3] = {v[i][0] - x[0], v[i][1] - x[1], v[i][2] - x[2]};
        if (dot3(q, n) < 0) { inside |= 1u << i; num_inside++; }
    }
    for (int i = 0; i < 12; i++) {
        const int e0 = c_edges[i][0], e1 = c_edges[i][1];
        if (((inside >> e0) & 1u) + ((inside >> e1) & 1u) == 1) {
            const double q[3] = {x[0] - v[e0][0], x[1] - v[e0][1], x[2] - v[e0][2]};
            dist[nint] = dot3(q, n) / n[i / 4];
            iedge[i] = nint;
            any = i;
            nint++;
        } else iedge[i] = -1;
    }
    if (nint == 0) return (inside & 1u) ? (dx * dy * dz) : 0;
    {
        const int e0 = c_edges[any][0];
        xi[0] = v[e0][0]; xi[1] = v[e0][1]; xi[2] = v[e0][2];
        xi[any / 4] += dist[iedge[any]];
    }
    const unsigned flip = (num_inside > 4) ? 1u : 0u;
    for (int f = 0; f < 6; f++) {
        const int ni = f >> 1;
        int np = 0;
        bool brk = false;
        for (int i = 0; i < 4 && !brk; i++) {
            int e = c_faces[f][i], i1 = 0, i2 = 1;
            if (e < 0) { e = -e; i1 = 1; i2 = 0; }
            const int va = c_edges[e][i1], vb = c_edges[e][i2];
            if (np == 0 && (((inside >> va) & 1u) ^ flip)) {
                pts[np][0] = v[va][0]; pts[np][1] = v[va][1]; pts[np][2] = v[va][2];
                if (pts[0][ni] == xi[ni]) { brk = true; break; }
                np++;
            }
            if (iedge[e] >= 0) {
                const int e0 = c_edges[e][0];
                pts[np][0] = v[e0][0]; pts[np][1] = v[e0][1]; pts[np][2] = v[e0][2];
                pts[np][e / 4] += dist[iedge[e]];
                if (np == 0 && pts[0][ni] == xi[ni]) { brk = true; break; }
                np++;
            }
            if (i < 3 && (((inside >> vb) & 1u) ^ flip)) {
                pts[np][0] = v[vb][0]; pts[np][1] = v[vb][1]; pts[np][2] = v[vb][2];
                if (np == 0 && pts[0][ni] == xi[ni]) { brk = true; break; }
                np++;
            }
        }
        if (np < 3) continue;
        const int i1 = c_crossp[ni][0], i2 = c_crossp[ni][1];
        double area = 0;
        for (int i = 2; i < np; i++)
            area += fabs((pts[i - 1][i1] - pts[0][i1]) * (pts[i][i2] - pts[0][i2]) - (pts[i - 1][i2] - pts[0][i2]) * (pts[i][i1] - pts[0][i1]));
        const double d = pts[0][ni] - xi[ni];
        V += ((f & 1) ? 1 : -1) * d * area;
    }
    V *= (1.0 / 6.0);
    if (flip) V = dx * dy * dz - V;
    return V;
}

struct Frame {
    double p[3];
    double V;
    unsigned mask;
    int child, levels;
};

// integratePhiVoxel fg:16622-16752 for the fibres `list[0..n)` (indices into A.fib) of one voxel with centre p
__device__ double integrate_phi_voxel(const PhaseArgs& A, const int* list, int nlist, const double* p, double r_voxel0) {
    Frame st[PH_MAXDEPTH];
    int depth = 0;
    st[0].p[0] = p[0]; st[0].p[1] = p[1]; st[0].p[2] = p[2];
    st[0].V = 0;
    st[0].mask = (1u << nlist) - 1u;
    st[0].child = -1;
    st[0].levels = A.smooth_levels;
    double result = 0;
    while (depth >= 0) {
        Frame& F = st[depth];
        const double sc = ldexp(1.0, -depth);
        const double dx = A.dv[0] * sc, dy = A.dv[1] * sc, dz = A.dv[2] * sc;
        double ret = 0;
        bool pop = false;
        if (F.child < 0) {
            // entry of integratePhiVoxel: distances of the list at the centre (they are what the caller stored in info_list)
            double d[PH_MAXINFO], x[PH_MAXINFO][3];
            int i_min = -1;
            for (int i = 0; i < nlist; i++) {
                if (!((F.mask >> i) & 1u)) continue;
                d[i] = capsule_distance(A.fib[list[i]], F.p, x[i]);
                if (i_min < 0 || d[i] < d[i_min]) i_min = i;
            }
            const double r_voxel = 0.5 * sqrt(dx * dx + dy * dy + dz * dz);
            const double V_max = dx * dy * dz;
            int levels = F.levels;
            if (i_min < 0) { ret = 0; pop = true; }
            else if (fabs(d[i_min]) >= r_voxel) { ret = (d[i_min] < 0) ? V_max : 0; pop = true; }
            else {
                if (levels < 0) {
                    const double K = 1 / A.fib[list[i_min]].R;
                    const double Kd = r_voxel * K;
                    double err;
                    if (Kd > 1) err = 1;
                    else err = Kd * Kd * A.lvlpow[depth];
                    if (err < A.smooth_tol) levels = 0;
                }
                if (levels == 0 || depth + 1 >= PH_MAXDEPTH) {
                    const double x0[3] = {F.p[0] - 0.5 * dx, F.p[1] - 0.5 * dy, F.p[2] - 0.5 * dz};
                    double V = 0;
                    for (int i = 0; i < nlist; i++) {
                        if (!((F.mask >> i) & 1u)) continue;
                        double n[3];
                        capsule_grad(A.fib[list[i]], x[i], n);
                        V += halfspace_box_cut_volume(x[i], n, x0, dx, dy, dz);
                    }
                    ret = fmin(V, V_max);
                    pop = true;
                } else {
                    F.levels = levels - 1;
                    F.child = 0;
                    F.V = 0;
                }
            }
        }
        if (!pop) {
            if (F.child == 8) {
                ret = fmin(F.V, dx * dy * dz);
                pop = true;
            } else {
                const int c = F.child++;
                const int ci = c >> 2, cj = (c >> 1) & 1, ck = c & 1;
                const double hx = 0.5 * dx, hy = 0.5 * dy, hz = 0.5 * dz;
                const double x0[3] = {F.p[0] - 0.5 * dx, F.p[1] - 0.5 * dy, F.p[2] - 0.5 * dz};
                const double ps[3] = {x0[0] + (ci + 0.5) * hx, x0[1] + (cj + 0.5) * hy, x0[2] + (ck + 0.5) * hz};
                const double r_child = 0.5 * (0.5 * sqrt(dx * dx + dy * dy + dz * dz));
                unsigned sub = 0;
                for (int i = 0; i < nlist; i++) {
                    if (!((F.mask >> i) & 1u)) continue;
                    double xx[3];
                    const double d = capsule_distance(A.fib[list[i]], ps, xx);
                    if (fabs(d) >= r_child) {
                        if (d < 0) {
                            F.V += hx * hy * hz;
                            sub = 0;
                            break;
                        }
                        continue;
                    }
                    sub |= 1u << i;
                }
                if (sub) {
                    Frame& N = st[depth + 1];
                    N.p[0] = ps[0]; N.p[1] = ps[1]; N.p[2] = ps[2];
                    N.V = 0;
                    N.mask = sub;
                    N.child = -1;
                    N.levels = F.levels;
                    depth++;
                }
                continue;
            }
        }
        // return `ret` to the caller frame
        depth--;
        if (depth >= 0) st[depth].V += ret;
        else result = ret;
    }
    return result;
}

__global__ void __launch_bounds__(128) k_init_phi(PhaseArgs A) {
    const GridDev& g = A.g;
    const unsigned nvox = (unsigned)g.lnx * (unsigned)g.ny * (unsigned)g.nz;
    const double r_voxel = 0.5 * sqrt(A.dv[0] * A.dv[0] + A.dv[1] * A.dv[1] + A.dv[2] * A.dv[2]);
    const double V_voxel = A.dv[0] * A.dv[1] * A.dv[2];
    for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += gridDim.x * blockDim.x) {
        const unsigned row_ = v / (unsigned)g.nz;
        const int k = (int)(v - row_ * (unsigned)g.nz);
        const int i = (int)(row_ / (unsigned)g.ny);
        const int j = (int)(row_ - (unsigned)i * (unsigned)g.ny);
        const size_t o = (size_t)row_ * g.nzp + k;
        const int gi = g.x0 + i;
        const double p[3] = {A.dv[0] * (gi + 0.5) + A.x0[0], A.dv[1] * (j + 0.5) + A.x0[1], A.dv[2] * (k + 0.5) + A.x0[2]};
        const int cell = ((gi / A.cs - A.cx0) * A.ncy + j / A.cs) * A.ncz + k / A.cs;
        const int cb = A.cell_start[cell], ce = A.cell_start[cell + 1];
        double phi[FGB_MAX_PHASES];
        for (int m = 0; m < A.nmat; m++) {
            if (m == A.matrix_mat) { phi[m] = 1.0; continue; }
            int list[PH_MAXINFO], n = 0;
            for (int q = cb; q < ce; q++) {
                const int fi = A.cell_fibs[q];
                const CapsuleDev& f = A.fib[fi];
                if (f.mat != m) continue;
                double x[3];
                const double d = capsule_distance(f, p, x);
                if (d <= r_voxel) {
                    if (n < PH_MAXINFO) list[n++] = fi;
                    else atomicOr(A.flag, 2);
                }
            }
            phi[m] = (n > 0) ? integrate_phi_voxel(A, list, n, p, r_voxel) / V_voxel : 0.0;
        }
        // normalizePhi fg:17588-17646: the last material has the highest priority
        double rem = 1;
        for (int m = A.nmat - 1; m >= 0; m--) {
            const double vol = fmin(rem, phi[m]);
            A.phi[m][o] = vol;
            rem -= vol;
        }
        if (A.normals || A.orient) {
            // closest fibre of any material among the candidates of this cell (fg:6885-6924); voxels without a candidate get zeros
            int best = -1;
            double dbest = INFINITY;
            for (int q = cb; q < ce; q++) {
                double x[3];
                const double d = capsule_distance(A.fib[A.cell_fibs[q]], p, x);
                if (d < dbest) { dbest = d; best = A.cell_fibs[q]; }
            }
            double nrm[3] = {0, 0, 0}, ax[3] = {0, 0, 0};
            if (best >= 0) {
                capsule_grad(A.fib[best], p, nrm);
                ax[0] = A.fib[best].a[0]; ax[1] = A.fib[best].a[1]; ax[2] = A.fib[best].a[2];
            }
            for (int a = 0; a < 3; a++) {
                if (A.normals) A.normals[(size_t)a * g.plane + o] = nrm[a];
                if (A.orient) A.orient[(size_t)a * g.plane + o] = ax[a];
            }
        }
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------
static double hdot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double hnorm3(const double* a) { return std::sqrt(hdot3(a, a)); }

// orthonormal_vector fg:605-623
static void orthonormal_vector(const double* v, double* x) {
    int i_max = 0, i_min = 0;
    for (int i = 0; i < 3; i++) {
        if (std::fabs(v[i]) < std::fabs(v[i_min])) i_min = i;
        if (std::fabs(v[i]) > std::fabs(v[i_max])) i_max = i;
    }
    if (i_min == i_max) i_min = (i_max + 1) % 3;
    x[0] = v[0]; x[1] = v[1]; x[2] = v[2];
    x[i_min] = -v[i_max];
    x[i_max] = v[i_min];
    const double s = hdot3(x, v);
    for (int i = 0; i < 3; i++) x[i] = x[i] - s * x[i];
    const double n = hnorm3(x);
    for (int i = 0; i < 3; i++) x[i] = x[i] / n;
}

// distance of point q to the segment c1 + t a, t in [0, L]
static double seg_dist(const CapsuleDev& f, const double* q) {
    double pc[3] = {q[0] - f.c1[0], q[1] - f.c1[1], q[2] - f.c1[2]};
    double t = std::min(std::max(0.0, hdot3(pc, f.a)), f.L);
    double r[3] = {pc[0] - t * f.a[0], pc[1] - t * f.a[1], pc[2] - t * f.a[2]};
    return hnorm3(r);
}

extern "C" int fgb_init_phase_capsules(fgb_ctx* c, int nfib, const fgb_capsule* fibers, int matrix_mat, int smooth_levels, double smooth_tol,
                                       const double* x0, int with_normals, int with_orientation) {
    if (!c) return FGB_EINVAL;
    cudaSetDevice(c->device);
    if (c->nphases < 1) return fgb_fail(c, FGB_EINVAL, "fgb_init_phase_capsules: set the number of phases first");
    if (matrix_mat < 0 || matrix_mat >= c->nphases) return fgb_fail(c, FGB_EINVAL, "matrix material %d out of range", matrix_mat);
    if (nfib < 0 || (nfib > 0 && !fibers)) return fgb_fail(c, FGB_EINVAL, "invalid fibre list");
    // full_staggered: initPhi runs on the doubly fine grid (select_dfg fg:17154-17156); half_staggered: on the coarse grid, the fine
    // phases are injected afterwards (fg:17160-17164)
    const bool fine = c->dfg == 2;
    const GridDev& g = fine ? c->gf : c->g;
    PhaseArgs A;
    memset(&A, 0, sizeof(A));
    A.g = g;
    for (int a = 0; a < 3; a++) A.x0[a] = x0 ? x0[a] : 0.0;
    A.dv[0] = c->L[0] / g.nx; A.dv[1] = c->L[1] / g.ny; A.dv[2] = c->L[2] / g.nz;          // voxel size of the grid being initialised
    A.nmat = c->nphases;
    A.matrix_mat = matrix_mat;
    A.smooth_levels = smooth_levels;
    A.smooth_tol = smooth_tol;
    for (int l = 0; l < PH_MAXDEPTH + 2; l++) A.lvlpow[l] = std::pow(std::ldexp(1.0, -l), 2.0 / 3.0);
    // CapsuleFiber(c, a, L0, R) fg:5254-5277
    std::vector<CapsuleDev> fib(nfib > 0 ? nfib : 1);
    for (int q = 0; q < nfib; q++) {
        CapsuleDev& f = fib[q];
        const fgb_capsule& s = fibers[q];
        if (s.material < 0 || s.material >= c->nphases) return fgb_fail(c, FGB_EINVAL, "fibre %d: material %d out of range", q, s.material);
        f.R = std::fabs(s.R);
        f.L = std::max(0.0, std::fabs(s.L0) - (4.0 / 3.0) * f.R);
        const double na = hnorm3(s.a);
        if (na == 0 && f.L != 0) return fgb_fail(c, FGB_EINVAL, "CapsuleFiber: given nonzero fiber length without orientation vector!");
        double o[3];
        for (int a = 0; a < 3; a++) f.a[a] = (na != 0) ? s.a[a] / na : 0.0;
        for (int a = 0; a < 3; a++) f.c1[a] = s.c[a] - (f.L / 2) * f.a[a];
        orthonormal_vector(f.a, o);
        for (int a = 0; a < 3; a++) f.r[a] = o[a] * f.R;
        f.mat = s.material;
    }
    // uniform cell list over this rank's slab: a fibre is listed in every cell whose centre is within R + r_voxel + half cell diagonal
    // of its axis segment
    const int cs = 8;
    A.cs = cs;
    A.cx0 = g.x0 / cs;
    A.ncx = (g.x0 + g.lnx + cs - 1) / cs - A.cx0;
    A.ncy = (g.ny + cs - 1) / cs;
    A.ncz = (g.nz + cs - 1) / cs;
    const size_t ncells = (size_t)A.ncx * A.ncy * A.ncz;
    const double r_voxel = 0.5 * std::sqrt(A.dv[0] * A.dv[0] + A.dv[1] * A.dv[1] + A.dv[2] * A.dv[2]);
    const double cd[3] = {cs * A.dv[0], cs * A.dv[1], cs * A.dv[2]};
    const double hd = 0.5 * std::sqrt(cd[0] * cd[0] + cd[1] * cd[1] + cd[2] * cd[2]);
    std::vector<int> count(ncells + 1, 0);
    std::vector<std::pair<size_t, int>> entries;
    for (int q = 0; q < nfib; q++) {
        const CapsuleDev& f = fib[q];
        const double reach = f.R + r_voxel + hd;
        int lo[3], hi[3];
        const int nc[3] = {A.ncx, A.ncy, A.ncz};
        const int off[3] = {A.cx0, 0, 0};
        bool empty = false;
        for (int a = 0; a < 3; a++) {
            const double e0 = f.c1[a], e1 = f.c1[a] + f.L * f.a[a];
            const double mn = std::min(e0, e1) - reach - A.x0[a], mx = std::max(e0, e1) + reach - A.x0[a];
            lo[a] = std::max((int)std::floor(mn / cd[a]) - off[a], 0);
            hi[a] = std::min((int)std::floor(mx / cd[a]) - off[a], nc[a] - 1);
            if (lo[a] > hi[a]) empty = true;
        }
        if (empty) continue;
        for (int ci = lo[0]; ci <= hi[0]; ci++)
            for (int cj = lo[1]; cj <= hi[1]; cj++)
                for (int ck = lo[2]; ck <= hi[2]; ck++) {
                    const double cc[3] = {A.x0[0] + (ci + A.cx0 + 0.5) * cd[0], A.x0[1] + (cj + 0.5) * cd[1], A.x0[2] + (ck + 0.5) * cd[2]};
                    if (seg_dist(f, cc) > reach) continue;
                    const size_t cell = ((size_t)ci * A.ncy + cj) * A.ncz + ck;
                    entries.push_back(std::make_pair(cell, q));
                    count[cell + 1]++;
                }
    }
    for (size_t i = 0; i < ncells; i++) count[i + 1] += count[i];
    std::vector<int> fibs(entries.size() > 0 ? entries.size() : 1), cursor(count.begin(), count.end() - 1);
    for (auto& e : entries) fibs[cursor[e.first]++] = e.second;      // fibre order inside a cell = list order (ascending q)

    int *d_start = nullptr, *d_fibs = nullptr;
    CapsuleDev* d_fib = nullptr;
    auto cleanup = [&]() { if (d_start) cudaFree(d_start); if (d_fibs) cudaFree(d_fibs); if (d_fib) cudaFree(d_fib); };
#define PH_CUDA(call)                                                                                                     \
    do {                                                                                                                  \
        cudaError_t e__ = (call);                                                                                         \
        if (e__ != cudaSuccess) { cleanup(); return fgb_fail(c, FGB_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e__)); } \
    } while (0)
    PH_CUDA(cudaMalloc(&d_start, sizeof(int) * count.size()));
    PH_CUDA(cudaMalloc(&d_fibs, sizeof(int) * fibs.size()));
    PH_CUDA(cudaMalloc(&d_fib, sizeof(CapsuleDev) * fib.size()));
    PH_CUDA(cudaMemcpyAsync(d_start, count.data(), sizeof(int) * count.size(), cudaMemcpyHostToDevice, c->stream));
    PH_CUDA(cudaMemcpyAsync(d_fibs, fibs.data(), sizeof(int) * fibs.size(), cudaMemcpyHostToDevice, c->stream));
    PH_CUDA(cudaMemcpyAsync(d_fib, fib.data(), sizeof(CapsuleDev) * fib.size(), cudaMemcpyHostToDevice, c->stream));
    A.cell_start = d_start;
    A.cell_fibs = d_fibs;
    A.fib = d_fib;
    double** phi_dst = fine ? c->phi_f : c->phi;
    for (int m = 0; m < c->nphases; m++) {
        if (!phi_dst[m]) PH_CUDA(cudaMalloc(&phi_dst[m], sizeof(double) * g.plane));
        PH_CUDA(cudaMemsetAsync(phi_dst[m], 0, sizeof(double) * g.plane, c->stream));      // padding: zeros (the reference writes NaN there)
        A.phi[m] = phi_dst[m];
    }
    // normals / orientation are sampled on the fine grid whenever the doubly fine grid is in use (fg:14911-14937); with
    // half_staggered that needs a second pass over the fine grid, which this build does not provide
    if ((with_normals || with_orientation) && c->dfg == 1) {
        cleanup();
        return fgb_fail(c, FGB_EUNSUPPORTED, "normals / orientation from the device phase initialisation need gamma_scheme staggered or full_staggered");
    }
    double** nrm_dst = fine ? &c->normals_f : &c->normals;
    double** ori_dst = fine ? &c->orient_f : &c->orient;
    if (with_normals) {
        if (!*nrm_dst) PH_CUDA(cudaMalloc(nrm_dst, sizeof(double) * g.plane * 3));
        PH_CUDA(cudaMemsetAsync(*nrm_dst, 0, sizeof(double) * g.plane * 3, c->stream));
        A.normals = *nrm_dst;
    }
    if (with_orientation) {
        if (!*ori_dst) PH_CUDA(cudaMalloc(ori_dst, sizeof(double) * g.plane * 3));
        PH_CUDA(cudaMemsetAsync(*ori_dst, 0, sizeof(double) * g.plane * 3, c->stream));
        A.orient = *ori_dst;
    }
    A.flag = c->d_flag;
    c->phi_halo_valid = false;
    c->heatK_valid = false;
    const size_t nvox = (size_t)g.lnx * g.ny * g.nz;
    {
        ProfScope ps(c, "init_phase");
        unsigned grid = (unsigned)std::min<size_t>((nvox + 127) / 128, (size_t)c->sm_count * 64);
        if (grid < 1) grid = 1;
        k_init_phi<<<grid, 128, 0, c->stream>>>(A);
        c->launches++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { cleanup(); return fgb_fail(c, FGB_ECUDA, "launch of k_init_phi failed: %s", cudaGetErrorString(e)); }
    }
    if (c->dfg == 1) {
        for (int m = 0; m < c->nphases; m++) {
            if (!c->phi_f[m]) {
                PH_CUDA(cudaMalloc(&c->phi_f[m], sizeof(double) * c->gf.plane));
                PH_CUDA(cudaMemsetAsync(c->phi_f[m], 0, sizeof(double) * c->gf.plane, c->stream));
            }
            int rci = fgb_k_inject_phase(c, c->phi[m], c->phi_f[m]);
            if (rci) { cleanup(); return rci; }
        }
    }
    PH_CUDA(cudaMemcpyAsync(c->h_flag, c->d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PH_CUDA(cudaStreamSynchronize(c->stream));
    cleanup();
    if (*c->h_flag & 2) {
        cudaMemsetAsync(c->d_flag, 0, sizeof(int), c->stream);
        return fgb_fail(c, FGB_EUNSUPPORTED, "more than %d fibres of one material within a voxel radius of a voxel centre", PH_MAXINFO);
    }
    return FGB_OK;
}

// Phase::phi back to the host (writeRawPhase fg:17004 reads the same planes)
extern "C" int fgb_get_phase(fgb_ctx* c, int phase, double* phi_plane) {
    if (!c) return FGB_EINVAL;
    cudaSetDevice(c->device);
    const double* src = (c->dfg == 2) ? c->phi_f[phase] : c->phi[phase];          // full_staggered: the fine-grid plane
    if (phase < 0 || phase >= c->nphases || !src) return fgb_fail(c, FGB_EINVAL, "phase %d not initialised", phase);
    FGB_CUDA(c, cudaMemcpyAsync(phi_plane, src, sizeof(double) * (c->dfg == 2 ? c->gf.plane : c->g.plane), cudaMemcpyDeviceToHost, c->stream));
    FGB_CUDA(c, cudaStreamSynchronize(c->stream));
    return FGB_OK;
}
