/*
 * fg_phase.c -- CPU restatement (plain C) of fibergen's composite-voxel phase initialisation.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the checker for the device phase initialisation and the source of the
 * phase fractions the parity tests and bench.py feed to both the CUDA path and the oracle.  Nothing under fibergen_b200/
 * may call it.  "fg:N" = /root/reference/src/fibergen.cpp:N.
 *
 *   halfspace_box_cut_volume   fg:1385-1575     volume of a box cut by a half space (Gauss divergence theorem over the faces)
 *   CapsuleFiber               fg:5237-5330     distanceTo / distanceGrad / curvature of a capsule (sphere when L0 <= 4R/3)
 *   FiberCluster::closestFibers fg:3336-3361    all fibres of a material with signed distance <= r (bounding boxes are a filter only)
 *   integratePhiVoxel          fg:16622-16752   adaptive subdivision + half-space cuts
 *   initPhi                    fg:17489-17581   per voxel centre, matrix phase = 1
 *   normalizePhi               fg:17588-17646   last material has the highest priority, fractions sum to 1
 *
 * Pinned by the reference's own self-tests restated in tests/test_oracle_pinning.py (fg:23759-23864: cut volume against the
 * analytic values) and by the Hashin demo's documented result (demo/elasticity/hashin/project.xml:30-32).
 */
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    double c1[3];   /* base point 1 of the cylinder part */
    double a[3];    /* unit axis */
    double r[3];    /* vector orthogonal to a with length R (orthonormal_vector fg:605) */
    double R, L;    /* radius, length of the cylinder part: L = max(0, L0 - 4R/3) (fg:5258) */
    int mat;        /* material index */
} fgo_capsule;

static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double norm3(const double* a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

/* orthonormal_vector fg:605-623 (restated including its Gram-Schmidt line `x - <x,v> x`) */
static void orthonormal_vector(const double* v, double* x) {
    int i_max = 0, i_min = 0, i;
    for (i = 0; i < 3; i++) {
        if (fabs(v[i]) < fabs(v[i_min])) i_min = i;
        if (fabs(v[i]) > fabs(v[i_max])) i_max = i;
    }
    if (i_min == i_max) i_min = (i_max + 1) % 3;
    x[0] = v[0]; x[1] = v[1]; x[2] = v[2];
    x[i_min] = -v[i_max];
    x[i_max] = v[i_min];
    {
        const double s = dot3(x, v);
        double n;
        for (i = 0; i < 3; i++) x[i] = x[i] - s * x[i];
        n = norm3(x);
        for (i = 0; i < 3; i++) x[i] = x[i] / n;
    }
}

/* CapsuleFiber(c, a, L0, R) fg:5254-5277 */
void fgo_capsule_init(fgo_capsule* f, const double* c, const double* a, double L0, double R, int mat) {
    int i;
    double na = norm3(a), o[3];
    L0 = fabs(L0);
    f->R = fabs(R);
    f->L = fmax(0.0, L0 - (4.0 / 3.0) * f->R);
    for (i = 0; i < 3; i++) f->a[i] = (na != 0) ? a[i] / na : 0.0;
    for (i = 0; i < 3; i++) f->c1[i] = c[i] - (f->L / 2) * f->a[i];
    orthonormal_vector(f->a, o);
    for (i = 0; i < 3; i++) f->r[i] = o[i] * f->R;
    f->mat = mat;
}

/* CapsuleFiber::distanceTo fg:5298-5333: signed distance of p to the surface, x = closest surface point */
double fgo_capsule_distance(const fgo_capsule* f, const double* p, double* x) {
    double pc[3], t, d, q[3];
    int i;
    for (i = 0; i < 3; i++) pc[i] = p[i] - f->c1[i];
    t = dot3(pc, f->a);
    t = fmin(fmax(0.0, t), f->L);
    for (i = 0; i < 3; i++) x[i] = f->c1[i] + t * f->a[i];
    for (i = 0; i < 3; i++) q[i] = p[i] - x[i];
    d = norm3(q);
    if (d < DBL_EPSILON * f->R) {
        for (i = 0; i < 3; i++) x[i] += f->r[i];
    } else {
        for (i = 0; i < 3; i++) x[i] += q[i] * (f->R / d);
    }
    return d - f->R;
}

/* CapsuleFiber::distanceGrad fg:5279-5296 */
void fgo_capsule_grad(const fgo_capsule* f, const double* p, double* g) {
    double pc[3], t, n;
    int i;
    for (i = 0; i < 3; i++) pc[i] = p[i] - f->c1[i];
    t = dot3(pc, f->a);
    t = fmin(fmax(0.0, t), f->L);
    for (i = 0; i < 3; i++) g[i] = p[i] - f->c1[i] - t * f->a[i];
    n = norm3(g);
    if (n < sqrt(DBL_EPSILON)) {
        for (i = 0; i < 3; i++) g[i] = ((t < 0.5 * f->L) ? -1 : 1) * f->a[i];
    } else {
        for (i = 0; i < 3; i++) g[i] /= n;
    }
}

/* halfspace_box_cut_volume fg:1385-1575: volume of {y in box(x0, dx,dy,dz) : <y - x, n> < 0} */
double fgo_halfspace_box_cut_volume(const double* x, const double* n, const double* x0, double dx, double dy, double dz) {
    static const int edges[12][2] = {{0, 1}, {2, 4}, {3, 6}, {5, 7}, {0, 2}, {1, 4}, {3, 5}, {6, 7}, {0, 3}, {1, 6}, {2, 5}, {4, 7}};
    static const int faces[6][4] = {{8, 6, -10, -4}, {9, 7, -11, -5}, {0, 9, -2, -8}, {1, 11, -3, -10}, {0, 5, -1, -4}, {2, 7, -3, -6}};
    static const int face_normal_signs[6] = {-1, 1, -1, 1, -1, 1};
    static const int crossp_indices[3][2] = {{1, 2}, {2, 0}, {0, 1}};
    const double d3[3] = {dx, dy, dz};
    double v[8][3], dist[6], xi[3], points[5][3], V = 0;
    int inside[8], num_inside = 0, iedge[12], nint = 0, any = -1, i, f, flip;
    for (i = 0; i < 8; i++) memcpy(v[i], x0, sizeof(double) * 3);
    v[1][0] += dx;
    v[2][1] += dy;
    v[3][2] += dz;
    memcpy(v[4], v[1], sizeof(v[4])); v[4][1] += dy;
    memcpy(v[5], v[2], sizeof(v[5])); v[5][2] += dz;
    memcpy(v[6], v[3], sizeof(v[6])); v[6][0] += dx;
    memcpy(v[7], v[6], sizeof(v[7])); v[7][1] += dy;
    for (i = 0; i < 8; i++) {
        const double q[3] = {v[i][0] - x[0], v[i][1] - x[1], v[i][2] - x[2]};
        inside[i] = dot3(q, n) < 0;
        num_inside += inside[i];
    }
    for (i = 0; i < 12; i++) {
        if (inside[edges[i][0]] + inside[edges[i][1]] == 1) {
            const double* p0 = v[edges[i][0]];
            const double q[3] = {x[0] - p0[0], x[1] - p0[1], x[2] - p0[2]};
            dist[nint] = dot3(q, n) / n[i / 4];
            iedge[i] = nint;
            any = i;
            nint++;
        } else iedge[i] = -1;
    }
    if (nint == 0) return inside[0] ? (dx * dy * dz) : 0;
    memcpy(xi, v[edges[any][0]], sizeof(xi));
    xi[any / 4] += dist[iedge[any]];
    flip = (num_inside > 4);
    for (f = 0; f < 6; f++) {
        const int ni = f >> 1;
        int np = 0, brk = 0;
        for (i = 0; i < 4 && !brk; i++) {
            int e = faces[f][i], i1 = 0, i2 = 1;
            if (e < 0) { e = -e; i1 = 1; i2 = 0; }
            /* note the reference's operator precedence: num_points == 0 && (inside ^ flip) */
            if (np == 0 && (inside[edges[e][i1]] ^ flip)) {
                memcpy(points[np], v[edges[e][i1]], sizeof(points[0]));
                if (points[0][ni] == xi[ni]) { brk = 1; break; }
                np++;
            }
            if (iedge[e] >= 0) {
                memcpy(points[np], v[edges[e][0]], sizeof(points[0]));
                points[np][e / 4] += dist[iedge[e]];
                if (np == 0 && points[0][ni] == xi[ni]) { brk = 1; break; }
                np++;
            }
            if (i < 3 && (inside[edges[e][i2]] ^ flip)) {
                memcpy(points[np], v[edges[e][i2]], sizeof(points[0]));
                if (np == 0 && points[0][ni] == xi[ni]) { brk = 1; break; }
                np++;
            }
        }
        if (np < 3) continue;
        {
            const int i1 = crossp_indices[ni][0], i2 = crossp_indices[ni][1];
            double area = 0, d;
            for (i = 2; i < np; i++)
                area += fabs((points[i - 1][i1] - points[0][i1]) * (points[i][i2] - points[0][i2]) -
                             (points[i - 1][i2] - points[0][i2]) * (points[i][i1] - points[0][i1]));
            d = points[0][ni] - xi[ni];
            V += face_normal_signs[f] * d * area;
        }
    }
    (void)d3;
    V *= (1.0 / 6.0);
    if (flip) V = dx * dy * dz - V;
    return V;
}

typedef struct {
    const fgo_capsule* fiber;
    double d;
    double x[3];
} info_t;

/* integratePhiVoxel fg:16622-16752 */
static double integrate_phi_voxel(int levels, double tol, double r_voxel0, const double* p, double dx, double dy, double dz,
                                  info_t* info, int ninfo) {
    double r_voxel, x0[3], V = 0, V_max = dx * dy * dz;
    int i_min = 0, i, j, k, q;
    if (ninfo == 0) return 0;
    r_voxel = 0.5 * sqrt(dx * dx + dy * dy + dz * dz);
    for (i = 1; i < ninfo; i++)
        if (info[i].d < info[i_min].d) i_min = i;
    if (fabs(info[i_min].d) >= r_voxel) return (info[i_min].d < 0) ? dx * dy * dz : 0;
    x0[0] = p[0] - 0.5 * dx;
    x0[1] = p[1] - 0.5 * dy;
    x0[2] = p[2] - 0.5 * dz;
    if (levels < 0) {
        const double K = 1 / info[i_min].fiber->R;              /* CapsuleFiber::curvature fg:5470 */
        const double Kd = r_voxel * K;
        double err;
        if (Kd > 1) err = 1;
        else err = Kd * Kd * pow(r_voxel / r_voxel0, 2.0 / 3.0);
        if (err < tol) levels = 0;
    }
    if (levels == 0) {
        for (i = 0; i < ninfo; i++) {
            double n[3];
            fgo_capsule_grad(info[i].fiber, info[i].x, n);
            V += fgo_halfspace_box_cut_volume(info[i].x, n, x0, dx, dy, dz);
        }
        return fmin(V, V_max);
    }
    levels--;
    dx *= 0.5;
    dy *= 0.5;
    dz *= 0.5;
    r_voxel *= 0.5;
    {
        info_t* sub = (info_t*)malloc(sizeof(info_t) * (size_t)ninfo);
        double ps[3];
        for (i = 0; i < 2; i++) {
            ps[0] = x0[0] + (i + 0.5) * dx;
            for (j = 0; j < 2; j++) {
                ps[1] = x0[1] + (j + 0.5) * dy;
                for (k = 0; k < 2; k++) {
                    int nsub = 0;
                    ps[2] = x0[2] + (k + 0.5) * dz;
                    for (q = 0; q < ninfo; q++) {
                        info[q].d = fgo_capsule_distance(info[q].fiber, ps, info[q].x);
                        if (fabs(info[q].d) >= r_voxel) {
                            if (info[q].d < 0) {
                                V += dx * dy * dz;
                                nsub = 0;
                                break;
                            }
                            continue;
                        }
                        sub[nsub++] = info[q];
                    }
                    if (nsub != 0) V += integrate_phi_voxel(levels, tol, r_voxel0, ps, dx, dy, dz, sub, nsub);
                }
            }
        }
        free(sub);
    }
    return fmin(V, V_max);
}

/*
 * initPhi fg:17489-17581 + normalizePhi fg:17588-17646 for capsule fibres.
 * phi: nmat arrays of nx*ny*nz doubles (unpadded, x slowest), cell [x0, x0 + L); rows i in [i0, i1) of the global grid are written
 * at phi[m][(i - i0)*ny*nz + ...] so that a slab can be produced on its own.  Every fibre is taken as given (periodic images are
 * separate entries, as the reference's ghost fibres are).  Returns the number of interface voxels found.
 */
long fgo_init_phi(int nx, int ny, int nz, const double* L, const double* x0c, int nfib, const fgo_capsule* fibers, int nmat, int matrix_mat,
                  int smooth_levels, double smooth_tol, int i0, int i1, double* const* phi) {
    const double dxv = L[0] / nx, dyv = L[1] / ny, dzv = L[2] / nz;
    const double V_voxel = dxv * dyv * dzv;
    const double r_voxel = 0.5 * sqrt(dxv * dxv + dyv * dyv + dzv * dzv);
    long ninterface = 0;
    int m;
    for (m = 0; m < nmat; m++) {
        long i;
        if (m == matrix_mat) {
            const size_t n = (size_t)(i1 - i0) * ny * nz;
            size_t q;
            for (q = 0; q < n; q++) phi[m][q] = 1.0;
            continue;
        }
#pragma omp parallel for schedule(dynamic) reduction(+ : ninterface)
        for (i = i0; i < i1; i++) {
            info_t* info = (info_t*)malloc(sizeof(info_t) * (size_t)(nfib > 0 ? nfib : 1));
            double p[3];
            int j, k, q;
            p[0] = dxv * (i + 0.5) + x0c[0];
            for (j = 0; j < ny; j++) {
                p[1] = dyv * (j + 0.5) + x0c[1];
                for (k = 0; k < nz; k++) {
                    int ninfo = 0;
                    double val = 0;
                    p[2] = dzv * (k + 0.5) + x0c[2];
                    for (q = 0; q < nfib; q++) {
                        const fgo_capsule* f = &fibers[q];
                        double c[3], bb;
                        if (f->mat != m) continue;
                        /* bounding-ball filter (IBoundingBox::bbDistanceMin fg:3046): centre c1 + a L/2, radius L/2 + R */
                        c[0] = p[0] - (f->c1[0] + 0.5 * f->L * f->a[0]);
                        c[1] = p[1] - (f->c1[1] + 0.5 * f->L * f->a[1]);
                        c[2] = p[2] - (f->c1[2] + 0.5 * f->L * f->a[2]);
                        bb = norm3(c) - (0.5 * f->L + f->R);
                        if (bb > r_voxel) continue;
                        info[ninfo].d = fgo_capsule_distance(f, p, info[ninfo].x);
                        if (info[ninfo].d <= r_voxel) {
                            info[ninfo].fiber = f;
                            ninfo++;
                        }
                    }
                    if (ninfo > 0) val = integrate_phi_voxel(smooth_levels, smooth_tol, r_voxel, p, dxv, dyv, dzv, info, ninfo) / V_voxel;
                    phi[m][((size_t)(i - i0) * ny + j) * nz + k] = val;
                }
            }
            free(info);
        }
    }
    /* normalizePhi: the last material has the highest priority */
    {
        const size_t n = (size_t)(i1 - i0) * ny * nz;
        size_t q;
        for (q = 0; q < n; q++) {
            double rem = 1;
            int interface_ = 0;
            for (m = nmat - 1; m >= 0; m--) {
                const double vol = fmin(rem, phi[m][q]);
                phi[m][q] = vol;
                rem -= vol;
                if (!(vol == 0 || vol == 1)) interface_ = 1;
            }
            ninterface += interface_;
        }
    }
    return ninterface;
}
