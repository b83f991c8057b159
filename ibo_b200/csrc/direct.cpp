// Batched DIRECT: the reference's rectangle rules (cpp/direct.cpp:49-65 Rectangle, :111-141 samplef,
// :146-235 divrec, :372-498 selection + division order) restructured so that every iteration issues
// two batches -- all probe points of all potentially-optimal rectangles, then all child centres --
// instead of one objective call per sample.  Floating-point expressions that decide control flow
// (side lengths, centre-to-vertex distances, slopes, the epsilon test) are written exactly as the
// reference writes them so that, on identical objective values, the trajectory is identical:
// same samples, same rectangles, same (FMIN, XMIN, nsamples).
//
// The potentially-optimal test is evaluated from per-distance-class minima in O(R * #classes)
// instead of the reference's O(R^2) pair scan; it evaluates the same slope expressions on the
// extremal member of each class (fl(x - c) and fl(x / c), c > 0, are monotone in x), so accept /
// reject decisions are bit-identical.
#include "../../include/ibo_b200.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <limits>
#include <map>
#include <string>
#include <utility>
#include <vector>

namespace ibo {
void set_error(const std::string& s);
int eval_neg_acq(ibo_model* m, const double* Xs, long n, int acq, double ymax, double parm, int flags, double* y);

namespace {

typedef std::pair<unsigned, double> ind_val;
bool sort_by_val(const ind_val& a, const ind_val& b) { return a.second < b.second; }   // cpp/direct.cpp:44-47

const double MAX_DOUBLE = std::numeric_limits<double>::max();
const double MIN_DOUBLE = std::numeric_limits<double>::min();   // +2.2e-308, as in cpp/direct.h:19

struct Rects {   // structure-of-arrays rectangle store (unit cube coordinates)
    int N = 0;
    std::vector<double> lb, ub, center, d, y;
    size_t size() const { return d.size(); }
};

struct Driver {
    ibo_batch_objective_t f; void* user;
    int N;
    std::vector<double> lowerb, upperb;
    std::vector<char> fixed;
    double FMIN = MAX_DOUBLE;
    std::vector<double> XMIN;
    long nsamples = 0;
    std::vector<double> xbuf, ybuf;   // batch staging (box coordinates)

    // unit cube -> box (cpp/direct.cpp:113-120)
    void to_box(const double* x, double* out) const {
        for (int i = 0; i < N; i++) out[i] = fixed[i] ? lowerb[i] : x[i] * (upperb[i] - lowerb[i]) + lowerb[i];
    }
    // bookkeeping half of samplef (cpp/direct.cpp:122-129): strict <, first sample in call order wins
    void account(const double* x, double y) {
        nsamples += 1;
        if (y < FMIN) {
            FMIN = y;
            XMIN.resize(N);
            for (int i = 0; i < N; i++) XMIN[i] = lowerb[i] + (upperb[i] - lowerb[i]) * x[i];
        }
    }
    void eval(const std::vector<double>& unit_pts, long n, std::vector<double>& y) {
        xbuf.resize((size_t)n * N);
        y.resize(n);
        for (long p = 0; p < n; p++) to_box(&unit_pts[(size_t)p * N], &xbuf[(size_t)p * N]);
        if (n > 0) f(user, n, N, xbuf.data(), y.data());
    }
};

// centre and centre-to-vertex distance exactly as Rectangle::Rectangle (cpp/direct.cpp:54-61)
inline double center_and_d(const double* lb, const double* ub, double* c, int N) {
    double d = 0.0;
    for (int i = 0; i < N; i++) {
        c[i] = lb[i] + (ub[i] - lb[i]) / 2.;
        d += std::pow((lb[i] - c[i]), 2);
    }
    return std::sqrt(d);
}

struct Pending {           // one rectangle being divided
    size_t src;            // index in the rectangle store
    double maxlength;
    std::vector<unsigned> dims;     // long, non-fixed dims in ascending order
    long probe0 = 0;       // offset of its 2*k probe values in the phase-A batch
    long child0 = 0;       // offset of its 2*k child values in the phase-B batch
    std::vector<ind_val> I;
    // children, in the reference's return order [c1_dimA, c3_dimA, c1_dimB, ..., middle]
    std::vector<double> clb, cub, ccenter, cd;
    std::vector<double> old_lb, old_ub; double old_d = 0;
};

// Divides the given rectangles (already in processing order).  Appends the new rectangles to R in the
// reference's order and marks sources for removal.  seq: one rectangle per batch pair (reference call order).
void divide(Driver& D, Rects& R, const std::vector<size_t>& order, bool seq, std::vector<char>& removed) {
    const int N = D.N;
    size_t g0 = 0;
    std::vector<double> pts, yA, yB;
    while (g0 < order.size()) {
        size_t g1 = seq ? g0 + 1 : order.size();
        std::vector<Pending> P(g1 - g0);
        // ---- phase A: probe points at lb + w/3, lb + 2w/3 along every longest side (cpp/direct.cpp:156-192)
        pts.clear();
        long np = 0;
        for (size_t t = g0; t < g1; t++) {
            Pending& p = P[t - g0];
            p.src = order[t];
            const double* lb = &R.lb[p.src * N];
            const double* ub = &R.ub[p.src * N];
            const double* c = &R.center[p.src * N];
            double maxlength = ub[0] - lb[0];          // dim 0 even if fixed (reference quirk)
            for (int i = 1; i < N; i++)
                if (!D.fixed[i] && ub[i] - lb[i] > maxlength) maxlength = ub[i] - lb[i];
            p.maxlength = maxlength;
            p.probe0 = np;
            for (int i = 0; i < N; i++) {
                if (!D.fixed[i] && ub[i] - lb[i] == maxlength) {
                    p.dims.push_back((unsigned)i);
                    size_t o = pts.size();
                    pts.insert(pts.end(), c, c + N);
                    pts.insert(pts.end(), c, c + N);
                    pts[o + i] = lb[i] + maxlength / 3.;
                    pts[o + N + i] = lb[i] + 2. * maxlength / 3.;
                    np += 2;
                }
            }
        }
        D.eval(pts, np, yA);
        // ---- sort the dims by min(sf1, sf2) and build the children (cpp/direct.cpp:181-232)
        std::vector<double> ptsB;
        long nc = 0;
        for (size_t t = 0; t < P.size(); t++) {
            Pending& p = P[t];
            for (size_t a = 0; a < p.dims.size(); a++) {
                double sf1 = yA[p.probe0 + 2 * a], sf2 = yA[p.probe0 + 2 * a + 1];
                if (sf1 < sf2) p.I.push_back(ind_val(p.dims[a], sf1));
                else p.I.push_back(ind_val(p.dims[a], sf2));
            }
            std::sort(p.I.begin(), p.I.end(), sort_by_val);
            p.old_lb.assign(&R.lb[p.src * N], &R.lb[p.src * N] + N);
            p.old_ub.assign(&R.ub[p.src * N], &R.ub[p.src * N] + N);
            p.child0 = nc;
            for (size_t a = 0; a < p.I.size(); a++) {
                unsigned dd = p.I[a].first;
                double dwidth = p.old_ub[dd] - p.old_lb[dd];
                double split1 = p.old_lb[dd] + dwidth / 3.;
                double split2 = p.old_lb[dd] + 2. * dwidth / 3.;
                std::vector<double> lb1(p.old_lb), ub1(p.old_ub), lb3(p.old_lb), ub3(p.old_ub);
                ub1[dd] = split1;
                lb3[dd] = split2;
                std::vector<double> c1(N), c3(N);
                double d1 = center_and_d(lb1.data(), ub1.data(), c1.data(), N);
                p.old_lb[dd] = split1;
                p.old_ub[dd] = split2;
                double d3 = center_and_d(lb3.data(), ub3.data(), c3.data(), N);
                p.clb.insert(p.clb.end(), lb1.begin(), lb1.end()); p.cub.insert(p.cub.end(), ub1.begin(), ub1.end());
                p.ccenter.insert(p.ccenter.end(), c1.begin(), c1.end()); p.cd.push_back(d1);
                p.clb.insert(p.clb.end(), lb3.begin(), lb3.end()); p.cub.insert(p.cub.end(), ub3.begin(), ub3.end());
                p.ccenter.insert(p.ccenter.end(), c3.begin(), c3.end()); p.cd.push_back(d3);
                ptsB.insert(ptsB.end(), c1.begin(), c1.end());
                ptsB.insert(ptsB.end(), c3.begin(), c3.end());
                nc += 2;
            }
            // the middle third keeps the old centre and y; d is recomputed from the shrunk bounds (:226-231)
            double d = 0.0;
            const double* oc = &R.center[p.src * N];
            for (int i = 0; i < N; i++) d += std::pow((p.old_lb[i] - oc[i]), 2);
            p.old_d = std::sqrt(d);
        }
        D.eval(ptsB, nc, yB);
        // ---- replay the reference's call order for FMIN / nsamples, then append the rectangles
        for (size_t t = 0; t < P.size(); t++) {
            Pending& p = P[t];
            for (size_t a = 0; a < 2 * p.dims.size(); a++) D.account(&pts[(size_t)(p.probe0 + a) * N], yA[p.probe0 + a]);
            for (size_t a = 0; a < 2 * p.I.size(); a++) D.account(&ptsB[(size_t)(p.child0 + a) * N], yB[p.child0 + a]);
            for (size_t a = 0; a < 2 * p.I.size(); a++) {
                R.lb.insert(R.lb.end(), &p.clb[a * N], &p.clb[a * N] + N);
                R.ub.insert(R.ub.end(), &p.cub[a * N], &p.cub[a * N] + N);
                R.center.insert(R.center.end(), &p.ccenter[a * N], &p.ccenter[a * N] + N);
                R.d.push_back(p.cd[a]);
                R.y.push_back(yB[p.child0 + a]);
            }
            // middle rectangle (copy of the source with shrunk bounds)
            std::vector<double> oc(&R.center[p.src * N], &R.center[p.src * N] + N);
            double oy = R.y[p.src];
            R.lb.insert(R.lb.end(), p.old_lb.begin(), p.old_lb.end());
            R.ub.insert(R.ub.end(), p.old_ub.begin(), p.old_ub.end());
            R.center.insert(R.center.end(), oc.begin(), oc.end());
            R.d.push_back(p.old_d);
            R.y.push_back(oy);
            removed.resize(R.size(), 0);
            removed[p.src] = 1;
        }
        g0 = g1;
    }
}

void compact(Rects& R, std::vector<char>& removed) {
    const int N = R.N;
    size_t w = 0;
    for (size_t r = 0; r < R.size(); r++) {
        if (r < removed.size() && removed[r]) continue;
        if (w != r) {
            std::memmove(&R.lb[w * N], &R.lb[r * N], sizeof(double) * N);
            std::memmove(&R.ub[w * N], &R.ub[r * N], sizeof(double) * N);
            std::memmove(&R.center[w * N], &R.center[r * N], sizeof(double) * N);
            R.d[w] = R.d[r]; R.y[w] = R.y[r];
        }
        w++;
    }
    R.lb.resize(w * N); R.ub.resize(w * N); R.center.resize(w * N); R.d.resize(w); R.y.resize(w);
    removed.assign(w, 0);
}

// potentially-optimal rectangles, ascending index (cpp/direct.cpp:378-456)
void select(const Rects& R, double FMIN, std::vector<size_t>& potopts) {
    const double epsilon = 10e-10;
    potopts.clear();
    // distance classes: exact d -> (min y, second-smallest y is not needed: I3 compares against others)
    std::map<double, double> cls;   // d -> min y
    for (size_t r = 0; r < R.size(); r++) {
        auto it = cls.find(R.d[r]);
        if (it == cls.end()) cls[R.d[r]] = R.y[r];
        else if (R.y[r] < it->second) it->second = R.y[r];
    }
    std::vector<double> cd, cy;
    for (auto& kv : cls) { cd.push_back(kv.first); cy.push_back(kv.second); }
    const size_t C = cd.size();
    for (size_t j = 0; j < R.size(); j++) {
        const double dj = R.d[j], yj = R.y[j];
        double maxI1 = MIN_DOUBLE, minI2 = MAX_DOUBLE;
        bool breaked = false;
        for (size_t c = 0; c < C && !breaked; c++) {
            if (cd[c] < dj) {
                double val = (yj - cy[c]) / (dj - cd[c]);
                if (val > maxI1) maxI1 = val;
            } else if (cd[c] > dj) {
                double val = (cy[c] - yj) / (cd[c] - dj);
                if (val < minI2) { minI2 = val; if (minI2 <= 0.) breaked = true; }
            } else {
                if (yj > cy[c]) breaked = true;     // some other rectangle of the same size is better
            }
        }
        if (!breaked && maxI1 != MIN_DOUBLE && minI2 != MAX_DOUBLE && minI2 < maxI1) breaked = true;
        if (breaked) continue;
        if (minI2 == MAX_DOUBLE) potopts.push_back(j);
        else if (FMIN == 0.0) { if (yj <= dj * minI2) potopts.push_back(j); }
        else if (epsilon <= (FMIN - yj) / std::abs(FMIN) + (dj / std::abs(FMIN)) * minI2) potopts.push_back(j);
    }
}

int run_direct(ibo_batch_objective_t f, void* user, int ndim, const double* lb, const double* ub, int maxiter, int maxtime,
               int maxsample, int flags, double* fmin, double* xmin, long* nsamples, int* iterations) {
    if (!f || ndim < 1 || !lb || !ub) { set_error("bad argument"); return IBO_E_BADARG; }
    const bool seq = (flags & IBO_FLAG_DIRECT_SEQ) != 0;
    time_t start = time(NULL);
    Driver D;
    D.f = f; D.user = user; D.N = ndim;
    D.lowerb.assign(lb, lb + ndim); D.upperb.assign(ub, ub + ndim);
    D.fixed.resize(ndim);
    for (int i = 0; i < ndim; i++) D.fixed[i] = (lb[i] == ub[i]);
    Rects R; R.N = ndim;
    // first rectangle: the unit cube, sampled at its centre (cpp/direct.cpp:349-357)
    {
        std::vector<double> l(ndim, 0.0), u(ndim, 1.0), c(ndim);
        double d = center_and_d(l.data(), u.data(), c.data(), ndim);
        std::vector<double> y;
        D.eval(c, 1, y);
        D.account(c.data(), y[0]);
        R.lb = l; R.ub = u; R.center = c; R.d.push_back(d); R.y.push_back(y[0]);
    }
    std::vector<char> removed(1, 0);
    {
        std::vector<size_t> order(1, 0);
        divide(D, R, order, seq, removed);
        compact(R, removed);
    }
    int iteration = 0;
    bool done = false;
    std::vector<size_t> potopts, order;
    while (iteration < maxiter && !done) {
        iteration++;
        select(R, D.FMIN, potopts);
        if (potopts.empty()) break;    // "could not divide any more" (cpp/direct.cpp:473-477)
        // division order: highest index first; stop after the rectangle that pushes nsamples past
        // maxsample (cpp/direct.cpp:479-492).  Each rectangle costs 4 samples per longest side, which is
        // known before evaluating, so the cut is taken up front.
        order.clear();
        long ns = D.nsamples;
        for (size_t t = potopts.size(); t-- > 0;) {
            size_t j = potopts[t];
            const double* l = &R.lb[j * ndim];
            const double* u = &R.ub[j * ndim];
            double maxlength = u[0] - l[0];
            for (int i = 1; i < ndim; i++)
                if (!D.fixed[i] && u[i] - l[i] > maxlength) maxlength = u[i] - l[i];
            int k = 0;
            for (int i = 0; i < ndim; i++)
                if (!D.fixed[i] && u[i] - l[i] == maxlength) k++;
            order.push_back(j);
            ns += 4L * k;
            if (ns > (long)(unsigned)maxsample) { done = true; break; }
        }
        divide(D, R, order, seq, removed);
        compact(R, removed);
        if (time(NULL) - start > maxtime) break;
        if (D.nsamples > (long)(unsigned)maxsample) break;
    }
    if (fmin) *fmin = D.FMIN;
    if (xmin) for (int i = 0; i < ndim; i++) xmin[i] = D.XMIN.empty() ? lb[i] : D.XMIN[i];
    if (nsamples) *nsamples = D.nsamples;
    if (iterations) *iterations = iteration;
    return IBO_OK;
}

struct ScalarAdapter { objective_t f; };
void scalar_batch(void* user, long n, int ndim, const double* X, double* y) {
    ScalarAdapter* a = static_cast<ScalarAdapter*>(user);
    std::vector<double> x(ndim);
    for (long p = 0; p < n; p++) {
        std::memcpy(x.data(), X + (size_t)p * ndim, sizeof(double) * ndim);
        y[p] = a->f(ndim, x.data());
    }
}

struct GpuObjective { ibo_model* m; int acq; double ymax, parm; int flags; int rc; };
void gpu_batch(void* user, long n, int ndim, const double* X, double* y) {
    GpuObjective* g = static_cast<GpuObjective*>(user);
    (void)ndim;
    if (g->rc != IBO_OK) { for (long i = 0; i < n; i++) y[i] = 0.0; return; }
    g->rc = eval_neg_acq(g->m, X, n, g->acq, g->ymax, g->parm, g->flags, y);
    if (g->rc != IBO_OK) for (long i = 0; i < n; i++) y[i] = 0.0;
}

}  // namespace
}  // namespace ibo

using namespace ibo;

extern "C" int ibo_direct_batched(ibo_batch_objective_t f, void* user, int ndim, const double* lb, const double* ub, int maxiter,
                                  int maxtime, int maxsample, int flags, double* fmin, double* xmin, long* nsamples, int* iterations) {
    return run_direct(f, user, ndim, lb, ub, maxiter, maxtime, maxsample, flags, fmin, xmin, nsamples, iterations);
}

extern "C" int ibo_acqmax(ibo_model* m, const double* lb, const double* ub, int acq, double ymax, double parm, int flags,
                          int maxiter, int maxtime, int maxsample, double* opt, double* optx, long* nsamples, int* iterations) {
    if (!m || acq < 0 || acq > 2) { set_error("bad argument"); return IBO_E_BADARG; }
    GpuObjective g{m, acq, ymax, parm, flags, IBO_OK};
    double fmin = 0;
    int rc = run_direct(gpu_batch, &g, ibo_model_dim(m), lb, ub, maxiter, maxtime, maxsample, flags, &fmin, optx, nsamples, iterations);
    if (rc) return rc;
    if (g.rc) return g.rc;
    if (opt) *opt = -fmin;
    return IBO_OK;
}

// ---- legacy symbols ---------------------------------------------------------------------------
extern "C" const double* direct(objective_t objective, int ndim, double* lb, double* ub, int maxiter, int maxtime, int maxsample) {
    if (!objective || ndim < 1) return NULL;
    ScalarAdapter a{objective};
    double* res = (double*)malloc(sizeof(double) * (ndim + 1));   // caller frees (cpp/direct.cpp:564-569)
    if (!res) return NULL;
    int rc = run_direct(scalar_batch, &a, ndim, lb, ub, maxiter, maxtime, maxsample, IBO_FLAG_DIRECT_SEQ, &res[0], &res[1], NULL, NULL);
    if (rc) { free(res); return NULL; }
    return res;
}

extern "C" const double* acqmaxGP(int ndim, double* lb, double* ub, double* invR, double* X, double* Y, int nx, int acqfunc,
                                  int kerneltype, double* hyperparams, int npbases, double* pbasismeans, double* pbasisbeta,
                                  double pbasistheta, double* pbasislowerb, double* pbasiswidth, double parm, double noise,
                                  int maxiter, int maxtime, int maxsample) {
    if (acqfunc < 0 || acqfunc > 2) {                     // cpp/optimizeGP.cpp:342-346
        set_error("[C++] unknown acquisition function");
        return NULL;
    }
    if (kerneltype < 0 || kerneltype > 3 || nx < 1 || ndim < 1) return NULL;
    // hyper-parameter layout as GP_Maximizer::posterior reads it (cpp/optimizeGP.cpp:70-112);
    // sf2 = 1 for kernels 0-2 (:303-310), exp(2 log(hyper[1])) for Matern-5/2 (the magnitude slot).
    double sf2 = 1.0;
    int nh = (kerneltype == 0) ? ndim : 1;
    if (kerneltype == 3) sf2 = std::exp(2.0 * std::log(hyperparams[1]));
    ibo_model* m = NULL;
    int info = 0;
    int rc = ibo_model_create_from_inverse(0, kerneltype, hyperparams, nh, X, Y, nx, ndim, noise, invR, sf2, npbases, pbasismeans,
                                           pbasisbeta, pbasistheta, pbasislowerb, pbasiswidth, &m, &info);
    if (rc) return NULL;
    double maxY = Y[0];
    for (int i = 0; i < nx; i++) if (Y[i] > maxY) maxY = Y[i];      // cpp/optimizeGP.cpp:317-322
    double* res = (double*)malloc(sizeof(double) * (ndim + 1));
    double opt = 0;
    rc = res ? ibo_acqmax(m, lb, ub, acqfunc, maxY, parm, IBO_FLAG_MODE_CPP, maxiter, maxtime, maxsample, &opt, &res[1], NULL, NULL) : IBO_E_NOMEM;
    ibo_model_destroy(m);
    if (rc) { free(res); return NULL; }
    res[0] = -opt;     // the reference returns FMIN of the negated acquisition
    return res;
}
