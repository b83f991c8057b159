/*
 * TEST INFRASTRUCTURE ONLY (oracle/).  Not part of the product path.
 *
 * Plain-C restatement of the reference's per-candidate evaluator, used as the timed CPU baseline
 * ("port") when oracle/_ref (the reference compiled from its own sources) is unavailable, and as a
 * third, independent implementation in tests/test_oracle.py.  Follows, line by line in meaning:
 *   GP_Maximizer::posterior   cpp/optimizeGP.cpp:57-170   (kernel switch :70-112, prior :116-139,
 *                                                          sig2 clip [1e-8, 10] :149-157)
 *   GP_Maximizer::aMb         cpp/optimizeGP.cpp:174-191  (dense mat-vec + dot)
 *   negei / negpi / negucb    cpp/optimizeGP.cpp:194-236
 * Build: gcc -O3 -fPIC -shared -o oracle/_build/liboracle_port.so oracle/oracle_port.c -lm
 */
#include <math.h>
#include <stdlib.h>

typedef struct {
    int nd, nx, kerneltype, npbases;
    const double *invR, *X, *Y, *hyper, *pmeans, *pbeta, *plowerb, *pwidth;
    double ptheta, parm, noise, sf2, maxY;
} port_model;

static double aMb(int n, const double* a, const double* M, const double* b, double* Mb)
{
    for (int i = 0; i < n; i++) {
        double s = 0;
        for (int j = 0; j < n; j++) s += M[(size_t)i * n + j] * b[j];
        Mb[i] = s;
    }
    double x = 0;
    for (int i = 0; i < n; i++) x += Mb[i] * a[i];
    return x;
}

static void posterior(const port_model* m, const double* x, double* mu, double* sigma, double* r, double* tmp, double* ymu)
{
    const int NX = m->nx, NA = m->nd;
    for (int i = 0; i < NX; i++) {
        double z = 0;
        switch (m->kerneltype) {
        case 0: for (int j = 0; j < NA; j++) z += 1 / pow(m->hyper[j], 2) * pow(m->X[NA * i + j] - x[j], 2);
                r[i] = m->sf2 * exp(-.5 * z); break;
        case 1: for (int j = 0; j < NA; j++) z += pow((m->X[NA * i + j] - x[j]) / m->hyper[0], 2);
                r[i] = m->sf2 * exp(-.5 * z); break;
        case 2: for (int j = 0; j < NA; j++) z += pow((m->X[NA * i + j] - x[j]) / m->hyper[0], 2);
                z = sqrt(3) * sqrt(z); r[i] = m->sf2 * (1.0 + z) * exp(-z); break;
        default: for (int j = 0; j < NA; j++) z += pow(m->X[NA * i + j] - x[j], 2);
                z = sqrt(z);
                r[i] = m->sf2 * (1.0 + sqrt(5) * z / m->hyper[0] + 5 * z * z / (3 * m->hyper[0] * m->hyper[0])) * exp(-(sqrt(5) * z / m->hyper[0]));
                break;
        }
    }
    double ypred;
    if (m->npbases > 0) {
        double pm = 0.0;
        for (int i = 0; i < m->npbases; i++) {
            double d = 0;
            for (int j = 0; j < NA; j++) d += pow((x[j] - m->plowerb[j]) / m->pwidth[j] - m->pmeans[i * NA + j], 2);
            pm += m->pbeta[i] * exp(-m->ptheta * d);
        }
        for (int i = 0; i < NX; i++) ymu[i] = m->Y[i] - pm;
        ypred = pm + aMb(NX, r, m->invR, ymu, tmp);
    } else {
        ypred = aMb(NX, r, m->invR, m->Y, tmp);
    }
    double sig2 = 1. + m->noise - aMb(NX, r, m->invR, r, tmp);
    if (sig2 < 1e-8) sig2 = 1e-8; else if (sig2 > 10.) sig2 = 10.;
    *sigma = sqrt(sig2);
    *mu = ypred;
}

/* acq: 0 EI, 1 PI, 2 UCB.  out[m] = negated acquisition (what the reference minimises). */
void port_eval(int nd, const double* invR, const double* X, const double* Y, int nx, int kerneltype, const double* hyper,
               int npbases, const double* pmeans, const double* pbeta, double ptheta, const double* plowerb, const double* pwidth,
               double parm, double noise, int acq, long M, const double* Xs, double* out, double* mu_out, double* sigma_out)
{
    port_model m = {nd, nx, kerneltype, npbases, invR, X, Y, hyper, pmeans, pbeta, plowerb, pwidth, ptheta, parm, noise, 1.0, Y[0]};
    if (kerneltype > 2) m.sf2 = exp(2.0 * log(hyper[nd]));            /* cpp/optimizeGP.cpp:311-314 */
    for (int i = 0; i < nx; i++) if (Y[i] > m.maxY) m.maxY = Y[i];
    double* r = (double*)malloc(sizeof(double) * nx * 3);
    for (long c = 0; c < M; c++) {
        double mu, sigma;
        posterior(&m, Xs + c * nd, &mu, &sigma, r, r + nx, r + 2 * nx);
        double v;
        if (acq == 2) v = -(mu + parm * sigma);
        else {
            double ydiff = mu - m.maxY - parm, Z = ydiff / sigma;
            double cdf = 0.5 * (1. + erf(Z / sqrt(2.)));
            if (acq == 1) v = -cdf;
            else { double pdf = exp(-(Z * Z / 2.)) / (sqrt(2. * M_PI)); v = -(ydiff * cdf + sigma * pdf); }
        }
        out[c] = v;
        if (mu_out) mu_out[c] = mu;
        if (sigma_out) sigma_out[c] = sigma;
    }
    free(r);
}
