// TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product path.
//
// Thin harness around the *unmodified* reference translation unit cpp/optimizeGP.cpp
// (pulled in by #include at build time from /root/reference/cpp; nothing is copied into
// this repo).  It exposes the reference's own per-candidate evaluator
//   GP_Maximizer::posterior / negei / negpi / negucb   (reference cpp/optimizeGP.cpp:57-236)
// over a batch of candidates so that tests can pin the oracle restatement against the real
// reference arithmetic and bench.py can time the reference CPU path (`--impl reference`,
// cpu_baseline.kind == "reference").
//
// The statics are assigned exactly as acqmaxGP does (reference cpp/optimizeGP.cpp:285-322);
// after that they are read-only, so evaluating different candidates from several threads
// is safe (posterior() only mallocs private scratch).
#include <iostream>
#include <sstream>
#include <thread>
#include <vector>
#include "optimizeGP.cpp"   // -I/root/reference/cpp

extern "C" void ref_set_model(int ndim, double* invR, double* X, double* Y, int nx, int kerneltype,
                              double* hyperparams, int npbases, double* pbasismeans,
                              double* pbasisbeta, double pbasistheta, double* pbasislowerb,
                              double* pbasiswidth, double parm, double noise)
{
    GP_Maximizer::invR = invR;
    GP_Maximizer::NA = ndim;
    GP_Maximizer::NX = nx;
    GP_Maximizer::X = X;
    GP_Maximizer::Y = Y;
    GP_Maximizer::parm = parm;
    GP_Maximizer::noise = noise;
    GP_Maximizer::kerneltype = kerneltype;
    GP_Maximizer::hyperparams = hyperparams;
    GP_Maximizer::npbases = npbases;
    GP_Maximizer::pbasismeans = pbasismeans;
    GP_Maximizer::pbasisbeta = pbasisbeta;
    GP_Maximizer::pbasistheta = pbasistheta;
    GP_Maximizer::pbasislowerb = pbasislowerb;
    GP_Maximizer::pbasiswidth = pbasiswidth;
    switch (kerneltype) {
    case 0: case 1: case 2: GP_Maximizer::sf2 = 1.0; break;
    default: GP_Maximizer::sf2 = exp(2.0 * log(hyperparams[ndim])); break;
    }
    GP_Maximizer::maxY = Y[0];
    for (int i = 0; i < nx; i++) if (Y[i] > GP_Maximizer::maxY) GP_Maximizer::maxY = Y[i];
}

// acq: 0 EI, 1 PI, 2 UCB (values are the *negated* objective the reference minimises).
// mu/sigma may be NULL. nthreads<=1 => plain loop on the calling thread.
extern "C" void ref_eval(int acq, long M, double* Xs, double* negacq, double* mu, double* sigma, int nthreads)
{
    const int nd = GP_Maximizer::NA;
    // kerneltype 3 streams every r[i] to cout (reference cpp/optimizeGP.cpp:109): swallow it.
    std::ostringstream sink;
    std::streambuf* old = NULL;
    if (GP_Maximizer::kerneltype == 3) { old = std::cout.rdbuf(sink.rdbuf()); nthreads = 1; }
    auto work = [&](long lo, long hi) {
        for (long m = lo; m < hi; m++) {
            double* x = Xs + m * nd;
            double v;
            if (acq == 0) v = GP_Maximizer::negei(nd, x);
            else if (acq == 1) v = GP_Maximizer::negpi(nd, x);
            else v = GP_Maximizer::negucb(nd, x);
            negacq[m] = v;
            if (mu || sigma) {
                double a, b;
                GP_Maximizer::posterior(nd, x, a, b);
                if (mu) mu[m] = a;
                if (sigma) sigma[m] = b;
            }
            if (old) sink.str("");
        }
    };
    if (nthreads <= 1) work(0, M);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++) {
            long lo = M * t / nthreads, hi = M * (t + 1) / nthreads;
            th.push_back(std::thread(work, lo, hi));
        }
        for (size_t t = 0; t < th.size(); t++) th[t].join();
    }
    if (old) std::cout.rdbuf(old);
}
