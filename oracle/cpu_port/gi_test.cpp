// CPU harness for the block-cooperative GI solver (SerialBlock instantiation) -- test infrastructure.
#include <vector>
#include <cstring>
#include "ftmpc_gi.cuh"
#include "ftmpc_linalg.cuh"
using namespace ftmpc;

struct CsrCons {
    const int* ptr; const int* idx; const double* val; const double* beta;
    inline void row(int p, SparseRow& r) const {
        r.nnz = ptr[p + 1] - ptr[p];
        for (int k = 0; k < r.nnz; ++k) { r.idx[k] = idx[ptr[p] + k]; r.val[k] = val[ptr[p] + k]; }
        r.beta = beta[p];
    }
    inline double slack(int p, const double* v, double sb) const { return cons_slack_generic(*this, p, v, sb); }
};

// min 1/2 x'Gx + a'x  s.t. rows (CSR, <=12 nnz each): n_i'x >= beta_i ; first meq are equalities
extern "C" int gi_test(int n, int m, int meq, const double* G, const double* a, const int* ptr, const int* idx,
                       const double* val, const double* beta, double* x, double* lam, int* iters, int* nact) {
    const int ld = n | 1;
    std::vector<double> E((size_t)n * ld, 0.0), Ui((size_t)n * (n + 1) / 2 + 1), xe(n), s(m), u(n + 2), d(n), ze(n), r(n),
        cs(2 * n + 2), tmp(n + 2), sub(n + 2), dg(n), y(n);
    std::vector<int> act(n + 2), pos(m), itmp(n + 2);
    std::vector<double> esign(m + 1);
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) E[(size_t)i * ld + j] = G[i * n + j];
    SerialBlock blk;
    if (chol_lower(blk, n, ld, E.data(), 1e-300)) return -1;
    tri_inv_transpose(blk, n, ld, E.data(), dg.data());
    for (int i = 0; i < n; ++i) { double v = 0; for (int rr = 0; rr <= i; ++rr) v += E[(size_t)rr * ld + i] * a[rr]; y[i] = v; }
    for (int i = 0; i < n; ++i) { double v = 0; for (int k = 0; k < n; ++k) v += E[(size_t)i * ld + k] * y[k]; xe[i] = -v; }
    GiWork w{E.data(), Ui.data(), xe.data(), s.data(), u.data(), d.data(), ze.data(), r.data(), cs.data(), tmp.data(),
             sub.data(), act.data(), pos.data(), itmp.data(), esign.data()};
    CsrCons cons{ptr, idx, val, beta};
    int st = gi_solve(blk, cons, w, n, n, ld, m, meq, lam, 20 * (n + m), 1e-11, iters, nact);
    std::memcpy(x, xe.data(), n * sizeof(double));
    return st;
}
