// ftmpc_terminal.cuh -- terminal cost V_f(e) with exact gradient and Hessian.
//
// Replaces the sympy -> CasADi terminal cost the reference loads with
//   load_terminal_ingredients    ft_mpc/controllers/tools/terminal_ingredients.py:451-474
// from ft_mpc/config/terminal.yaml and uses at ft_mpc/controllers/spiraling_mpc.py:195-196.
// The expression is carried as a numeric term table (tools/gen_terminal_data.py):
//   V_f(e) = c0 + sum_k c_k prod_i e_i^p_ki + sum_j d_j (prod_i e_i^q_ji + eps_j)^w_j
#pragma once
#include "ftmpc.h"
#include "ftmpc_block.cuh"

namespace ftmpc {

FT_HD double ipow(double x, int p) {
    double r = 1.0;
    for (int i = 0; i < p; ++i) r *= x;
    return r;
}

// monomial value and derivatives.  g[9], H[81] are ACCUMULATED with weight wgt (H symmetric, full).
// Returns the monomial value.  mg/mH (optional) receive the raw monomial gradient/Hessian instead.
FT_HD double monomial(const int8_t* p, const double* e, int* vars, double* a0, double* a1, double* a2, int& nvr) {
    nvr = 0;
    double val = 1.0;
    for (int i = 0; i < FTMPC_NE; ++i) {
        const int pi = p[i];
        if (pi > 0) {
            vars[nvr] = i;
            a0[nvr] = ipow(e[i], pi);
            a1[nvr] = pi * ipow(e[i], pi - 1);
            a2[nvr] = (pi > 1) ? pi * (pi - 1) * ipow(e[i], pi - 2) : 0.0;
            val *= a0[nvr];
            ++nvr;
        }
    }
    return val;
}

// V only
FT_HD double terminal_value(const ftmpc_config& c, const double* e) {
    double v = c.term_const;
    for (int k = 0; k < c.n_poly; ++k) {
        double m = 1.0;
        for (int i = 0; i < FTMPC_NE; ++i) m *= ipow(e[i], c.poly_e[k][i]);
        v += c.poly_c[k] * m;
    }
    for (int k = 0; k < c.n_root; ++k) {
        double m = 1.0;
        for (int i = 0; i < FTMPC_NE; ++i) m *= ipow(e[i], c.root_e[k][i]);
        const double base = m + c.root_eps[k];
        v += c.root_c[k] * ((c.root_pow[k] == 0.25) ? sqrt(sqrt(base)) : pow(base, c.root_pow[k]));
    }
    return v;
}

// V, gradient g[9], Hessian H[81] (row-major, symmetric)
FT_HD double terminal_eval(const ftmpc_config& c, const double* e, double* g, double* H) {
    for (int i = 0; i < FTMPC_NE; ++i) g[i] = 0.0;
    for (int i = 0; i < FTMPC_NE * FTMPC_NE; ++i) H[i] = 0.0;
    double v = c.term_const;
    int vars[FTMPC_NE], nvr;
    double a0[FTMPC_NE], a1[FTMPC_NE], a2[FTMPC_NE], mg[FTMPC_NE];
    const int nterms = c.n_poly + c.n_root;
    for (int k = 0; k < nterms; ++k) {
        const bool is_root = k >= c.n_poly;
        const int kk = is_root ? k - c.n_poly : k;
        const int8_t* p = is_root ? c.root_e[kk] : c.poly_e[kk];
        const double m = monomial(p, e, vars, a0, a1, a2, nvr);
        // outer function  phi(m): value, phi', phi''
        double ph, ph1, ph2;
        if (!is_root) {
            ph = c.poly_c[kk] * m; ph1 = c.poly_c[kk]; ph2 = 0.0;
        } else {
            const double base = m + c.root_eps[kk], w = c.root_pow[kk];
            const double pw = (w == 0.25) ? sqrt(sqrt(base)) : pow(base, w);
            ph = c.root_c[kk] * pw;
            ph1 = c.root_c[kk] * w * pw / base;
            ph2 = c.root_c[kk] * w * (w - 1.0) * pw / (base * base);
        }
        v += ph;
        // monomial gradient
        for (int a = 0; a < nvr; ++a) {
            double pr = a1[a];
            for (int b = 0; b < nvr; ++b) if (b != a) pr *= a0[b];
            mg[a] = pr;
            g[vars[a]] += ph1 * pr;
        }
        // Hessian: phi' * m_ij + phi'' * m_i m_j
        for (int a = 0; a < nvr; ++a) {
            for (int b = 0; b <= a; ++b) {
                double mij;
                if (a == b) {
                    mij = a2[a];
                    for (int cc = 0; cc < nvr; ++cc) if (cc != a) mij *= a0[cc];
                } else {
                    mij = a1[a] * a1[b];
                    for (int cc = 0; cc < nvr; ++cc) if (cc != a && cc != b) mij *= a0[cc];
                }
                const double h = ph1 * mij + ph2 * mg[a] * mg[b];
                H[vars[a] * FTMPC_NE + vars[b]] += h;
                if (a != b) H[vars[b] * FTMPC_NE + vars[a]] += h;
            }
        }
    }
    return v;
}

#if defined(__CUDACC__)
// ---- term-parallel evaluation (device only) -----------------------------------------------------------
// One thread per term of V_f.  A term touches at most three of the nine error coordinates (checked at
// ftmpc_create), so its value / gradient / Hessian contributions are closed-form register arithmetic.
#define FTMPC_TERM_REC 91          /* record per term: value, gradient (9), Hessian (81) */
struct TermDesc {
    int var[3], pw[3];
    double c, eps, w;
    bool root;
};
__device__ __forceinline__ TermDesc term_desc(const ftmpc_config& cg, int k) {
    TermDesc d;
    d.root = k >= cg.n_poly;
    const int kk = d.root ? k - cg.n_poly : k;
    const int8_t* p = d.root ? cg.root_e[kk] : cg.poly_e[kk];
    d.c = d.root ? cg.root_c[kk] : cg.poly_c[kk];
    d.eps = d.root ? cg.root_eps[kk] : 0.0;
    d.w = d.root ? cg.root_pow[kk] : 1.0;
    d.var[0] = d.var[1] = d.var[2] = 0;
    d.pw[0] = d.pw[1] = d.pw[2] = 0;
    int nv = 0;
#pragma unroll
    for (int i = 0; i < FTMPC_NE; ++i) {
        const int pi = p[i];
        if (pi > 0) {
            if (nv == 0) { d.var[0] = i; d.pw[0] = pi; }
            else if (nv == 1) { d.var[1] = i; d.pw[1] = pi; }
            else { d.var[2] = i; d.pw[2] = pi; }
            ++nv;
        }
    }
    return d;
}
__device__ __forceinline__ double ipow_dyn(double x, int p) {
    double r = 1.0;
    for (int i = 0; i < p; ++i) r *= x;
    return r;
}
// value of one term; when rec != nullptr also its gradient / Hessian contributions (rec must be zero-filled)
__device__ __forceinline__ double term_eval(const TermDesc& d, const double* e, double* rec) {
    double a0[3], a1[3], a2[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double x = e[d.var[a]];
        const int p = d.pw[a];
        a0[a] = ipow_dyn(x, p);
        a1[a] = (p > 0) ? p * ipow_dyn(x, p - 1) : 0.0;
        a2[a] = (p > 1) ? p * (p - 1) * ipow_dyn(x, p - 2) : 0.0;
    }
    const double m = a0[0] * a0[1] * a0[2];
    double ph, ph1, ph2;
    if (!d.root) {
        ph = d.c * m; ph1 = d.c; ph2 = 0.0;
    } else {
        const double base = m + d.eps;
        const double pw = (d.w == 0.25) ? sqrt(sqrt(base)) : pow(base, d.w);
        ph = d.c * pw;
        ph1 = d.c * d.w * pw / base;
        ph2 = d.c * d.w * (d.w - 1.0) * pw / (base * base);
    }
    if (rec) {
        const double o0 = a0[1] * a0[2], o1 = a0[0] * a0[2], o2 = a0[0] * a0[1];
        const double mg[3] = {a1[0] * o0, a1[1] * o1, a1[2] * o2};
        rec[0] = ph;
#pragma unroll
        for (int a = 0; a < 3; ++a) rec[1 + d.var[a]] += ph1 * mg[a];
        const double oth[3] = {o0, o1, o2};
        const double single[3] = {a0[2], a0[1], a0[0]};       // the factor not in the pair (0,1), (0,2), (1,2)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            rec[10 + d.var[a] * FTMPC_NE + d.var[a]] += ph1 * a2[a] * oth[a] + ph2 * mg[a] * mg[a];
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                if (b < a) {
                    const double h = ph1 * a1[a] * a1[b] * single[a + b - 1] + ph2 * mg[a] * mg[b];
                    rec[10 + d.var[a] * FTMPC_NE + d.var[b]] += h;
                    rec[10 + d.var[b] * FTMPC_NE + d.var[a]] += h;
                }
            }
        }
    }
    return ph;
}
#endif  // __CUDACC__

}  // namespace ftmpc
