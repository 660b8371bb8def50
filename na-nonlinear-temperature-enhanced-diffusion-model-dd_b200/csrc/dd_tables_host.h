// dd_tables_host.h -- host-side preparation of the separable MMS tables (shared by the C ABI layer
// and the test-only host build): profile de-duplication and the separable cell-average tables.
#pragma once

#include <string.h>

#include <vector>

#include "dd_types.h"

struct DDHostTables {
    int nterms = 1, nprof = 1;
    int var_prof[DD_NVAR] = {0, 0, 0, 0, 0};
    std::vector<double> X[DD_NVAR][3], Y[DD_NVAR][3];  // per profile
    std::vector<double> QX1, QY1, QX2, QY2, QX3, QY3;
};

// X[v][d]: nterms x (N+1); XQ[q] (q: cp, T, cl): nterms x (N+1) x 3; same in y with M+1.
inline void dd_prepare_tables(int nterms, int N, int M, const double* const X[5][3], const double* const Y[5][3],
                              const double* const XQ[3], const double* const YQ[3], DDHostTables* out) {
    const size_t nx = (size_t)nterms * (N + 1), ny = (size_t)nterms * (M + 1);
    out->nterms = nterms;
    out->nprof = 0;
    for (int v = 0; v < DD_NVAR; ++v) {
        int found = -1;
        for (int w = 0; w < v && found < 0; ++w) {
            bool same = true;
            for (int d = 0; d < 3 && same; ++d)
                same = !memcmp(X[v][d], X[w][d], nx * sizeof(double)) && !memcmp(Y[v][d], Y[w][d], ny * sizeof(double));
            if (same) found = out->var_prof[w];
        }
        if (found >= 0) {
            out->var_prof[v] = found;
        } else {
            const int p = out->nprof++;
            out->var_prof[v] = p;
            for (int d = 0; d < 3; ++d) {
                out->X[p][d].assign(X[v][d], X[v][d] + nx);
                out->Y[p][d].assign(Y[v][d], Y[v][d] + ny);
            }
        }
    }
    // Gauss weights of the reference's avg_int (src/prob1base.py:508): the x and y sums factorise
    const double w[3] = {5.0 / 9.0, 8.0 / 9.0, 5.0 / 9.0};
    auto build = [&](const double* A, const double* B, int n1, bool pair, std::vector<double>& out1) {
        // A, B: nterms x n1 x 3.  pair == false: out[r][i] = sum_a w_a A[r][i][a];
        // pair == true: out[r*R+s][i] = sum_a w_a A[r][i][a] B[s][i][a]
        const int R = nterms;
        out1.assign((size_t)(pair ? R * R : R) * n1, 0.0);
        for (int r = 0; r < R; ++r)
            for (int s = 0; s < (pair ? R : 1); ++s)
                for (int i = 0; i < n1; ++i) {
                    double acc = 0.0;
                    for (int a = 0; a < 3; ++a) {
                        const double va = A[((size_t)r * n1 + i) * 3 + a];
                        acc += w[a] * (pair ? va * B[((size_t)s * n1 + i) * 3 + a] : va);
                    }
                    out1[(size_t)(pair ? r * R + s : r) * n1 + i] = acc;
                }
    };
    // XQ order: 0 = cp, 1 = T, 2 = cl
    build(XQ[0], nullptr, N + 1, false, out->QX1);
    build(YQ[0], nullptr, M + 1, false, out->QY1);
    build(XQ[0], XQ[2], N + 1, true, out->QX2);
    build(YQ[0], YQ[2], M + 1, true, out->QY2);
    build(XQ[0], XQ[1], N + 1, true, out->QX3);
    build(YQ[0], YQ[1], M + 1, true, out->QY3);
}
