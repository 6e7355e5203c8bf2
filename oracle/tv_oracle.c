/*
 * oracle/tv_oracle.c -- TEST / BENCH INFRASTRUCTURE ONLY (never linked into libpyxu_b200.so).
 *
 * Multi-threaded C restatement of the reference's CPU execution of one PD3O iteration on a
 * TV-denoising problem  f = alpha*||x - y||^2, g = i_{x>=0} | 0, h = lam*L21, K = Gradient (forward
 * differences, 'constant' boundary), written pass by pass the way the reference runs it on its
 * NumPy/Numba backend: each operator call is its own sweep over the volume
 *   (reference: src/pyxu/opt/solver/pds.py:747-761 PD3O.m_step;
 *    K / K^T: operator/linop/diff.py:1113-1265 -> stencil/_stencil.py:232-305 [numba parallel stencil];
 *    g.prox: func/indicator.py:203-206;  f.grad: func/norm.py:96-98 (+ ArgShift/Scale rules);
 *    h.fenchel_prox: abc/operator.py:906-944 around func/norm.py:352-364).
 * Used as (a) a second, independent checker of the fixtures in tests/, (b) the host-core CPU baseline
 * timed by bench.py (OpenMP over all cores, like numba's parallel=True).
 *
 * Parity: checked against tests/golden/solvers.npz (real reference output) in
 * tests/test_oracle_golden.py::test_c_port_matches_reference_fixture.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define DEFINE_TV(T, SUF)                                                                                   \
    /* ktz[s] = sum_d ( z_d[s - e_d] - z_d[s] ),  z_d[-1] = 0   (adjoint of forward differences) */          \
    static void grad_adjoint_##SUF(const T* z, T* out, int D, int64_t n0, int64_t n1, int64_t n2) {          \
        const int64_t N = n0 * n1 * n2, s0 = n1 * n2, s1 = n2;                                               \
        const T *z0 = z, *z1 = z + N, *z2 = z + 2 * N;                                                       \
        _Pragma("omp parallel for schedule(static)") for (int64_t i0 = 0; i0 < n0; ++i0)                    \
            for (int64_t i1 = 0; i1 < n1; ++i1)                                                              \
                for (int64_t i2 = 0; i2 < n2; ++i2) {                                                        \
                    const int64_t s = i0 * s0 + i1 * s1 + i2;                                                \
                    T acc = 0;                                                                               \
                    if (D == 3) {                                                                            \
                        acc += (i0 > 0 ? z0[s - s0] : (T)0) - z0[s];                                         \
                        acc += (i1 > 0 ? z1[s - s1] : (T)0) - z1[s];                                         \
                        acc += (i2 > 0 ? z2[s - 1] : (T)0) - z2[s];                                          \
                    } else {                                                                                 \
                        acc += (i1 > 0 ? z0[s - s1] : (T)0) - z0[s];                                         \
                        acc += (i2 > 0 ? z1[s - 1] : (T)0) - z1[s];                                          \
                    }                                                                                        \
                    out[s] = acc;                                                                            \
                }                                                                                            \
    }                                                                                                        \
    /* (K w)_d[s] = w[s + e_d] - w[s],  w[n] = 0 */                                                           \
    static void grad_apply_##SUF(const T* w, T* out, int D, int64_t n0, int64_t n1, int64_t n2) {            \
        const int64_t N = n0 * n1 * n2, s0 = n1 * n2, s1 = n2;                                               \
        _Pragma("omp parallel for schedule(static)") for (int64_t i0 = 0; i0 < n0; ++i0)                    \
            for (int64_t i1 = 0; i1 < n1; ++i1)                                                              \
                for (int64_t i2 = 0; i2 < n2; ++i2) {                                                        \
                    const int64_t s = i0 * s0 + i1 * s1 + i2;                                                \
                    const T c = w[s];                                                                        \
                    if (D == 3) {                                                                            \
                        out[s] = (i0 + 1 < n0 ? w[s + s0] : (T)0) - c;                                       \
                        out[N + s] = (i1 + 1 < n1 ? w[s + s1] : (T)0) - c;                                   \
                        out[2 * N + s] = (i2 + 1 < n2 ? w[s + 1] : (T)0) - c;                                \
                    } else {                                                                                 \
                        out[s] = (i1 + 1 < n1 ? w[s + s1] : (T)0) - c;                                       \
                        out[N + s] = (i2 + 1 < n2 ? w[s + 1] : (T)0) - c;                                    \
                    }                                                                                        \
                }                                                                                            \
    }                                                                                                        \
    int tv_pd3o_##SUF(const T* y, T* x, T* u, T* z, int D, int64_t n0, int64_t n1, int64_t n2, double alpha_, \
                      double tau_, double sigma_, double rho_, double lam_, int positivity, int iters) {      \
        const int64_t N = n0 * n1 * n2;                                                                      \
        const T alpha = (T)alpha_, tau = (T)tau_, sigma = (T)sigma_, rho = (T)rho_, lam = (T)lam_;           \
        T* t1 = (T*)malloc(sizeof(T) * N);                                                                   \
        T* ut = (T*)malloc(sizeof(T) * N);                                                                   \
        T* w = (T*)malloc(sizeof(T) * N);                                                                    \
        T* kw = (T*)malloc(sizeof(T) * N * D);                                                               \
        if (!t1 || !ut || !w || !kw) return -1;                                                              \
        for (int it = 0; it < iters; ++it) {                                                                 \
            grad_adjoint_##SUF(z, t1, D, n0, n1, n2);                        /* K^T z */                      \
            _Pragma("omp parallel for schedule(static)") for (int64_t s = 0; s < N; ++s) {                  \
                const T v = u[s] - tau * t1[s];                              /* x = prox_g(u - tau K^T z) */  \
                x[s] = positivity ? (v > 0 ? v : (T)0) : v;                                                  \
            }                                                                                                \
            _Pragma("omp parallel for schedule(static)") for (int64_t s = 0; s < N; ++s) {                  \
                const T gf = (x[s] - y[s]) * ((T)2 * alpha);                 /* grad f(x) */                  \
                ut[s] = x[s] - tau * gf;                                     /* u_temp */                     \
            }                                                                                                \
            _Pragma("omp parallel for schedule(static)") for (int64_t s = 0; s < N; ++s)                    \
                w[s] = x[s] + ut[s] - u[s];                                                                  \
            grad_apply_##SUF(w, kw, D, n0, n1, n2);                          /* K w */                        \
            const T tp = ((T)1 / sigma) * lam;                                                               \
            _Pragma("omp parallel for schedule(static)") for (int64_t s = 0; s < N; ++s) {                  \
                T p[3], a[3], nn = 0;                                        /* fenchel prox of lam*L21 */    \
                for (int d = 0; d < D; ++d) {                                                                \
                    p[d] = z[d * N + s] + sigma * kw[d * N + s];                                             \
                    a[d] = p[d] / sigma;                                                                     \
                    nn += a[d] * a[d];                                                                       \
                }                                                                                            \
                const T nrm = (T)sqrt((double)nn);                                                           \
                const T sc = (T)1 - tp / (nrm > tp ? nrm : tp);                                              \
                for (int d = 0; d < D; ++d) {                                                                \
                    const T zt = p[d] - sigma * (a[d] * sc);                                                 \
                    z[d * N + s] = ((T)1 - rho) * z[d * N + s] + rho * zt;                                   \
                }                                                                                            \
            }                                                                                                \
            _Pragma("omp parallel for schedule(static)") for (int64_t s = 0; s < N; ++s)                    \
                u[s] = ((T)1 - rho) * u[s] + rho * ut[s];                                                    \
        }                                                                                                    \
        free(t1); free(ut); free(w); free(kw);                                                               \
        return 0;                                                                                            \
    }                                                                                                        \
    void tv_grad_##SUF(const T* x, T* out, int adjoint, int D, int64_t n0, int64_t n1, int64_t n2) {          \
        if (adjoint) grad_adjoint_##SUF(x, out, D, n0, n1, n2); else grad_apply_##SUF(x, out, D, n0, n1, n2); \
    }

DEFINE_TV(float, f32)
DEFINE_TV(double, f64)

void tv_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#endif
}

int tv_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
