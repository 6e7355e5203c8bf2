/*
 * pyxu_b200 -- C ABI of the B200-native hot path (libpyxu_b200.so).
 *
 * Drop-in boundary for the inner loop of Pyxu's PD3O / CondatVu / PGD solvers on stencil-based
 * imaging problems.  Plain pointers and sizes only: every array argument is a DEVICE pointer
 * (what a DLPack capsule's `data + byte_offset` holds), C-contiguous, dtype given by
 * `dtype` (PXB_F32 / PXB_F64); `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *
 * Every entry point returns 0 on success, a negative PXB_E* code otherwise; pxb_last_error()
 * returns a thread-local human-readable message.  No call synchronises the device.
 *
 * Array convention (same as the reference): an array of shape (..., N) with N = prod(arg_shape)
 * is seen as (batch, n0, n1, n2) with batch = prod(leading dims); arg_shapes of rank < 3 are
 * left-padded with 1 (kernel extent 1, center 0, mode constant along the padded axes).
 *
 * Each declaration cites the reference interface it replaces (file:line in AdriaJ/pyxu).
 */
#ifndef PYXU_B200_H
#define PYXU_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PXB_ABI_VERSION 3

enum pxb_dtype { PXB_F32 = 0, PXB_F64 = 1 };

/* boundary modes: numpy.pad names used by Pad/Stencil (reference: src/pyxu/operator/linop/pad.py:158-172) */
enum pxb_mode { PXB_CONSTANT = 0, PXB_WRAP = 1, PXB_REFLECT = 2, PXB_SYMMETRIC = 3, PXB_EDGE = 4 };

enum pxb_error {
    PXB_OK = 0,
    PXB_EINVAL = -1,   /* bad argument (message says which) */
    PXB_ECUDA = -2,    /* CUDA runtime error (message carries cudaGetErrorString) */
    PXB_ENOSUP = -3    /* valid request outside the compiled kernel envelope */
};

#define PXB_MAX_DIRS 3   /* directions of a Gradient stack (<= spatial rank) */
#define PXB_MAX_GTAP 16  /* taps of one 1-D derivative kernel carried by value */

/* ------------------------------------------------------------------------------------------ */
/* Domain decomposition along axis 0 (z-slabs).  The arrays handed to a call hold the planes    */
/* this rank owns; `halo_lo` / `halo_hi` planes of valid neighbour data sit immediately before  */
/* / after them in memory when the corresponding side is "open".  A closed side is a true       */
/* domain boundary and gets the boundary mode.  Single-GPU: both closed (all zeros).            */
/* ------------------------------------------------------------------------------------------ */
typedef struct pxb_slab {
    int32_t open_lo; /* 1: planes [-halo, 0) hold the lower neighbour's data */
    int32_t open_hi; /* 1: planes [n0, n0+halo) hold the upper neighbour's data */
    int32_t halo;    /* planes allocated on EACH side of the owned planes (0 on a single GPU).  Arrays are
                        (halo + n0 + halo, n1, n2) per component; pointers passed to a call address owned
                        plane 0.  halo > 0 requires batch == 1. */
    int32_t plane_alloc; /* planes allocated per component (stride between components / batch items, in planes);
                            0 means n0 + 2*halo.  Lets a call address a sub-range of planes of a larger slab
                            (boundary planes first, interior while the halo exchange is in flight). */
} pxb_slab;

/* ------------------------------------------------------------------------------------------ */
/* Stencil / Convolve  (reference: src/pyxu/operator/linop/stencil/stencil.py:356-461,         */
/* _stencil.py:139-198).  One dense correlation kernel of extent ksize[] whose entry            */
/* center[] sits on the output sample; out-of-domain reads follow mode[] per axis.              */
/* Separable stencils are issued by the host as one call per axis with ksize = 1 elsewhere.     */
/* ------------------------------------------------------------------------------------------ */
typedef struct pxb_stencil_desc {
    int32_t dtype;
    int32_t _pad;
    int64_t batch;
    int64_t shape[3];
    int32_t ksize[3];
    int32_t center[3];
    int32_t mode[3];
    pxb_slab slab;
    const void* coef; /* DEVICE pointer: prod(ksize) coefficients of `dtype`, C-order */
} pxb_stencil_desc;

/* y = S x  (Stencil.apply, stencil.py:441-450)  /  x = S^T y  (Stencil.adjoint, stencil.py:452-461).
 * `in` and `out` must not alias. */
int pxb_stencil_apply(const pxb_stencil_desc* d, const void* in, void* out, void* stream);
int pxb_stencil_adjoint(const pxb_stencil_desc* d, const void* in, void* out, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Tiled 2-D stencil for 'constant' boundaries: the whole correlation over the last two axes in ONE pass, also when */
/* the kernel is separable (the reference chains one 1-D stencil per axis, stencil.py:497-538, 441-450), with the    */
/* input window of each tile staged in shared memory by TMA (out-of-image samples zero-filled == numpy.pad           */
/* 'constant', pad.py:252-258).  Arrays are (nimg, n1, n2); nimg collects every leading dimension.                  */
/*   out[i] = alpha * S(in)[i] + beta * add[i % add_period]          (add nullable; add_period <= 0: same size)     */
/* S: dense == 0: taps coef1 (ksize[0], along axis n1) and coef2 (ksize[1], along n2), entry center[] on the output  */
/*    dense == 1: ksize[0] x ksize[1] coefficients of `dtype` at DEVICE pointer `coef`, C-order.                     */
/* The transpose of such an S is the same call with both factors reversed and center = ksize - 1 - center            */
/* (stencil.py:452-461 with zero padding).  Returns PXB_ENOSUP outside the envelope (more than 16 row taps; more     */
/* than 13 (fp32) / 11 (fp64) column taps; last axis not a multiple of 4 / 2 samples; unaligned arrays).             */
/* ------------------------------------------------------------------------------------------ */
typedef struct pxb_stencil2d {
    int32_t dtype;
    int32_t dense;
    int64_t nimg;
    int64_t shape[2];
    int32_t ksize[2];
    int32_t center[2];
    double coef1[16];
    double coef2[16];
    const void* coef;
    double alpha, beta;
    const void* add;
    int64_t add_period;
    /* Input images of another extent than the output (`shape`), and where the output grid sits in them: output sample    */
    /* (y, x) correlates the input samples (y + origin[0] - center[0] + q1, x + origin[1] - center[1] + q2); samples        */
    /* outside the input are zeros.  in_shape = {0, 0}: same extent, origin ignored.  in_shape[1] must be a multiple of     */
    /* 4 (fp32) / 2 (fp64).  This is how the folding boundary modes run: S o Pad reads the padded array (pxb_pad2d) and      */
    /* writes the trimmed one; S^T o Trim^T reads the array and writes the padded extent, which pxb_pad2d_adjoint folds.    */
    int64_t in_shape[2];
    int32_t origin[2];
} pxb_stencil2d;
int pxb_stencil2d_apply(const pxb_stencil2d* d, const void* in, void* out, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Pad over the two trailing axes and its transpose (reference: src/pyxu/operator/linop/pad.py:236-375), the pieces of  */
/* Stencil = Trim o S0 o Pad (stencil.py:76-84) that the tiled stencil does not absorb when a mode folds.               */
/*   ext is (nimg, ext_shape[0], ext_shape[1]); the array sits at rows [org[0], org[0] + shape[0]), columns             */
/*   [org[1], org[1] + shape[1]) of it and is extended by lo[] / hi[] samples before / after (org >= lo; anything        */
/*   beyond the padded extent is filler: alignment of the rows to 16 bytes).                                             */
/*   pxb_pad2d:          ext[e1 + org[0], e2 + org[1]] = in[m1(e1), m2(e2)]   for -lo <= e < shape + hi, else 0          */
/*                       (m = numpy.pad's index map of mode[]; 'constant' cells are 0)                                   */
/*   pxb_pad2d_adjoint:  out[t1, t2] = alpha * sum over the cells (e1, e2) of the padded extent with m1(e1) = t1,        */
/*                       m2(e2) = t2 of ext[e1 + org[0], e2 + org[1]]   +   beta * add[i % add_period]                   */
/* ------------------------------------------------------------------------------------------ */
typedef struct pxb_pad2d_desc {
    int32_t dtype;
    int32_t _pad;
    int64_t nimg;
    int64_t shape[2];
    int64_t ext_shape[2];
    int32_t org[2];
    int32_t lo[2];
    int32_t hi[2];
    int32_t mode[2];
} pxb_pad2d_desc;
int pxb_pad2d(const pxb_pad2d_desc* d, const void* in, void* ext, void* stream);
int pxb_pad2d_adjoint(const pxb_pad2d_desc* d, const void* ext, void* out, double alpha, double beta, const void* add,
                      int64_t add_period, void* stream);


/* ------------------------------------------------------------------------------------------ */
/* Gradient stack: ndir first-order 1-D derivative stencils, direction k acting along           */
/* axis[k] (reference: src/pyxu/operator/linop/diff.py:1113-1265 Gradient,                      */
/* :952-1056 _stack_diff_ops, :157-261 finite-difference coefficients).                         */
/* Output layout (batch, ndir, n0, n1, n2) == vstack of the per-direction operators.            */
/* ------------------------------------------------------------------------------------------ */
typedef struct pxb_grad_desc {
    int32_t dtype;
    int32_t ndir;
    int64_t batch;
    int64_t shape[3];
    int32_t mode[3];              /* per AXIS */
    int32_t axis[PXB_MAX_DIRS];   /* per direction: axis in [0,3) of the left-padded shape */
    int32_t ntap[PXB_MAX_DIRS];
    int32_t center[PXB_MAX_DIRS];
    double coef[PXB_MAX_DIRS][PXB_MAX_GTAP];
    pxb_slab slab;
} pxb_grad_desc;

int pxb_gradient_apply(const pxb_grad_desc* d, const void* x, void* z, void* stream);   /* z = K x   */
int pxb_gradient_adjoint(const pxb_grad_desc* d, const void* z, void* x, void* stream); /* x = K^T z */

/* ------------------------------------------------------------------------------------------ */
/* Proximal maps of the primal term g (reference: src/pyxu/operator/func/norm.py:47-52 L1Norm, */
/* :100-104 SquaredL2Norm, :400-403 PositiveL1Norm; func/indicator.py:203-206 PositiveOrthant, */
/* :58-69 LInfinityBall; ScaleRule abc/arithmetic.py:182).  prox_{tau*g}.                      */
/* ------------------------------------------------------------------------------------------ */
enum pxb_prox_kind {
    PXB_PROX_NONE = 0,  /* g = 0                      : v                                   */
    PXB_PROX_POS = 1,   /* g = i_{x>=0}               : max(v, 0)                           */
    PXB_PROX_BOX = 2,   /* g = i_{p0<=x<=p1}          : clip(v, p0, p1)                     */
    PXB_PROX_L1 = 3,    /* g = p0*||x||_1             : soft(v, p0*tau)                     */
    PXB_PROX_POSL1 = 4, /* g = p0*(||x||_1 + i_+)     : max(v - p0*tau, 0)                  */
    PXB_PROX_SQL2 = 5   /* g = p0*||x||_2^2           : v / (2*p0*tau + 1)                  */
};
typedef struct pxb_prox_spec {
    int32_t kind;
    int32_t _pad;
    double p0, p1;
} pxb_prox_spec;

/* out[i] = prox_{tau g}( a*x[i] + b*y[i % ny] + c*z[i % nz] );  y, z may be NULL (term dropped).
 * One pass: the linear combination every solver forms before calling g.prox
 * (pds.py:431-434, :750-753; pgd.py:179-191).  out may alias x. */
int pxb_prox_lincomb(int dtype, const pxb_prox_spec* g, double tau, int64_t n, void* out,
                     double a, const void* x, double b, const void* y, int64_t ny,
                     double c, const void* z, int64_t nz, void* stream);

/* out[i] = a*x[i] + b*y[i % ny] + c*z[i % nz]  (ScaleRule / ArgShiftRule / relaxation algebra). */
int pxb_lincomb(int dtype, int64_t n, void* out, double a, const void* x, double b, const void* y,
                int64_t ny, double c, const void* z, int64_t nz, void* stream);

/* The factor of a separable 3-D stencil along the SLOWEST axis ('constant' boundary) as one streaming pass:
 *   out[q, :, :] = sum_j coef[j] * in[q + j - c0, :, :]      (coef: HOST array of k0 doubles)
 * (reference: the axis-0 link of the chain stencil.py:497-538 builds for separable kernels).  Slab cuts: planes beyond an
 * open side are read from the ghost planes (slab->halo >= the reach on that side).  Envelope: 2..9 taps, plane size a
 * multiple of 4 (fp32) / 2 (fp64) samples, 16-byte aligned arrays; PXB_ENOSUP otherwise (use pxb_stencil_apply). */
int pxb_stencil_axis0_apply(int dtype, int64_t batch, const int64_t* shape, const pxb_slab* slab, int k0, int c0,
                            const double* coef, const void* in, void* out, void* stream);
/* The same streaming pass with a FOLDING boundary mode along axis 0 (single-domain arrays), both directions of
 * Trim o S0 o Pad (stencil.py:76-84, 130-146):
 *   adjoint == 0:  out[q] = sum_j coef[j] * in[m(q + j - c0)]            (m: numpy.pad's index map of `mode`, pad.py:252-302)
 *   adjoint != 0:  coef / c0 hold the REVERSED taps and the mirrored centre (k0 - 1 - centre); out[t] = sum over the padded
 *                  coordinates e with m(e) = t of  sum_j coef[j] * in[e + j - c0]  (in zero outside the array): the
 *                  transpose, Pad^T o S0^T o Trim^T (pad.py:307-375).
 * Envelope as pxb_stencil_axis0_apply; PXB_ENOSUP otherwise (use pxb_stencil_apply / pxb_stencil_adjoint). */
int pxb_stencil_axis0_fold(int dtype, int64_t batch, const int64_t* shape, int k0, int c0, const double* coef, int mode,
                           int adjoint, const void* in, void* out, void* stream);

/* A separable 3-D stencil ('constant' boundaries) in ONE pass over HBM where the reference chains three 1-D stencils
 * (stencil.py:497-538): thread blocks march along axis 0 with the in-plane-filtered planes in a register ring.
 *   out = alpha * (S_0 S_1 S_2)(in) + beta * add[i % add_period]     (add: dense (batch, n0, n1, n2), nullable)
 * coef0/1/2: the 1-D factors along axes 0/1/2 (HOST values), center[] their entries on the output sample.
 * Envelope: 3, 5, 7 or 9 taps along axis 0, <= 16 along axis 1, <= 13 (fp32) / 11 (fp64) along axis 2, last axis a
 * multiple of 4 / 2 samples, 16-byte aligned arrays; PXB_ENOSUP otherwise (use pxb_stencil_axis0_apply +
 * pxb_stencil2d_apply).  Slab cuts: as pxb_stencil_axis0_apply. */
typedef struct pxb_stencil3d {
    int32_t dtype;
    int32_t _pad;
    int64_t batch;
    int64_t shape[3];
    int32_t ksize[3];
    int32_t center[3];
    double coef0[16];
    double coef1[16];
    double coef2[16];
    double alpha, beta;
    const void* add;
    int64_t add_period;
    pxb_slab slab;
} pxb_stencil3d;
int pxb_stencil3d_apply(const pxb_stencil3d* d, const void* in, void* out, void* stream);
/* Which kernel serves pxb_stencil3d_apply: 0 = automatic (K x K x K taps, K in {3, 5, 7, 9}, centred along the rows and a dense or
 * absent epilogue operand: the fully unrolled instances; anything else: the general marching kernel), 1 = the general kernel
 * always.  No reference counterpart: for A/B measurements and tests. */
int pxb_set_stencil3d_path(int path);

/* A DENSE (full-rank) 3-D stencil ('constant' boundaries; e.g. a measured 7x7x7 PSF) in ONE pass over HBM.  The reference takes
 * any dense kernel and evaluates it sample by sample over the padded array (stencil.py:356-461, _stencil.py:232-305); here
 * thread blocks march along axis 0 and scatter every staged input plane into the register accumulators of the output planes it
 * contributes to, the coefficients being constant-bank operands of the FMAs (k0*k1*k2 FMAs per sample, 8 B/voxel in fp32).
 *   out = alpha * S(in) + beta * add[i % add_period]     (add: dense (batch, n0, n1, n2), nullable)
 * coef: k0*k1*k2 HOST values, row-major; center[]: the kernel's entry on the output sample.  The adjoint of a zero-padded
 * correlation is the correlation with the reversed kernel and the mirrored centre (the caller passes those).
 * Envelope: every extent <= 7 and the kernel filling at least half of the enclosing cube of 3, 5 or 7 taps; last axis a multiple
 * of 4 / 2 samples, 16-byte aligned arrays; PXB_ENOSUP otherwise (use pxb_stencil_apply, or pxb_stencil2d_apply per kernel
 * plane).  Slab cuts: as pxb_stencil_axis0_apply. */
typedef struct pxb_stencil3d_dense {
    int32_t dtype;
    int32_t _pad;
    int64_t batch;
    int64_t shape[3];
    int32_t ksize[3];
    int32_t center[3];
    const double* coef;
    double alpha, beta;
    const void* add;
    int64_t add_period;
    pxb_slab slab;
} pxb_stencil3d_dense;
int pxb_stencil3d_dense_apply(const pxb_stencil3d_dense* d, const void* in, void* out, void* stream);

/* One accelerated proximal-gradient (FISTA) iteration on f = alpha_f*||A x + shift||^2, g pointwise, A such a stencil
 * (reference: src/pyxu/opt/solver/pgd.py:173-191), as TWO tiled passes instead of five:
 *   which == 0:  out = r = d.alpha * A((1+a) x - a x_prev) + d.beta * d.add      (d describes A; the extrapolated point
 *                y is formed in shared memory from the two input windows, never written)
 *   which == 1:  out = x_new = prox_{tau g}((1+a) x - a x_prev + d.alpha * A^T r) (d describes A^T, d.alpha = -tau);
 *                out may be the x_prev buffer.  norms (nullable): per row += { sum (x_new - x)^2, sum x^2 }
 *                with row = image / imgs_per_row  (RelError[x], stop.py:353-382).
 * Envelope as pxb_stencil2d_apply. */
typedef struct pxb_fista_step {
    const void* x;
    const void* x_prev;
    const void* r;       /* which == 1 */
    double a;            /* momentum a_k = k / (k + 1 + d);  0: plain proximal gradient */
    double tau;
    pxb_prox_spec g;
    double* norms;
    int64_t imgs_per_row;
} pxb_fista_step;
int pxb_stencil2d_fista(const pxb_stencil2d* d, const pxb_fista_step* f, int which, void* out, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Dual term h (reference: src/pyxu/operator/func/norm.py:352-364 L21Norm.prox, :47-52 L1Norm;  */
/* fenchel_prox abc/operator.py:906-944).  Arrays are (outer, group, inner); the l2 norm runs   */
/* over `group` (for TV: outer = batch, group = ndir, inner = voxels).                          */
/* ------------------------------------------------------------------------------------------ */
enum pxb_dual_kind { PXB_DUAL_NONE = 0, PXB_DUAL_L21 = 1, PXB_DUAL_L1 = 2 };

/* out = prox_{tau*lam*||.||_{2,1}}(x)                (L21Norm.prox through ScaleRule) */
int pxb_prox_l21(int dtype, int64_t outer, int64_t group, int64_t inner, double lam, double tau,
                 const void* x, void* out, void* stream);

/* Dual update of CondatVu / PD3O (pds.py:437-441, :756-760):
 *   p = z + sigma*t ;  z <- (1-rho)*z + rho*prox_{sigma h*}(p),  h = lam*L21 | lam*L1.
 * norms (nullable): per `outer` row, += { sum (z_new-z_old)^2, sum z_old^2 }  (RelError[z]). */
int pxb_dual_update(int dtype, int kind, int64_t outer, int64_t group, int64_t inner, double lam,
                    double sigma, double rho, void* z, const void* t, double* norms, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Fused primal / dual half-iterations for h o K with K a Gradient stack (TV-type problems).    */
/* One pass over every voxel each.                                                              */
/* ------------------------------------------------------------------------------------------ */
enum pxb_algo { PXB_PD3O = 0, PXB_CV = 1 };

enum pxb_fterm_kind {
    PXB_F_NONE = 0,    /* f = 0 */
    PXB_F_SQL2 = 1,    /* f = alpha*||x + shift||^2, grad = 2*alpha*(x + shift)  (SquaredL2Norm + ArgShift + Scale) */
    PXB_F_GRADARR = 2  /* CV only: grad f(x_k) precomputed in `garr` (non-local f, e.g. blur data term) */
};
typedef struct pxb_fterm {
    int32_t kind;
    int32_t _pad;
    double alpha;
    const void* shift;    /* DEVICE, nullable (zero shift) */
    int64_t shift_period; /* shift[i % shift_period] (broadcast over stacked problems) */
    const void* garr;     /* DEVICE, PXB_F_GRADARR */
} pxb_fterm;

typedef struct pxb_pds_params {
    double tau, sigma, rho;
    pxb_prox_spec g;
    pxb_fterm f;
    int32_t hkind; /* pxb_dual_kind */
    int32_t _pad;
    double lam;
} pxb_pds_params;

/* Primal half-step, per voxel s (K^T z gathered in-kernel from z unless `ktz` is given):
 *  PD3O (pds.py:747-761):  x = prox_g(u - tau*K^T z);  ut = x - tau*grad f(x);
 *                          w = x + ut - u;  u <- (1-rho)*u + rho*ut
 *        `xu` holds u (in/out), `x_out` receives x, `w` receives w.
 *  CV   (pds.py:429-442):  xt = prox_g(x - tau*grad f(x) - tau*K^T z);  w = 2*xt - x;
 *                          x <- rho*xt + (1-rho)*x
 *        `xu` holds x (in/out), `x_out` unused (may be NULL), `w` receives 2*xt - x.
 *  norms (nullable): per batch row += { sum (x_new-x_old)^2, sum x_old^2 } where x_old is the
 *  previous content of x_out (PD3O) / xu (CV)  -> RelError[x] without an extra pass. */
int pxb_pds_primal(int algo, const pxb_grad_desc* K, const pxb_pds_params* p, void* xu,
                   const void* z, const void* ktz, void* x_out, void* w, double* norms, void* stream);

/* Dual half-step:  z <- (1-rho)*z + rho*prox_{sigma h*}(z + sigma*K w)  with K w gathered in-kernel.
 * norms as in pxb_dual_update. */
int pxb_pds_dual(const pxb_grad_desc* K, const pxb_pds_params* p, const void* w, void* z,
                 double* norms, void* stream);

/* One whole iteration in ONE pass (primal half-step + dual half-step + both RelError norms), out of place:
 *   reads  xu_in (PD3O: u, CV: x), z_in      -- left untouched (they become the next call's output buffers)
 *   writes xu_out, z_out, and for PD3O x_out (nullable: x is then not materialised)
 * Same algebra as pxb_pds_primal followed by pxb_pds_dual (pds.py:429-442, :747-761); w never leaves the SM.
 * norms_x / norms_z (nullable): per batch row += { sum (new-old)^2, sum old^2 } for x (PD3O: against the previous
 * content of x_out) and z.
 * Envelope: K a 2- or 3-direction first-order Gradient (any boundary mode on single-domain arrays, 'constant' on slabs;
 * see pxb_set_iter_modes), 16-byte aligned arrays, last
 * axis a multiple of 4 (fp32) / 2 (fp64) samples, h = lam*L21 | lam*L1, f pointwise (PD3O) or any (CV, garr).
 * Returns PXB_ENOSUP outside the envelope: callers fall back to the two-pass form.
 * With an open slab side the ghost planes of xu_in, z_in (and of the shift array) must hold the neighbour's data. */
int pxb_pds_iter(int algo, const pxb_grad_desc* K, const pxb_pds_params* p, const void* xu_in, const void* z_in,
                 void* xu_out, void* z_out, void* x_out, double* norms_x, double* norms_z, void* stream);
/* Which implementation pxb_pds_iter uses: 0 = automatic (3-D: TMA-staged pipeline, 2-D: direct loads),
 * 1 = direct-load form only, 2 = TMA form only (PXB_ENOSUP when it does not apply).  Environment variable
 * PXB_TV_ITER=direct|tma sets the initial value.  For A/B measurements and tests. */
int pxb_set_iter_path(int path);
/* Folding boundary modes (wrap / reflect / symmetric / edge; pad.py:252-302) inside pxb_pds_iter: 1 (default) = the
 * single-kernel forms serve them too (the samples within two of a folding face, and the out-of-domain rim of the w
 * tiles, go through the per-sample boundary map / pre-image gather; single-domain arrays only: slabs answer PXB_ENOSUP),
 * 0 = PXB_ENOSUP for every folding mode, so that callers take pxb_pds_primal + pxb_pds_dual (A/B measurements, tests),
 * -1 = back to the initial value (environment variable PXB_TV_ITER_MODES=0|1, else 1). */
int pxb_set_iter_modes(int on);
/* Up to `n` iterations queued back to back on the stream, the stopping rule tested ON THE DEVICE after each one, so that the
 * host is out of the loop (reference loop: abc/solver.py:588-652 tests the criterion on the host before every step; for a
 * 512x512 problem an iteration is ~3 us of GPU work against ~30 us of host turnaround).
 *   (xu_a, z_a) holds the current iterate; iteration i reads pair (i even ? a : b) and writes the other one.
 *   x           PD3O: x, rewritten by every iteration (needed for RelError[x]; nullable when eps_x <= 0).
 *   norms       DEVICE double[n][2][batch][2], zeroed by the caller: iteration i accumulates its RelError sums
 *               {sum (new-old)^2, sum old^2} for x ([i][0]) and z ([i][1]) there -- the host replays its criterion / history
 *               from them afterwards, so the log is the same as with one launch per iteration.
 *   rule        RelError thresholds as in stop.py:353-382: a variable's test holds when sqrt(num) <= eps*sqrt(den) for every
 *               (all = 1) / any (all = 0) batch row; `table` bit (2*px + pz) tells whether the composed criterion stops for
 *               outcomes (px, pz) -- the host builds it from its criterion tree with the non-norm leaves (MaxIter ...) false.
 *   ctl         DEVICE pxb_iter_ctl, zeroed by the caller before the first batch: `done` counts the iterations carried out
 *               (the iterate is in pair a when done is even), `stop` is raised by the iteration that met the rule; launches
 *               behind it return without touching anything.
 * rule == NULL (norms, ctl then unused): n plain iterations, for criteria only the host evaluates (MaxIter); x is then left alone.
 * Envelope and return codes as pxb_pds_iter (PXB_ENOSUP: nothing was launched). */
typedef struct pxb_stop_rule {
    double eps_x, eps_z;
    int32_t all_x, all_z;
    int32_t table;
    int32_t _pad;
} pxb_stop_rule;
typedef struct pxb_iter_ctl {
    int32_t stop;
    int32_t done;
    uint32_t ticket; /* internal: thread blocks of the running iteration that have finished */
    int32_t _pad;
} pxb_iter_ctl;
int pxb_pds_iter_n(int algo, const pxb_grad_desc* K, const pxb_pds_params* p, void* xu_a, void* z_a, void* xu_b, void* z_b,
                   void* x, double* norms, int n, const pxb_stop_rule* rule, pxb_iter_ctl* ctl, void* stream);
/* pxb_pds_iter on a z-slab with the halo exchange FUSED into the kernel (no reference counterpart: the reference distributes
 * through Dask arrays).  The thread blocks that produce the slab's first / last owned plane also store it -- u, z_0, z_1, z_2 of
 * the first plane, z_0 of the last one -- into the neighbours' ghost planes through peer memory (NVLink), then bump a counter in
 * the neighbour's memory; the blocks that read a ghost plane wait until the neighbour's counter says its previous iteration has
 * delivered (epoch * tiles per plane).  Those blocks are scheduled first.  All pointers are DEVICE addresses valid on this GPU
 * (peer mappings of the neighbours' allocations); a null neighbour side is a closed side of the volume.
 *   dn_u, dn_z    lower neighbour: the ghost plane ABOVE its last owned plane, in the buffers it will READ next iteration
 *                 (primal; z component 0, the other components dn_zvol elements apart)
 *   up_z0         upper neighbour: the ghost plane BELOW its first owned plane of z component 0
 *   dn_flag, up_flag   counters in the neighbours' memory (uint32, monotonically increasing over the solve)
 *   lo_wait, hi_wait   counters in this rank's memory which the lower / upper neighbour bumps
 *   epoch         iterations completed before this one (0 for the first: the initial ghost planes are exchanged by the host)
 * Envelope: as pxb_pds_iter for 3-D volumes (TMA form), batch 1.  PXB_ENOSUP: nothing was launched. */
typedef struct pxb_peer {
    void* dn_u;
    void* dn_z;
    int64_t dn_zvol;
    void* up_z0;
    uint32_t* dn_flag;
    uint32_t* up_flag;
    const uint32_t* lo_wait;
    const uint32_t* hi_wait;
    int64_t epoch;
} pxb_peer;
int pxb_pds_iter_p2p(int algo, const pxb_grad_desc* K, const pxb_pds_params* p, const void* xu_in, const void* z_in, void* xu_out,
                     void* z_out, void* x_out, double* norms_x, double* norms_z, const pxb_peer* peer, void* stream);
/* same, with the number of planes one thread block marches through fixed by the caller (tuning / tests) */
int pxb_pds_iter_chunked(int algo, const pxb_grad_desc* K, const pxb_pds_params* p, const void* xu_in, const void* z_in,
                         void* xu_out, void* z_out, void* x_out, double* norms_x, double* norms_z, int chunk, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Stopping-criterion norms (reference: src/pyxu/opt/stop.py:273-284 AbsError, :353-382         */
/* RelError).  Per row r of a (rows, n) array:  out[2r] += sum (x-y)^2 (y NULL -> sum x^2),     */
/* out[2r+1] += sum y^2.  `out` is a DEVICE double[2*rows] the caller zeroes.                   */
/* ------------------------------------------------------------------------------------------ */
int pxb_sqnorms(int dtype, int64_t rows, int64_t n, const void* x, const void* y, double* out, void* stream);

/* misc */
int pxb_abi_version(void);
/* Lets kernels launched on the calling thread's current device reach memory of `peer_device` (needed once per pair before
 * pxb_pds_iter_p2p is handed pointers into a neighbour's allocation).  PXB_ENOSUP when the two GPUs have no peer path. */
int pxb_enable_peer_access(int peer_device);
const char* pxb_last_error(void);
/* number of kernels this library has launched in the calling process (bench's gpu_launches) */
int64_t pxb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* PYXU_B200_H */
