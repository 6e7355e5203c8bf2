"""
CPU oracle for the pyxu_b200 hot path  --  TEST INFRASTRUCTURE ONLY.

This module is a plain-NumPy restatement of the algorithms the reference (AdriaJ/pyxu) runs on the
path named by BASELINE.json's north_star: Stencil/Convolve, Gradient (finite differences) and its
adjoint, the L1 / L21 / PositiveOrthant / box proximal maps, SquaredL2Norm, the PD3O / CondatVu /
PGD iterations and the RelError / AbsError / MaxIter stopping rules.

Who may import it: `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / reference
arm.  Nothing under `pyxu_b200/` imports it; the product path is CUDA-only.

Parity status: PINNED.  Every function below is checked (tests/test_oracle_golden.py) against
  * the known-answer vectors printed in the reference's docstrings
    (reference: src/pyxu/operator/linop/stencil/stencil.py:172-342, :829-860; pad.py:27-55),
  * the reference's own test recipe (scipy.ndimage.correlate on a Pad()-extended array,
    reference: src/pyxu_tests/operator/linop/test_stencil.py:143-181), and
  * fixtures produced by importing the real reference in the build container
    (tests/golden/make_golden.py -> tests/golden/*.npz).

All arrays follow the reference convention: shape (..., N) with N = prod(arg_shape), C-order.
"""
import itertools
import math

import numpy as np

MODES = ("constant", "wrap", "reflect", "symmetric", "edge")


# ----------------------------------------------------------------------------------------------
# Pad / Trim  (reference: src/pyxu/operator/linop/pad.py:236-375)
# ----------------------------------------------------------------------------------------------
def _canon_mode(mode, ndim):
    if isinstance(mode, str):
        mode = (mode,) * ndim
    mode = tuple(m.strip().lower() for m in mode)
    assert len(mode) == ndim and set(mode) <= set(MODES)
    return mode


def _axis_source_index(n, lhs, rhs, mode):
    """For one axis: index map padded-position -> source position (or -1 for 'zero').

    Restates the slice arithmetic of Pad.apply() (pad.py:252-302) as an explicit index table.
    """
    idx = np.arange(-lhs, n + rhs)
    if mode == "constant":
        src = np.where((idx >= 0) & (idx < n), idx, -1)
    elif mode == "wrap":
        assert max(lhs, rhs) <= n  # pad.py:219-229
        src = idx % n
    elif mode == "reflect":
        assert max(lhs, rhs) <= n - 1
        src = np.where(idx < 0, -idx, np.where(idx >= n, 2 * (n - 1) - idx, idx))
    elif mode == "symmetric":
        assert max(lhs, rhs) <= n
        src = np.where(idx < 0, -idx - 1, np.where(idx >= n, 2 * n - 1 - idx, idx))
    elif mode == "edge":
        src = np.clip(idx, 0, n - 1)
    else:
        raise ValueError(mode)
    return src


def pad_apply(x, arg_shape, pad_width, mode):
    """(..., prod(arg_shape)) -> (..., prod(pad_shape)).  Pad.apply(), pad.py:236-305."""
    arg_shape = tuple(arg_shape)
    D = len(arg_shape)
    mode = _canon_mode(mode, D)
    sh = x.shape[:-1]
    out = x.reshape(*sh, *arg_shape)
    for ax in range(D):  # separable: each axis is extended from the already-extended earlier ones
        lhs, rhs = pad_width[ax]
        src = _axis_source_index(arg_shape[ax], lhs, rhs, mode[ax])
        taken = np.take(out, np.maximum(src, 0), axis=len(sh) + ax)
        mask_shape = [1] * taken.ndim
        mask_shape[len(sh) + ax] = -1
        out = taken * (src >= 0).reshape(mask_shape).astype(x.dtype)
    return out.reshape(*sh, -1)


def pad_adjoint(y, arg_shape, pad_width, mode):
    """(..., prod(pad_shape)) -> (..., prod(arg_shape)).  Pad.adjoint(), pad.py:307-375.

    The adjoint folds every padded sample back onto the source sample it was copied from.
    """
    arg_shape = tuple(arg_shape)
    D = len(arg_shape)
    mode = _canon_mode(mode, D)
    pad_shape = tuple(n + l + r for n, (l, r) in zip(arg_shape, pad_width))
    sh = y.shape[:-1]
    out = y.reshape(*sh, *pad_shape)
    for ax in range(D):
        lhs, rhs = pad_width[ax]
        n = arg_shape[ax]
        src = _axis_source_index(n, lhs, rhs, mode[ax])
        moved = np.moveaxis(out, len(sh) + ax, 0)
        acc = np.zeros((n,) + moved.shape[1:], dtype=y.dtype)
        valid = src >= 0
        np.add.at(acc, src[valid], moved[valid])
        out = np.moveaxis(acc, 0, len(sh) + ax)
    return out.reshape(*sh, -1)


# ----------------------------------------------------------------------------------------------
# Zero-boundary stencil  (reference: src/pyxu/operator/linop/stencil/_stencil.py:232-305)
# numba.stencil(mode="constant", cval=0): outputs whose neighbourhood is not fully inside the
# array are set to 0; elsewhere out[i] = sum_q k[q] * a[i - c + q].
# ----------------------------------------------------------------------------------------------
def stencil0(a, kernel, center):
    """a: (S, N_1..N_D); kernel: (K_1..K_D); center: (D,).  Returns array like `a`."""
    D = kernel.ndim
    out = np.zeros_like(a)
    shape = a.shape[1:]
    lo = [int(c) for c in center]
    hi = [int(k - 1 - c) for k, c in zip(kernel.shape, center)]
    if any(n - l - h <= 0 for n, l, h in zip(shape, lo, hi)):
        return out
    core = tuple(slice(l, n - h) for n, l, h in zip(shape, lo, hi))
    acc = np.zeros((a.shape[0],) + tuple(n - l - h for n, l, h in zip(shape, lo, hi)), dtype=a.dtype)
    for q in itertools.product(*[range(k) for k in kernel.shape]):
        w = kernel[q]
        if w == 0:
            continue
        sl = tuple(slice(qq, qq + n - l - h) for qq, n, l, h in zip(q, shape, lo, hi))
        acc += w * a[(slice(None),) + sl]
    out[(slice(None),) + core] = acc
    return out


# ----------------------------------------------------------------------------------------------
# Stencil / Convolve  (reference: src/pyxu/operator/linop/stencil/stencil.py:356-461, :497-576, :794-887)
# ----------------------------------------------------------------------------------------------
class Stencil:
    def __init__(self, arg_shape, kernel, center, mode="constant", dtype=np.float64):
        if not isinstance(arg_shape, (tuple, list)):
            arg_shape = (arg_shape,)
        self.arg_shape = tuple(int(n) for n in arg_shape)
        D = len(self.arg_shape)
        assert len(center) == D
        # canonical representation (stencil.py:497-538)
        if isinstance(kernel, np.ndarray):
            assert kernel.ndim == D
            self._k = [np.asarray(kernel, dtype=dtype)]
            self._c = [np.array(center, dtype=int)]
            self.separable = False
        else:
            assert len(kernel) == D
            self._k, self._c = [], []
            for i in range(D):
                sh = [1] * D
                sh[i] = -1
                self._k.append(np.asarray(kernel[i], dtype=dtype).reshape(sh))
                c = np.zeros(D, dtype=int)
                c[i] = center[i]
                self._c.append(c)
            self.separable = True
        self.mode = _canon_mode(mode, D)
        self.dtype = np.dtype(dtype)
        # pad widths (stencil.py:540-561)
        pw = []
        for i in range(D):
            if not self.separable:
                c, n = int(self._c[0][i]), self._k[0].shape[i]
            else:
                c, n = int(self._c[i][i]), self._k[i].size
            p = max(c, n - c - 1) if self.mode[i] == "constant" else n - 1
            pw.append((p, p))
        self.pad_width = tuple(pw)
        self.pad_shape = tuple(n + 2 * p for n, (p, _) in zip(self.arg_shape, pw))
        # adjoint kernels: flipped kernel, mirrored center (stencil.py:563-576)
        self._k_bw = [np.flip(k) for k in self._k]
        self._c_bw = [np.array(k.shape) - c - 1 for k, c in zip(self._k, self._c)]
        self.dim = self.codim = int(np.prod(self.arg_shape))

    # Lipschitz bound (stencil.py:639-656, pad.py:377-394, Trim -> 1)
    @property
    def lipschitz(self):
        full = 1
        for k in self._k:
            full = full * k
        L_st = np.abs(np.asarray(full)).sum()
        L_pad = 1.0
        for n, m, (l, r) in zip(self.arg_shape, self.mode, self.pad_width):
            if m == "constant":
                L = 1
            elif m in ("wrap", "symmetric"):
                L = np.sqrt(1 + np.ceil((l + r) / n))
            elif m == "reflect":
                L = np.sqrt(1 + np.ceil((l + r) / (n - 2)))
            else:
                L = np.sqrt(1 + max(l, r))
            L_pad *= L
        return float(L_st * L_pad)

    def _chain(self, x, ks, cs):
        for k, c in zip(ks, cs):
            x = stencil0(x, k, c)
        return x

    def _trim(self, y):
        sl = tuple(slice(p, p + n) for n, (p, _) in zip(self.arg_shape, self.pad_width))
        return y[(slice(None),) + sl]

    def _fw_bw(self):
        return (self._k, self._c), (self._k_bw, self._c_bw)

    def apply(self, arr):
        fw, _ = self._fw_bw()
        arr = np.asarray(arr, dtype=self.dtype)
        sh = arr.shape[:-1]
        x = pad_apply(arr, self.arg_shape, self.pad_width, self.mode).reshape(-1, *self.pad_shape)
        y = self._chain(x, *fw)
        return self._trim(y).reshape(*sh, -1)

    def adjoint(self, arr):
        _, bw = self._fw_bw()
        arr = np.asarray(arr, dtype=self.dtype)
        sh = arr.shape[:-1]
        x = np.zeros((int(np.prod(sh, dtype=int)),) + self.pad_shape, dtype=self.dtype)
        sl = tuple(slice(p, p + n) for n, (p, _) in zip(self.arg_shape, self.pad_width))
        x[(slice(None),) + sl] = arr.reshape(-1, *self.arg_shape)  # Trim.adjoint == zero-pad
        y = self._chain(x, *bw).reshape(*sh, -1)
        return pad_adjoint(y, self.arg_shape, self.pad_width, self.mode)

    __call__ = apply


class Convolve(Stencil):
    """Convolution == correlation with the flipped kernel / mirrored center (stencil.py:878-887)."""

    def _fw_bw(self):
        return (self._k_bw, self._c_bw), (self._k, self._c)


# ----------------------------------------------------------------------------------------------
# Finite differences / Gradient  (reference: src/pyxu/operator/linop/diff.py:157-261, :1113-1265)
# ----------------------------------------------------------------------------------------------
def fd_coefficients(order=1, scheme="forward", accuracy=1, sampling=1.0, dtype=np.float64):
    """Returns (coefs, center).  diff.py:215-258."""
    if scheme == "central":
        n = 2 * ((order + 1) // 2) - 1 + accuracy
        ids = np.arange(-(n // 2), n // 2 + 1, dtype=int)
    elif scheme == "forward":
        ids = np.arange(0, order + accuracy, dtype=int)
    elif scheme == "backward":
        ids = np.arange(-(order + accuracy) + 1, 1, dtype=int)
    else:
        raise ValueError(scheme)
    mat = np.vander(ids, increasing=True).T.astype(dtype)
    rhs = np.zeros(len(ids), dtype=dtype)
    rhs[order] = math.factorial(order)
    coefs = np.linalg.solve(mat, rhs)
    coefs /= sampling**order
    return coefs, int(ids.tolist().index(0))


class Gradient:
    """Stack of first-order partial derivatives, one separable Stencil per direction.

    Output layout (D, N_1..N_D) raveled, i.e. vstack of the per-direction operators
    (diff.py:952-1056, blocks.vstack).
    """

    def __init__(self, arg_shape, mode="constant", scheme="forward", accuracy=1, sampling=1.0, dtype=np.float64):
        self.arg_shape = tuple(int(n) for n in arg_shape)
        D = len(self.arg_shape)
        samp = (sampling,) * D if np.isscalar(sampling) else tuple(sampling)
        sch = (scheme,) * D if isinstance(scheme, str) else tuple(scheme)
        acc = (accuracy,) * D if np.isscalar(accuracy) else tuple(accuracy)
        self.ops = []
        for d in range(D):
            coefs, c = fd_coefficients(1, sch[d], acc[d], samp[d], dtype)
            kern = [np.array([1.0], dtype=dtype)] * D  # identity along the other axes (diff.py:715)
            kern[d] = coefs
            center = [0] * D
            center[d] = c
            self.ops.append(Stencil(self.arg_shape, kern, center, mode, dtype))
        self.dim = int(np.prod(self.arg_shape))
        self.codim = D * self.dim

    @property
    def lipschitz(self):
        # vstack rule: L = sqrt(sum_i L_i^2)  (reference: src/pyxu/operator/blocks.py, vstack)
        return float(np.sqrt(sum(op.lipschitz**2 for op in self.ops)))

    def apply(self, arr):
        return np.concatenate([op.apply(arr) for op in self.ops], axis=-1)

    def adjoint(self, arr):
        parts = np.split(arr, len(self.ops), axis=-1)
        out = self.ops[0].adjoint(parts[0])
        for op, p in zip(self.ops[1:], parts[1:]):
            out = out + op.adjoint(p)
        return out

    __call__ = apply


# ----------------------------------------------------------------------------------------------
# Functionals  (reference: src/pyxu/operator/func/norm.py, func/indicator.py, abc/operator.py:906-944)
# ----------------------------------------------------------------------------------------------
def l1_prox(x, tau):  # norm.py:47-52
    return np.fmax(0, np.fabs(x) - tau) * np.sign(x)


def l1_apply(x):
    return np.abs(x).sum(axis=-1, keepdims=True)


def positive_l1_prox(x, tau):  # norm.py:400-403
    return np.fmax(0, x - tau)


def positive_orthant_prox(x, tau=None):  # indicator.py:203-206
    return x.clip(0, None)


def linf_ball_prox(x, radius):  # indicator.py:58-69 with ord=inf: x - prox_{r l1}(x)  == clip
    return x - l1_prox(x, radius)


def box_prox(x, lo, hi):
    return np.clip(x, lo, hi)


def sql2_grad(x):  # norm.py:96-98
    return 2 * x


def sql2_prox(x, tau):  # norm.py:100-104
    return x / (2 * tau + 1)


def l21_apply(x, arg_shape, l2_axis=(0,)):  # norm.py:338-350
    sh = x.shape[:-1]
    a = x.reshape(sh + tuple(arg_shape))
    ax = tuple(len(sh) + int(i) for i in l2_axis)
    n = np.sqrt((a**2).sum(axis=ax, keepdims=True))
    return n.reshape(*sh, -1).sum(axis=-1, keepdims=True)


def l21_prox(x, tau, arg_shape, l2_axis=(0,)):  # norm.py:352-364
    sh = x.shape[:-1]
    a = x.reshape(sh + tuple(arg_shape))
    ax = tuple(len(sh) + int(i) for i in l2_axis)
    n = np.sqrt((a**2).sum(axis=ax, keepdims=True))
    out = a * (1 - tau / np.fmax(n, tau))
    return out.reshape(*sh, -1)


def fenchel_prox(prox, x, sigma):  # abc/operator.py:906-944 (Moreau identity)
    return x - sigma * prox(x / sigma, 1 / sigma)


# ----------------------------------------------------------------------------------------------
# Stopping criteria  (reference: src/pyxu/opt/stop.py:29-68, :222-297, :300-396)
# ----------------------------------------------------------------------------------------------
class MaxIter:
    def __init__(self, n):
        self.n, self.i = int(n), 0

    def stop(self, state):
        self.i += 1
        return self.i > self.n


class RelError:
    def __init__(self, eps, var="x", satisfy_all=True):
        self.eps, self.var, self.all = eps, var, satisfy_all
        self.prev, self.val = None, None

    def stop(self, state):
        x = state[self.var]
        if self.prev is None:
            self.prev = x.copy()
            return False
        num = np.linalg.norm(x - self.prev, axis=-1, keepdims=True)
        den = np.linalg.norm(self.prev, axis=-1, keepdims=True)
        rule = np.all if self.all else np.any
        decision = bool(rule(num <= self.eps * den))
        with np.errstate(all="ignore"):
            self.val = np.nan_to_num(num / den, nan=0.0)
        self.prev = x.copy()
        return decision


class AbsError:
    def __init__(self, eps, var="x", satisfy_all=True):
        self.eps, self.var, self.all = eps, var, satisfy_all
        self.val = None

    def stop(self, state):
        self.val = np.linalg.norm(state[self.var], axis=-1, keepdims=True)
        rule = np.all if self.all else np.any
        return bool(rule(self.val <= self.eps))


def run(state, step, stop_crit):
    """Solver._step() control flow: test the criterion, then step (abc/solver.py:588-652)."""
    n = 0
    while not stop_crit.stop(state):
        step(state)
        n += 1
    return n


# ----------------------------------------------------------------------------------------------
# Primal-dual splitting  (reference: src/pyxu/opt/solver/pds.py)
# Problem:  min_x f(x) + g(x) + h(Kx);  callables: grad_f(x), prox_g(x, tau), prox_h(z, tau),
# K / KT (apply / adjoint).  `None` stands for the NullFunc / NullOp of the reference.
# ----------------------------------------------------------------------------------------------
def pd3o_step_sizes(beta, K_lipschitz, has_h, tau=None, sigma=None, tuning_strategy=1):
    """PD3O._set_step_sizes + _optimize_step_sizes (pds.py:763-864) + momentum (pds.py:183-204)."""
    gamma = beta if tuning_strategy != 2 else beta / 1.9
    tau = None if tau == 0 else tau
    sigma = None if sigma == 0 else sigma
    L = K_lipschitz
    if tau is not None and sigma is None:
        sigma = 0 if not has_h else 1 / (tau * L**2)
    elif tau is None and sigma is not None:
        tau = 1 / gamma if not has_h else min(1 / (sigma * L**2), 1 / gamma)
    elif tau is None and sigma is None:
        if beta > 0:
            if not has_h:
                tau, sigma = 1 / gamma, 0
            else:
                # linprog (pds.py:849-864): max log t + log s, s.t. log t + log s <= log .99 - 2 log L,
                # log t <= -log gamma, t = s.  Closed form of that LP:
                t = min(0.5 * (math.log(0.99) - 2 * math.log(L)), math.log(1 / gamma))
                tau = sigma = math.exp(t)
        else:
            if not has_h:
                tau, sigma = 1, 0
            else:
                tau = sigma = 1 / L
    delta = 2 if beta == 0 else 2 - beta * tau / 2
    rho = 1.0 if tuning_strategy != 3 else delta - 0.1
    return float(tau), float(sigma), float(rho)


def cv_step_sizes(beta, K_lipschitz, has_h, f_quadratic, tau=None, sigma=None, tuning_strategy=1):
    """CondatVu._set_step_sizes (pds.py:444-517)."""
    gamma = beta if tuning_strategy != 2 else beta / 1.9
    tau = None if tau == 0 else tau
    sigma = None if sigma == 0 else sigma
    L = K_lipschitz
    if tau is not None and sigma is None:
        sigma = 0 if not has_h else ((1 / tau) - gamma) * (1 / L**2)
    elif tau is None and sigma is not None:
        tau = 1 / gamma if not has_h else 1 / (gamma + sigma * L**2)
    elif tau is None and sigma is None:
        if beta > 0:
            if not has_h:
                tau, sigma = 1 / gamma, 0
            else:
                tau = sigma = (1 / L**2) * ((-gamma / 2) + math.sqrt((gamma**2 / 4) + L**2))
        else:
            if not has_h:
                tau, sigma = 1, 0
            else:
                tau = sigma = 1 / L
    delta = 2 if (beta == 0 or (f_quadratic and gamma <= beta)) else 2 - beta / (2 * gamma)
    rho = 1.0 if tuning_strategy != 3 else delta - 0.1
    return float(tau), float(sigma), float(rho)


def _zero(x, *a):
    return np.zeros_like(x)


def _ident(x, *a):
    return x


def pd3o_init(x0, K, z0=None, g_null=False, h_null=False):
    """PD3O.m_init (pds.py:722-745)."""
    st = dict(x=x0, z=(K(x0.copy()) if z0 is None else z0))
    st["u"] = x0 * 1.01 if (g_null and h_null) else x0.copy()
    return st


def pd3o_step(st, tau, sigma, rho, grad_f, prox_g, prox_h, K, KT):
    """PD3O.m_step (pds.py:747-761).  prox_h=None <=> h is NullFunc."""
    x = prox_g(st["u"] - tau * KT(st["z"]), tau)
    u_temp = x - tau * grad_f(x)
    if prox_h is not None:
        z_temp = fenchel_prox(prox_h, st["z"] + sigma * K(x + u_temp - st["u"]), sigma)
        st["z"] = (1 - rho) * st["z"] + rho * z_temp
    st["u"] = (1 - rho) * st["u"] + rho * u_temp
    st["x"] = x


def cv_step(st, tau, sigma, rho, grad_f, prox_g, prox_h, K, KT):
    """CondatVu.m_step (pds.py:429-442)."""
    x_temp = prox_g(st["x"] - tau * grad_f(st["x"]) - tau * KT(st["z"]), tau)
    if prox_h is not None:
        u = 2 * x_temp - st["x"]
        z_temp = fenchel_prox(prox_h, st["z"] + sigma * K(u), sigma)
        st["z"] = rho * z_temp + (1 - rho) * st["z"]
    st["x"] = rho * x_temp + (1 - rho) * st["x"]


# ----------------------------------------------------------------------------------------------
# PGD / FISTA  (reference: src/pyxu/opt/solver/pgd.py:129-191)
# ----------------------------------------------------------------------------------------------
def pgd_init(x0):
    return dict(x=x0, x_prev=x0, k=0)


def pgd_step(st, tau, grad_f, prox_g, acceleration=True, d=75):
    k = st["k"]
    a = (k / (k + 1 + d)) if acceleration else 0.0
    st["k"] = k + 1
    y = st["x"] + a * (st["x"] - st["x_prev"])
    z = y - tau * grad_f(y)
    st["x_prev"], st["x"] = st["x"], prox_g(z, tau)


# ----------------------------------------------------------------------------------------------
# Convenience problem builders used by tests / bench (compositions the reference expresses through
# operator arithmetic: ScaleRule, ArgShiftRule, ChainRule -- src/pyxu/abc/arithmetic.py:65-260,
# :479-665, :1034-1345).
# ----------------------------------------------------------------------------------------------
def tv_problem(y, arg_shape, lam, mode="constant", blur=None, positivity=True, dtype=np.float64, data_scale=0.5):
    """f = data_scale*||A x - y||^2 (A = blur or Id), g = PositiveOrthant or 0, h = lam*L21, K = Gradient."""
    y = np.asarray(y, dtype=dtype)
    D = len(arg_shape)
    K = Gradient(arg_shape, mode=mode, dtype=dtype)
    if blur is None:
        grad_f = lambda x: (2 * data_scale) * (x - y)
        beta = 2 * data_scale
    else:
        grad_f = lambda x: (2 * data_scale) * blur.adjoint(blur.apply(x) - y)
        beta = 2 * data_scale * blur.lipschitz**2
    prox_g = positive_orthant_prox if positivity else _ident
    prox_h = lambda z, t: l21_prox(z, lam * t, (D,) + tuple(arg_shape), (0,))
    return dict(grad_f=grad_f, prox_g=prox_g, prox_h=prox_h, K=K.apply, KT=K.adjoint, beta=beta, K_lipschitz=K.lipschitz)
