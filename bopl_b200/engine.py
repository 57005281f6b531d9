"""Device engine: one `bopl_handle` (one GPU) behind a small Python object.

PyTorch is used only to own device buffers and to name the stream; every computation goes through the C-ABI
(`bopl_b200/_lib.py` -> libbopl_b200.so).  NumPy arrays in / NumPy arrays out for the reference-facing calls,
torch CUDA tensors in / out for callers that keep data resident in HBM.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import KERNEL_KINDS, UTILITY_KINDS, VAR_CLIP, VAR_INCLUDE_NOISE, BoplError, check


class GPState(object):
    """Flat fp64 arrays of m attributes x H hyper-samples of exact GPs — what `bopl_set_model` consumes.

    kinds (m,H) of GPy kernel names; X (n,d); lengthscale (m,H,d); variance, noise, ymean (m,H);
    L (m,H,n,n) lower Cholesky factors of K + (noise+1e-8) I, C order; alpha (m,H,n); Winv (m,H,n,n) or None;
    Linv (m,H,n,n) = L^-1 (dtrtri) or None — enables the small-batch matrix-vector path.
    Field map from the reference's live objects: SURVEY.md §7.1-10."""

    def __init__(self, kinds, X, lengthscale, variance, noise, ymean, L, alpha, Winv=None, Linv=None, Y=None):
        self.X = np.ascontiguousarray(X, dtype=np.float64)
        self.n, self.d = self.X.shape
        # raw targets (m,n) when known: needed only to start a Thompson sample path (Engine.path_begin)
        self.Y = None if Y is None else np.ascontiguousarray(np.asarray(Y, dtype=np.float64).reshape(-1, self.n))
        self.variance = np.ascontiguousarray(variance, dtype=np.float64)
        self.m, self.H = self.variance.shape
        self.kinds = np.asarray(kinds).reshape(self.m, self.H)
        self.lengthscale = np.ascontiguousarray(lengthscale, dtype=np.float64).reshape(self.m, self.H, self.d)
        self.noise = np.ascontiguousarray(noise, dtype=np.float64).reshape(self.m, self.H)
        self.ymean = np.ascontiguousarray(ymean, dtype=np.float64).reshape(self.m, self.H)
        # L / alpha are None for a state fitted on the device (Engine.fit_model): the factor never existed on the host
        self.L = None if L is None else np.ascontiguousarray(L, dtype=np.float64).reshape(self.m, self.H, self.n, self.n)
        self.alpha = None if alpha is None else np.ascontiguousarray(alpha, dtype=np.float64).reshape(self.m, self.H, self.n)
        self.on_device = False
        self.Winv = None if Winv is None else \
            np.ascontiguousarray(Winv, dtype=np.float64).reshape(self.m, self.H, self.n, self.n)
        self.Linv = None if Linv is None else \
            np.ascontiguousarray(Linv, dtype=np.float64).reshape(self.m, self.H, self.n, self.n)


def _torch():
    import torch
    return torch


class Engine(object):
    """One GPU.  Methods taking `*_dev` arguments expect torch CUDA fp64 tensors on this engine's device."""

    def __init__(self, device=0, posterior_kernel=None):
        self.lib = _lib.load()
        torch = _torch()
        if not torch.cuda.is_available():
            raise BoplError(-2, "no CUDA device: bopl_b200 has no CPU fallback")
        self.device_index = int(device)
        self.device = torch.device("cuda", self.device_index)
        h = ctypes.c_void_p()
        check(self.lib.bopl_create(ctypes.byref(h), self.device_index))
        self.h = h
        self.state = None
        self.sms = self.lib.bopl_device_sms(self.h)
        if posterior_kernel is not None:
            self.set_posterior_kernel(posterior_kernel)

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.bopl_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return ctypes.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    def _check(self, rc):
        check(rc, self.h)

    def to_dev(self, a):
        torch = _torch()
        if isinstance(a, torch.Tensor):
            return a.to(device=self.device, dtype=torch.float64).contiguous()
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.device)

    def empty(self, *shape):
        torch = _torch()
        return torch.empty(*shape, dtype=torch.float64, device=self.device)

    @property
    def launch_count(self):
        return int(self.lib.bopl_launch_count(self.h))

    # ------------------------------------------------------------------ model
    def set_posterior_kernel(self, kind):
        """'int8x7' (update on the int8 tensor pipe, 7-digit = 55-bit exact split; default), 'fp64' (DMMA blocked solve),
        'int8' (6-digit split: opt-in reduced precision) or 'int8auto' (6 digits when an upload-time probe of the model at its training
        points meets 1e-9 relative on the variance, else 7; `posterior_kernel` then reports what the loaded model runs).
        Select before the model is uploaded / fitted: the digit images of the factor are built then (`bopl_set_posterior_kernel`)."""
        self._check(self.lib.bopl_set_posterior_kernel(self.h, _lib.POSTERIOR_KERNELS[kind]))

    @property
    def posterior_kernel(self):
        k = int(self.lib.bopl_get_posterior_kernel(self.h))
        return {v: n for n, v in _lib.POSTERIOR_KERNELS.items()}[k]

    def set_model(self, state):
        if state.L is None:
            raise BoplError(-3, "this GPState was fitted on the device (Engine.fit_model) and holds no host factor to upload")
        kinds = np.ascontiguousarray([[KERNEL_KINDS[str(k)] for k in row] for row in state.kinds], dtype=np.int32)
        p = lambda a: ctypes.c_void_p(a.ctypes.data)
        self._check(self.lib.bopl_set_model(
            self.h, state.n, state.d, state.m, state.H, p(kinds), p(state.X), p(state.lengthscale), p(state.variance),
            p(state.noise), p(state.ymean), p(state.L), p(state.alpha),
            p(state.Winv) if state.Winv is not None else None,
            p(state.Linv) if state.Linv is not None else None))
        self.state = state
        self.n, self.d, self.m, self.H = state.n, state.d, state.m, state.H

    def fit_model(self, X, Y, kinds, variance, lengthscale, noise, want_winv=True, want_linv=True, max_tries=10):
        """Fit every (attribute, hyper-sample) posterior ON THE DEVICE (`bopl_fit_model`): K, blocked Cholesky, alpha,
        K^-1, L^-1 — nothing but X, Y and the hyper-parameters crosses the bus.  A matrix that is not positive definite
        is retried with growing diagonal jitter exactly as `GPy/util/linalg.py:53-78` (jitchol) does.
        Returns a `GPState` without the factor arrays (read them back with `get_factor` if needed)."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        n, d = X.shape
        variance = np.ascontiguousarray(variance, dtype=np.float64)
        m, H = variance.shape
        lengthscale = np.asarray(lengthscale, dtype=np.float64).reshape(m, H, -1)
        if lengthscale.shape[2] == 1 and d > 1:
            lengthscale = np.repeat(lengthscale, d, axis=2)
        lengthscale = np.ascontiguousarray(lengthscale)
        noise = np.ascontiguousarray(np.asarray(noise, dtype=np.float64).reshape(m, H))
        kinds = np.asarray(kinds)
        kinds = np.repeat(kinds.reshape(m, 1), H, axis=1) if kinds.size == m else kinds.reshape(m, H)
        kind_ids = np.ascontiguousarray([[KERNEL_KINDS[str(k)] for k in row] for row in kinds], dtype=np.int32)
        Ym = np.ascontiguousarray(np.stack([np.asarray(Y[j], dtype=np.float64).reshape(n) for j in range(m)]))
        p = lambda a: ctypes.c_void_p(a.ctypes.data)
        logdet = np.empty((m, H))
        jitter = np.zeros((m, H))
        tries = np.zeros((m, H), dtype=np.int64)
        info = np.zeros((m, H), dtype=np.int32)
        while True:
            rc = self.lib.bopl_fit_model(self.h, n, d, m, H, p(kind_ids), p(X), p(Ym), p(lengthscale), p(variance), p(noise),
                                         p(jitter), int(bool(want_winv)), int(bool(want_linv)), p(logdet))
            if rc != _lib.ERR_NUMERIC:
                break
            # jitchol (GPy/util/linalg.py:53-78) PER MATRIX: only the instances whose factorisation failed get jitter —
            # mean(diag) * 1e-6, then x10 per retry, at most max_tries times; the healthy instances keep their factors unchanged
            self._check(self.lib.bopl_fit_info(self.h, m * H, p(info)))
            failed = info != 0
            if not failed.any() or np.any(tries[failed] >= max_tries):
                break
            jitter[failed] = ((variance + noise + 1e-8) * 1e-6)[failed] * (10.0 ** tries[failed])
            tries[failed] += 1
        if rc != 0:
            self.state = None                   # the device holds no model now (bopl_fit_model frees it first): fail clearly later
        self._check(rc)
        state = GPState(kinds, X, lengthscale, variance, noise, np.repeat(Ym.mean(axis=1)[:, None], H, axis=1), None, None, Y=Ym)
        state.on_device = True
        state.has_winv, state.has_linv = bool(want_winv), bool(want_linv)
        state.logdet, state.jitter = logdet, jitter
        self.state = state
        self.n, self.d, self.m, self.H = n, d, m, H
        return state

    def get_factor(self, j, hsel, want_winv=False, want_linv=False):
        """(L, alpha, Winv, Linv) of attribute j under hyper-sample hsel, read back from the device (`bopl_get_factor`)."""
        n = self.n
        L, alpha = np.empty((n, n)), np.empty(n)
        Winv = np.empty((n, n)) if want_winv else None
        Linv = np.empty((n, n)) if want_linv else None
        p = lambda a: ctypes.c_void_p(a.ctypes.data) if a is not None else None
        self._check(self.lib.bopl_get_factor(self.h, int(j), int(hsel), p(L), p(alpha), p(Winv), p(Linv)))
        return L, alpha, Winv, Linv

    def log_marginal(self, X, Yc, kinds, variance, lengthscale, noise, want_grad=True):
        """Log marginal likelihood (+ gradient wrt variance, lengthscales, noise) of G independent GPs on the same inputs:
        one batched likelihood evaluation of MAP / HMC (`bopl_log_marginal`).  Yc (G,n) centred targets; variance, noise (G,);
        lengthscale (G,d).  Returns ll (G,), grad (G,d+2) or None, info (G,) (non-zero: not positive definite)."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        n, d = X.shape
        Yc = np.ascontiguousarray(Yc, dtype=np.float64).reshape(-1, n)
        G = Yc.shape[0]
        variance = np.ascontiguousarray(variance, dtype=np.float64).reshape(G)
        noise = np.ascontiguousarray(noise, dtype=np.float64).reshape(G)
        lengthscale = np.ascontiguousarray(lengthscale, dtype=np.float64).reshape(G, d)
        kind_ids = np.ascontiguousarray([KERNEL_KINDS[str(k)] for k in np.asarray(kinds).reshape(G)], dtype=np.int32)
        ll = np.empty(G)
        grad = np.empty((G, d + 2)) if want_grad else None
        info = np.zeros(G, dtype=np.int32)
        p = lambda a: ctypes.c_void_p(a.ctypes.data) if a is not None else None
        self._check(self.lib.bopl_log_marginal(self.h, n, d, G, p(kind_ids), p(X), p(Yc), p(lengthscale), p(variance), p(noise),
                                               p(ll), p(grad), p(info)))
        return ll, grad, info

    # ------------------------------------------------------------------ posterior (device tensors)
    def cross_cov_dev(self, Xc, hsel, j):
        N = Xc.shape[0]
        K = self.empty(N, self.n)
        self._check(self.lib.bopl_cross_cov(self.h, Xc.data_ptr(), N, hsel, j, K.data_ptr(), self._stream()))
        return K

    def posterior_dev(self, Xc, hsel, include_noise=True, clip=True, want_mean=True, want_var=True):
        N = Xc.shape[0]
        mean = self.empty(self.m, N) if want_mean else None
        var = self.empty(self.m, N) if want_var else None
        flags = (VAR_INCLUDE_NOISE if include_noise else 0) | (VAR_CLIP if clip else 0)
        self._check(self.lib.bopl_posterior(self.h, Xc.data_ptr(), N, hsel, mean.data_ptr() if want_mean else None,
                                            var.data_ptr() if want_var else None, flags, self._stream()))
        return mean, var

    def posterior_mean_dev(self, Xc, hsel):
        N = Xc.shape[0]
        mean = self.empty(self.m, N)
        self._check(self.lib.bopl_posterior_mean(self.h, Xc.data_ptr(), N, hsel, mean.data_ptr(), self._stream()))
        return mean

    def train_mean_dev(self):
        F = self.empty(self.H, self.m, self.n)
        self._check(self.lib.bopl_train_mean(self.h, F.data_ptr(), self._stream()))
        return F

    def posterior_grad_dev(self, Xc, hsel, want_mean=True, want_var=True):
        N = Xc.shape[0]
        dmean = self.empty(self.m, N, self.d) if want_mean else None
        dvar = self.empty(self.m, N, self.d) if want_var else None
        self._check(self.lib.bopl_posterior_grad(self.h, Xc.data_ptr(), N, hsel,
                                                 dmean.data_ptr() if want_mean else None,
                                                 dvar.data_ptr() if want_var else None, self._stream()))
        return dmean, dvar

    # ------------------------------------------------------------------ acquisition (device tensors)
    def incumbent_dev(self, kind, theta):
        Lt, p = theta.shape
        best = self.empty(self.H, Lt)
        self._check(self.lib.bopl_incumbent(self.h, UTILITY_KINDS[kind], theta.data_ptr(), Lt, p, best.data_ptr(),
                                            self._stream()))
        return best

    def uei_mc_dev(self, mean, var, Z, kind, theta, wts, best, scale, acq=None, accumulate=False):
        N = mean.shape[1]
        Lt, p = theta.shape
        if acq is None:
            acq = self.empty(N)
        self._check(self.lib.bopl_uei_mc(self.h, mean.data_ptr(), var.data_ptr(), N, Z.data_ptr(), Z.shape[0],
                                         UTILITY_KINDS[kind], theta.data_ptr(), Lt, p, wts.data_ptr(), best.data_ptr(),
                                         float(scale), int(bool(accumulate)), acq.data_ptr(), self._stream()))
        return acq

    def uei_dev(self, Xc, Z, kind, theta, wts, best, nH, acq=None):
        N = Xc.shape[0]
        Lt, p = theta.shape
        if acq is None:
            acq = self.empty(N)
        self._check(self.lib.bopl_uei(self.h, Xc.data_ptr(), N, Z.data_ptr(), Z.shape[0], UTILITY_KINDS[kind],
                                      theta.data_ptr(), Lt, p, wts.data_ptr(), best.data_ptr(), nH, acq.data_ptr(),
                                      self._stream()))
        return acq

    def uei_grad_dev(self, Xc, Z, kind, theta, wts, best, nH):
        N = Xc.shape[0]
        Lt, p = theta.shape
        acq = self.empty(N)
        dacq = self.empty(N, self.d)
        self._check(self.lib.bopl_uei_grad(self.h, Xc.data_ptr(), N, Z.data_ptr(), Z.shape[0], UTILITY_KINDS[kind],
                                           theta.data_ptr(), Lt, p, wts.data_ptr(), best.data_ptr(), nH,
                                           acq.data_ptr(), dacq.data_ptr(), self._stream()))
        return acq, dacq

    def uei_affine_dev(self, Xc, theta, wts, best, nH, with_grad=False):
        N = Xc.shape[0]
        Lt = theta.shape[0]
        acq = self.empty(N)
        dacq = self.empty(N, self.d) if with_grad else None
        self._check(self.lib.bopl_uei_affine(self.h, Xc.data_ptr(), N, theta.data_ptr(), Lt, wts.data_ptr(),
                                             best.data_ptr(), nH, acq.data_ptr(),
                                             dacq.data_ptr() if with_grad else None, self._stream()))
        return acq, dacq

    def expected_utility_dev(self, Xc, Z, kind, theta, wts, nH, with_grad=False):
        N = Xc.shape[0]
        Lt, p = theta.shape
        val = self.empty(N)
        dval = self.empty(N, self.d) if with_grad else None
        self._check(self.lib.bopl_expected_utility(self.h, Xc.data_ptr(), N, Z.data_ptr(), Z.shape[0], UTILITY_KINDS[kind],
                                                   theta.data_ptr(), Lt, p, wts.data_ptr(), nH, val.data_ptr(),
                                                   dval.data_ptr() if with_grad else None, self._stream()))
        return val, dval

    def uei_constrained_dev(self, Xc, theta, wts, best, nH, with_grad=False):
        N = Xc.shape[0]
        Lt = theta.shape[0]
        acq = self.empty(N)
        dacq = self.empty(N, self.d) if with_grad else None
        self._check(self.lib.bopl_uei_constrained(self.h, Xc.data_ptr(), N, theta.data_ptr(), Lt, wts.data_ptr(),
                                                  best.data_ptr(), nH, acq.data_ptr(),
                                                  dacq.data_ptr() if with_grad else None, self._stream()))
        return acq, dacq

    def topk_dev(self, acq, k, index_offset=0):
        torch = _torch()
        N = acq.shape[0]
        val = self.empty(k)
        idx = torch.empty(k, dtype=torch.int64, device=self.device)
        self._check(self.lib.bopl_topk(self.h, acq.data_ptr(), N, k, int(index_offset), val.data_ptr(), idx.data_ptr(),
                                       self._stream()))
        return val, idx

    def uei_grad_host(self, Xc, Z, kind, theta, wts, best, nH):
        """NumPy in / NumPy out value + gradient through `bopl_uei_grad_host` (one library call, one synchronisation)."""
        c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        Xc, Z, theta, wts, best = c(Xc), c(Z), c(theta), c(wts), c(best)
        N = Xc.shape[0]
        Lt, p = theta.shape
        acq = np.empty(N)
        dacq = np.empty((N, self.d))
        ptr = lambda a: ctypes.c_void_p(a.ctypes.data)
        self._check(self.lib.bopl_uei_grad_host(self.h, ptr(Xc), N, ptr(Z), Z.shape[0], UTILITY_KINDS[kind], ptr(theta), Lt, p,
                                                ptr(wts), ptr(best), nH, ptr(acq), ptr(dacq)))
        return acq, dacq

    # ------------------------------------------------------------------ Thompson sample path (uTS.py:40-50)
    def path_begin(self, hsel=0, Y=None, capacity=640):
        """Start a posterior sample path on hyper-sample `hsel` (`bopl_path_begin`).  Y (m,n): the raw targets of the
        loaded model (default: the ones the state carries)."""
        Y = self.state.Y if Y is None else np.ascontiguousarray(np.asarray(Y, dtype=np.float64).reshape(self.m, self.n))
        if Y is None:
            raise BoplError(-3, "the loaded GPState carries no targets; pass Y")
        self._check(self.lib.bopl_path_begin(self.h, int(hsel), int(capacity), ctypes.c_void_p(Y.ctypes.data)))

    def path_eval(self, x, z):
        """One evaluation of the sample path at x (d,) with the caller's standard-normal draws z (m,): returns
        (y, mean, var), each (m,); (x, y) joins the path (`bopl_path_eval_host`)."""
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))
        z = np.ascontiguousarray(np.asarray(z, dtype=np.float64).reshape(-1))
        if x.size != self.d or z.size != self.m:
            raise ValueError("x must have %d entries and z %d" % (self.d, self.m))
        out = np.empty((3, self.m))
        ptr = lambda a: ctypes.c_void_p(a.ctypes.data)
        self._check(self.lib.bopl_path_eval_host(self.h, ptr(x), ptr(z), ptr(out[0]), ptr(out[1]), ptr(out[2])))
        return out[0], out[1], out[2]

    @property
    def path_length(self):
        return int(self.lib.bopl_path_length(self.h))

    def path_end(self):
        self._check(self.lib.bopl_path_end(self.h))

    # ------------------------------------------------------------------ host-buffer path (the reference-facing call)
    def uei_host(self, Xc, Z, kind, theta, wts, best, nH, k=0, index_offset=0, want_acq=True, out=None):
        """NumPy in / NumPy out through `bopl_uei_host` (chunked H2D overlapped with compute inside the library).
        Xc may be a pinned torch CPU tensor or a NumPy array; it must be C-contiguous fp64."""
        torch = _torch()

        def hp(a):
            if isinstance(a, torch.Tensor):
                assert a.device.type == "cpu" and a.dtype == torch.float64 and a.is_contiguous()
                return ctypes.c_void_p(a.data_ptr()), a
            a = np.ascontiguousarray(a, dtype=np.float64)
            return ctypes.c_void_p(a.ctypes.data), a

        px, Xc = hp(Xc)
        pz, Z = hp(Z)
        pt, theta = hp(theta)
        pw, wts = hp(wts)
        pb, best = hp(best)
        N = Xc.shape[0]
        Lt, p = theta.shape
        acq = None
        if want_acq:
            acq = out if out is not None else np.empty(N, dtype=np.float64)
            pa = ctypes.c_void_p(acq.data_ptr() if isinstance(acq, torch.Tensor) else acq.ctypes.data)
        else:
            pa = None
        tv = np.empty(max(k, 1), dtype=np.float64)
        ti = np.empty(max(k, 1), dtype=np.int64)
        self._check(self.lib.bopl_uei_host(self.h, px, N, pz, Z.shape[0], UTILITY_KINDS[kind], pt, Lt, p, pw, pb, nH, pa,
                                           k, int(index_offset), ctypes.c_void_p(tv.ctypes.data),
                                           ctypes.c_void_p(ti.ctypes.data)))
        if k:
            return acq, tv[:k], ti[:k]
        return acq
