"""Whole-sequence interface to the B200 library (the moihgp_cuda_* entry points).

``MOIHGPSequences`` keeps the reference's model vocabulary (``update(params)``, ``params``,
``negLogLikelihood``) but takes whole sequences ``Y[N][T][p]`` where the reference's callers loop
over observations (RegressionObjective::operator(), moihgp_regression.h:34-52; OnlineObjective,
moihgp_online.h:40-72; MOIHGPRegression::predict, moihgp_regression.h:127-139).

NumPy arrays go through the host entry points (copies inside the call); torch CUDA tensors go
through the ``*_dev`` entry points on torch's current stream and stay resident in HBM.
"""
import ctypes

import numpy as np

from . import _lib

SMOOTH_NONE, SMOOTH_REFERENCE_LITERAL, SMOOTH_RTS = -1, 0, 1
_KERNELS = {"Matern32": 32, "Matern52": 52}


def _np(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    return a.data_ptr()  # torch tensor


def _torch_stream(device):
    """cudaStream_t of torch's current stream on `device`.  torch reports its default stream as handle 0, which the C ABI
    reads as "the handle's own stream": pass CUDA's explicit alias of the legacy default stream (cudaStreamLegacy = 0x1)
    instead, so that the library's kernels are ordered with torch's own work on that stream."""
    import torch
    return torch.cuda.current_stream(device).cuda_stream or 1


class MOIHGPSequences(object):

    def __init__(self, dt, num_output, num_latent, kernel="Matern32", threading=False, device=-1):
        self._lib = _lib.load()
        h = ctypes.c_void_p()
        rc = self._lib.moihgp_cuda_create(ctypes.byref(h), _KERNELS[kernel], dt, num_output, num_latent, int(threading), device)
        if rc != 0 or not h:
            raise RuntimeError("moihgp_cuda_create failed (no B200-class CUDA device? there is no CPU fallback)")
        self._h = h
        self.dt, self.num_output, self.num_latent, self.kernel = dt, num_output, num_latent, kernel
        self.igp_dim = int(self._lib.moihgp_cuda_igp_dim(h))
        self.num_param = int(self._lib.moihgp_cuda_num_param(h))
        self.num_igp_param = int(self._lib.moihgp_cuda_num_igp_param(h))

    @classmethod
    def adopt(cls, model):
        """The whole-sequence interface on the handle of an existing per-observation model (``pywrapper.MOIHGP``): ONE model on
        the device, shared - ``update`` through either object is seen by both.  ``model`` keeps ownership of the handle (and
        is kept alive by the returned object)."""
        import ctypes as _ct
        self = cls.__new__(cls)
        self._lib = _lib.load()
        self._h = _ct.c_void_p(model.handle)
        self._owner = model
        self.dt, self.num_output, self.num_latent, self.kernel = model.dt, model.num_output, model.num_latent, None
        self.igp_dim = int(self._lib.moihgp_cuda_igp_dim(self._h))
        self.num_param = int(self._lib.moihgp_cuda_num_param(self._h))
        self.num_igp_param = int(self._lib.moihgp_cuda_num_igp_param(self._h))
        return self

    def __del__(self):
        if getattr(self, "_h", None) and getattr(self, "_owner", None) is None:
            self._lib.moihgp_cuda_destroy(self._h)
        self._h = None

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("libmoihgp: " + self._lib.moihgp_cuda_last_error(self._h).decode())

    # ---- model ------------------------------------------------------------------------------
    def update(self, params):
        params = _np(params)
        assert params.size == self.num_param
        self._check(self._lib.moihgp_cuda_update(self._h, params.ctypes.data_as(_lib.c_double_p)))

    @property
    def params(self):
        out = np.zeros(self.num_param)
        self._check(self._lib.moihgp_cuda_get_params(self._h, out.ctypes.data_as(_lib.c_double_p)))
        return out

    @property
    def U(self):
        out = np.zeros((self.num_output, self.num_latent))
        self._check(self._lib.moihgp_cuda_get_U(self._h, out.ctypes.data_as(_lib.c_double_p)))
        return out

    def latent_consts(self, l):
        flat = np.zeros(256)
        n = self._lib.moihgp_cuda_latent_consts(self._h, l, flat.ctypes.data_as(_lib.c_double_p), flat.size)
        if n < 0:
            raise RuntimeError("moihgp_cuda_latent_consts failed")
        d, out, o = self.igp_dim, {}, 0
        lay = [("A", (d, d)), ("Q", (d, d)), ("K", (d,)), ("S", ()), ("PF", (d, d)), ("HA", (d,)), ("AKHA", (d, d))]
        for k in range(3):
            lay += [("dS%d" % k, ()), ("dA%d" % k, (d, d)), ("dK%d" % k, (d,)), ("dAKHA%d" % k, (d, d)), ("HdA%d" % k, (d,))]
        for name, shp in lay:
            m = int(np.prod(shp)) if shp else 1
            out[name] = flat[o:o + m].reshape(shp).copy() if shp else float(flat[o])
            o += m
        return out

    def latent_iters(self, l):
        out = (ctypes.c_int * 8)()
        self._check(self._lib.moihgp_cuda_latent_iters(self._h, l, out))
        return list(out)

    def smoother_consts(self, l, mode):
        d = self.igp_dim
        G, P = np.zeros((d, d)), np.zeros((d, d))
        self._check(self._lib.moihgp_cuda_smoother_consts(self._h, l, mode, G.ctypes.data_as(_lib.c_double_p), P.ctypes.data_as(_lib.c_double_p)))
        return G, P

    @property
    def nan_status(self):
        """0 / 1 / 2 (none / handled / overflow) for the last whole-sequence call; waits for the stream.  The *_device methods
        are asynchronous and cannot raise on an overflow of the missing-observation list themselves."""
        out = ctypes.c_int(0)
        self._check(self._lib.moihgp_cuda_nan_status(self._h, ctypes.byref(out)))
        return int(out.value)

    @property
    def launch_count(self):
        return int(self._lib.moihgp_cuda_launch_count(self._h))

    def set_path(self, path):
        """0 auto, 1 chunked scan, 2 many-chains (see moihgp_cuda_set_path)."""
        self._check(self._lib.moihgp_cuda_set_path(self._h, {"auto": 0, "scan": 1, "chain": 2}.get(path, path)))

    def set_chain_seqs_per_warp(self, n):
        """Tuning knob of the many-chains kernels (0 = automatic); see moihgp_cuda_set_chain_seqs_per_warp."""
        self._check(self._lib.moihgp_cuda_set_chain_seqs_per_warp(self._h, int(n)))

    def profile(self, enable):
        self._check(self._lib.moihgp_cuda_profile(self._h, int(enable)))

    def profile_read(self):
        """{kernel name: (total ms, launches)} since profile(True); clears the record."""
        out = {}
        for line in self._lib.moihgp_cuda_profile_read(self._h).decode().splitlines():
            name, ms, cnt = line.split()
            out[name] = (float(ms), int(cnt))
        return out

    def set_stream(self, cuda_stream_ptr):
        self._check(self._lib.moihgp_cuda_set_stream(self._h, cuda_stream_ptr))

    def synchronize(self):
        self._check(self._lib.moihgp_cuda_sync(self._h))

    # ---- fused pass: filter + smoother + NLL ----------------------------------------------------
    def filter_smoother_nll(self, Y, x0=None, smoother_mode=SMOOTH_REFERENCE_LITERAL, want_states=True, want_yhat=False, want_nll=True):
        """Y: [N,T,p] (or [T,p]) NumPy array -> dict of NumPy arrays X, Xs, Yhat, nll, xT.
        smoother_mode defaults to the reference's own recursion (IHGP::backwardSmoother as written, ihgp.h:108-113, SURVEY Q3 -
        unstable for some hyper-parameters, e.g. the default Matern-3/2 latent); SMOOTH_RTS is this library's extension."""
        Y = _np(Y)
        if Y.ndim == 2:
            Y = Y[None]
        N, T, p = Y.shape
        assert p == self.num_output
        L, d = self.num_latent, self.igp_dim
        x0 = None if x0 is None else _np(x0).reshape(N, L, d)
        X = np.empty((N, T, L, d)) if want_states else None
        Xs = np.empty((N, T, L, d)) if (want_states and smoother_mode >= 0) else None
        Yhat = np.empty((N, T, p)) if want_yhat else None
        nll = np.empty(N) if want_nll else None
        xT = np.empty((N, L, d))
        self._check(self._lib.moihgp_cuda_filter_smoother_nll(self._h, _ptr(Y), N, T, _ptr(x0), smoother_mode, _ptr(X), _ptr(Xs),
                                                              _ptr(Yhat), _ptr(nll), _ptr(xT)))
        return {"X": X, "Xs": Xs, "Yhat": Yhat, "nll": nll, "xT": xT}

    def filter_smoother_nll_values(self, Y, x0=None, smoother_mode=SMOOTH_REFERENCE_LITERAL, want_yhat=False):
        """The same pass returning only the function-value component H x = x(0) of the filtered / smoothed states,
        F and Fs [N,T,L] (d times fewer bytes across PCIe), plus nll, xT and optionally Yhat."""
        Y = _np(Y)
        if Y.ndim == 2:
            Y = Y[None]
        N, T, p = Y.shape
        L, d = self.num_latent, self.igp_dim
        x0 = None if x0 is None else _np(x0).reshape(N, L, d)
        F = np.empty((N, T, L))
        Fs = np.empty((N, T, L)) if smoother_mode >= 0 else None
        Yhat = np.empty((N, T, p)) if want_yhat else None
        nll, xT = np.empty(N), np.empty((N, L, d))
        self._check(self._lib.moihgp_cuda_filter_smoother_nll_values(self._h, _ptr(Y), N, T, _ptr(x0), smoother_mode, _ptr(F), _ptr(Fs),
                                                                     _ptr(Yhat), _ptr(nll), _ptr(xT)))
        return {"F": F, "Fs": Fs, "Yhat": Yhat, "nll": nll, "xT": xT}

    def filter_smoother_nll_device(self, Y, x0=None, smoother_mode=SMOOTH_REFERENCE_LITERAL, X=None, Xs=None, Yhat=None, nll=None, xT=None):
        """Device-resident variant: all arguments are torch CUDA float64 tensors (outputs pre-allocated by the
        caller, any may be None); runs asynchronously on torch's current stream."""
        import torch
        assert Y.is_cuda and Y.dtype == torch.float64 and Y.is_contiguous() and Y.dim() == 3
        N, T, _ = Y.shape
        self.set_stream(_torch_stream(Y.device))
        self._check(self._lib.moihgp_cuda_filter_smoother_nll_dev(self._h, _ptr(Y), N, T, _ptr(x0), smoother_mode, _ptr(X), _ptr(Xs),
                                                                  _ptr(Yhat), _ptr(nll), _ptr(xT)))

    # ---- objective: NLL + gradient ----------------------------------------------------------------
    def objective(self, Y, x0=None, dx0=None, want_state=False):
        """sum over sequences and time of negLogLikelihood(x, y, dx, grad) with step v2 advancing the state
        (RegressionObjective::operator(), moihgp_regression.h:42-50).  Returns (loss, grad[, xT, dxT])."""
        Y = _np(Y)
        if Y.ndim == 2:
            Y = Y[None]
        N, T, p = Y.shape
        assert p == self.num_output
        L, d = self.num_latent, self.igp_dim
        x0 = None if x0 is None else _np(x0).reshape(N, L, d)
        dx0 = None if dx0 is None else _np(dx0).reshape(N, L, 3, d)
        loss = np.zeros(1)
        grad = np.zeros(self.num_param)
        xT = np.empty((N, L, d)) if want_state else None
        dxT = np.empty((N, L, 3, d)) if want_state else None
        self._check(self._lib.moihgp_cuda_objective(self._h, _ptr(Y), N, T, _ptr(x0), _ptr(dx0), _ptr(loss), _ptr(grad), _ptr(xT), _ptr(dxT)))
        if want_state:
            return float(loss[0]), grad, xT, dxT
        return float(loss[0]), grad

    def objective_begin_device(self, Y, want_end=True):
        """Time-sharded evaluation, step 1 (torch CUDA tensor Y [N,T,p] = this rank's block of time): projection + summaries;
        returns the block's end state from a zero carry-in as (x [N,L,d], dx [N,L,3,d]) NumPy arrays (needs T % 256 == 0)."""
        import torch
        assert Y.is_cuda and Y.dtype == torch.float64 and Y.is_contiguous() and Y.dim() == 3
        N, T, _ = Y.shape
        L, d = self.num_latent, self.igp_dim
        self.set_stream(_torch_stream(Y.device))
        z = np.zeros((N, L, 4, d)) if want_end else None
        self._check(self._lib.moihgp_cuda_objective_begin_dev(self._h, _ptr(Y), N, T, _ptr(z)))
        return (z[:, :, 0].copy(), z[:, :, 1:].copy()) if want_end else None

    def objective_begin_async(self, Y, zend=None):
        """Like objective_begin_device, without a host round trip: the block's end state from a zero carry-in goes to the
        torch CUDA tensor ``zend`` [N,L,4,d] (None on the last block); asynchronous on torch's current stream."""
        N, T, _ = Y.shape
        self.set_stream(_torch_stream(Y.device))
        self._check(self._lib.moihgp_cuda_objective_begin_async(self._h, _ptr(Y), N, T, _ptr(zend)))

    def carry_in_device(self, ends, block_lengths, rank, xin, dxin, x0=None, dx0=None):
        """This block's true carry-in from the all-gathered block ends ``ends`` [G,N,L,4,d] (torch CUDA tensors throughout):
        z_in(g+1) = T(n_g) z_in(g) + ends_g chained over g < rank; written to xin [N,L,d], dxin [N,L,3,d]."""
        G = len(block_lengths)
        lens = (ctypes.c_longlong * G)(*[int(b) for b in block_lengths])
        N = xin.shape[0]
        self._check(self._lib.moihgp_cuda_carry_in_dev(self._h, _ptr(ends), G, lens, int(rank), N, _ptr(x0), _ptr(dx0), _ptr(xin), _ptr(dxin)))

    def objective_finish_device(self, Y, loss, grad, x0=None, dx0=None):
        """Step 2: loss / grad of the block from its true carry-in (torch CUDA tensors), reusing step 1's projection."""
        N, T, _ = Y.shape
        self._check(self._lib.moihgp_cuda_objective_finish_dev(self._h, _ptr(Y), N, T, _ptr(x0), _ptr(dx0), _ptr(loss), _ptr(grad), None, None))

    def block_transition(self, n):
        """(AKHA^n [L,d,d], E_k(n) [L,3,d,d]): how a block of n steps maps its carry-in (time-sharded evaluation)."""
        L, d = self.num_latent, self.igp_dim
        out = np.zeros((L, 4, d, d))
        self._check(self._lib.moihgp_cuda_block_transition(self._h, int(n), _ptr(out)))
        return out[:, 0].copy(), out[:, 1:].copy()

    def smoother_power(self, n, mode=SMOOTH_REFERENCE_LITERAL):
        """G[mode]^n per latent, [L,d,d]: how a backward value crosses a block of n steps (time-sharded smoother)."""
        L, d = self.num_latent, self.igp_dim
        out = np.zeros((L, d, d))
        self._check(self._lib.moihgp_cuda_smoother_power(self._h, int(mode), int(n), _ptr(out)))
        return out

    def fsn_block(self, phase, Y, seq_end, smoother_mode=SMOOTH_REFERENCE_LITERAL, x0=None, u_after=None, b_end=None, X=None, Xs=None, nll=None, xT=None):
        """One phase (1, 2, 3) of the fused pass on one block of a sequence sharded in time (torch CUDA tensors).
        Phase 1 returns (x_end [N,L,d], u_first [N,L]), phase 2 returns b_start [N,L,d], phase 3 fills X / Xs / nll / xT."""
        N, T, _ = Y.shape
        L, d = self.num_latent, self.igp_dim
        host = np.zeros((N, L, d + 1)) if phase == 1 else (np.zeros((N, L, d)) if phase == 2 else None)
        self._lib.moihgp_cuda_set_stream(self._h, _torch_stream(Y.device))
        self._check(self._lib.moihgp_cuda_fsn_block_dev(self._h, int(phase), _ptr(Y), N, T, 1 if seq_end else 0, int(smoother_mode), _ptr(x0),
                                                        _ptr(u_after), _ptr(b_end), _ptr(X), _ptr(Xs), _ptr(nll), _ptr(xT), _ptr(host)))
        if phase == 1:
            return host[:, :, :d].copy(), host[:, :, d].copy()
        return host

    def fsn_block_async(self, phase, Y, seq_end, smoother_mode=SMOOTH_REFERENCE_LITERAL, x0=None, u_after=None, b_end=None, X=None, Xs=None,
                        nll=None, xT=None, out=None):
        """fsn_block without a host round trip: the phase's output goes to the torch CUDA tensor ``out`` ([N,L,d+1] after
        phase 1, [N,L,d] after phase 2); asynchronous on torch's current stream."""
        N, T, _ = Y.shape
        self._lib.moihgp_cuda_set_stream(self._h, _torch_stream(Y.device))
        self._check(self._lib.moihgp_cuda_fsn_block_async(self._h, int(phase), _ptr(Y), N, T, 1 if seq_end else 0, int(smoother_mode), _ptr(x0),
                                                          _ptr(u_after), _ptr(b_end), _ptr(X), _ptr(Xs), _ptr(nll), _ptr(xT), _ptr(out)))

    def fsn_carry_device(self, direction, gathered, block_lengths, rank, out, smoother_mode=SMOOTH_REFERENCE_LITERAL, x0=None, u_after=None):
        """The carry exchanges of the time-sharded pass on the device (torch CUDA tensors): direction 0 turns the all-gathered
        phase-1 outputs [G,N,L,d+1] into this block's x_in (``out`` [N,L,d]) and ``u_after`` [N,L]; direction 1 turns the
        all-gathered phase-2 outputs [G,N,L,d] into this block's b_end (``out`` [N,L,d]; untouched on the last block)."""
        G = len(block_lengths)
        lens = (ctypes.c_longlong * G)(*[int(b) for b in block_lengths])
        self._check(self._lib.moihgp_cuda_fsn_carry_dev(self._h, int(direction), int(smoother_mode), _ptr(gathered), G, lens, int(rank), out.shape[0],
                                                        _ptr(x0), _ptr(out), _ptr(u_after)))

    def smooth(self, X, smoother_mode=SMOOTH_REFERENCE_LITERAL):
        """IHGP::backwardSmoother (ihgp.h:103-114) of every latent over caller-supplied filtered states X [N,T,L,d] (or
        [T,L,d]); returns Xs of the same shape.  Gains / covariances: ``smoother_consts``."""
        X = _np(X)
        squeeze = X.ndim == 3
        if squeeze:
            X = X[None]
        N, T, L, d = X.shape
        assert L == self.num_latent and d == self.igp_dim
        Xs = np.empty_like(X)
        self._check(self._lib.moihgp_cuda_smooth(self._h, _ptr(X), N, T, int(smoother_mode), _ptr(Xs)))
        return Xs[0] if squeeze else Xs

    # ---- device-resident streaming learner (moihgp_cuda_online_*) -----------------------------------------------
    def online_begin(self, windowsize):
        """Window, moving mean, carried state and proximal term in device memory; False if the shape does not qualify."""
        return self._lib.moihgp_cuda_online_begin(self._h, int(windowsize)) == 0

    def online_push(self, y, ma_given=None, want_ma=False):
        """OnlineObjective::push_back (moihgp_online.h:75-93); ``ma_given``: the caller's own centre instead of the window mean."""
        y = _np(y)
        ma_given = None if ma_given is None else _np(ma_given)
        ma = np.empty(self.num_output) if want_ma else None
        self._check(self._lib.moihgp_cuda_online_push(self._h, _ptr(y), _ptr(ma_given), _ptr(ma)))
        return ma

    def online_set_proximal(self, oldparams=None, B=None):
        """1/2 dparams' B dparams around ``oldparams`` (B None: identity); ``oldparams`` None: no proximal term on the device."""
        oldparams = None if oldparams is None else _np(oldparams)
        B = None if B is None else _np(B)
        self._check(self._lib.moihgp_cuda_online_set_proximal(self._h, _ptr(oldparams), _ptr(B)))

    def online_objective(self, params):
        """OnlineObjective::operator() (moihgp_online.h:40-72) at ``params`` on the resident window: one CUDA-graph launch."""
        params = _np(params)
        loss, grad = np.zeros(1), np.zeros(self.num_param)
        self._check(self._lib.moihgp_cuda_online_objective(self._h, _ptr(params), _ptr(loss), _ptr(grad)))
        return float(loss[0]), grad

    def online_state(self):
        L, d = self.num_latent, self.igp_dim
        x, dx = np.zeros((L, d)), np.zeros((L, 3, d))
        self._check(self._lib.moihgp_cuda_online_get_state(self._h, _ptr(x), _ptr(dx)))
        return x, dx

    def bind(self, Y):
        """Copy the observations to the device once; ``objective_bound`` then evaluates on them at the current parameters
        (the L-BFGS loop calls the objective tens of times on the same data).  ``bind(None)`` releases them."""
        if Y is None:
            self._check(self._lib.moihgp_cuda_bind_data(self._h, None, 0, 0))
            self._bound = None
            return
        Y = _np(Y)
        if Y.ndim == 2:
            Y = Y[None]
        N, T, p = Y.shape
        assert p == self.num_output
        self._check(self._lib.moihgp_cuda_bind_data(self._h, _ptr(Y), N, T))
        self._bound = (N, T)

    def objective_bound(self, x0=None, dx0=None):
        if getattr(self, "_bound", None) is None:
            raise RuntimeError("no data bound: call bind(Y) first")
        N, T = self._bound
        L, d = self.num_latent, self.igp_dim
        x0 = None if x0 is None else _np(x0).reshape(N, L, d)
        dx0 = None if dx0 is None else _np(dx0).reshape(N, L, 3, d)
        loss = np.zeros(1)
        grad = np.zeros(self.num_param)
        self._check(self._lib.moihgp_cuda_objective_bound(self._h, _ptr(x0), _ptr(dx0), _ptr(loss), _ptr(grad), None, None))
        return float(loss[0]), grad

    def objective_device(self, Y, loss, grad, x0=None, dx0=None, xT=None, dxT=None):
        import torch
        assert Y.is_cuda and Y.dtype == torch.float64 and Y.is_contiguous() and Y.dim() == 3
        N, T, _ = Y.shape
        self.set_stream(_torch_stream(Y.device))
        self._check(self._lib.moihgp_cuda_objective_dev(self._h, _ptr(Y), N, T, _ptr(x0), _ptr(dx0), _ptr(loss), _ptr(grad), _ptr(xT), _ptr(dxT)))
