"""Per-observation model class with the reference's interface (moihgp/pywrapper.py:10-270).

Same constructor, methods, properties and array layouts as the reference's ``MOIHGP``; it binds
the same 13 ``gpXX_*`` symbols (pywrapper.py:28-145), here served by the B200 library.  The
reference's own pywrapper.py also works unchanged against lib/libmoihgp.so.
Two deviations, both documented in SURVEY.md Q7: ``kernel="Matern52"`` is usable (the reference
raises AttributeError at pywrapper.py:59), and whether it means Matern-5/2 is governed by
MOIHGP_GP52_MATERN52 exactly as in the C library.
"""
import numpy as np

from . import _lib

c_double_p = _lib.c_double_p


class MOIHGP(object):

    def __init__(self, dt, num_output, num_latent, kernel="Matern32", threading=False):
        self.dt = dt
        self.__num_output = num_output
        self.__num_latent = num_latent
        self.__lib = _lib.load()
        if kernel == "Matern32":
            pre = "gp32_"
        elif kernel == "Matern52":
            pre = "gp52_"
        else:
            raise NotImplementedError("Unsupported kernel type.")
        f = lambda n: getattr(self.__lib, pre + n)
        self.__obj = f("new")(dt, num_output, num_latent, threading)
        self.__del = f("del")
        self.__num_param = int(f("num_param")(self.__obj))
        self.__num_igp_param = int(f("num_igp_param")(self.__obj))
        self.__igp_dim = int(f("igp_dim")(self.__obj))
        self.__step1, self.__step2, self.__step3, self.__step4 = f("step1"), f("step2"), f("step3"), f("step4")
        self.__update, self.__lik1, self.__lik2, self.__get_params = f("update"), f("lik1"), f("lik2"), f("get_params")
        # caller-owned, pre-allocated buffers (pywrapper.py:146-167)
        self.__params = np.zeros((self.num_param,), dtype=np.float64)
        self.__grad = np.zeros((self.num_param,), dtype=np.float64)
        self.__x = np.zeros((self.num_latent, self.igp_dim), dtype=np.float64)
        self.__y = np.zeros((self.num_output,), dtype=np.float64)
        self.__dx = np.zeros((self.num_latent, self.num_igp_param, self.igp_dim), dtype=np.float64)
        self.__xnew = np.zeros((self.num_latent, self.igp_dim), dtype=np.float64)
        self.__yhat = np.zeros((self.num_output,), dtype=np.float64)
        self.__dxnew = np.zeros((self.num_latent, self.num_igp_param, self.igp_dim), dtype=np.float64)
        p = lambda a: a.ctypes.data_as(c_double_p)
        self.__params_p, self.__grad_p, self.__x_p, self.__y_p = p(self.__params), p(self.__grad), p(self.__x), p(self.__y)
        self.__dx_p, self.__xnew_p, self.__yhat_p, self.__dxnew_p = p(self.__dx), p(self.__xnew), p(self.__yhat), p(self.__dxnew)

    @property
    def handle(self):
        """the library handle (what gpXX_new returned): the moihgp_cuda_* entry points accept it too"""
        return self.__obj

    def __del__(self):
        if getattr(self, "_MOIHGP__obj", None):
            self.__del(self.__obj)
            self.__obj = None

    def step(self, x, y=None, dx=None):
        self.__x[...] = x
        if y is None:
            self.__step4(self.__obj, self.__x_p, self.__xnew_p, self.__yhat_p)
            return self.__xnew.astype(np.float64), self.__yhat.astype(np.float64)
        self.__y[...] = y
        if dx is None:
            self.__step3(self.__obj, self.__x_p, self.__y_p, self.__xnew_p, self.__yhat_p)
            return self.__xnew.astype(np.float64), self.__yhat.astype(np.float64)
        self.__dx[...] = dx
        self.__step1(self.__obj, self.__x_p, self.__y_p, self.__dx_p, self.__xnew_p, self.__yhat_p, self.__dxnew_p)
        return self.__xnew.astype(np.float64), self.__yhat.astype(np.float64), self.__dxnew.astype(np.float64)

    def update(self, params):
        self.__params[...] = params
        self.__update(self.__obj, self.__params_p)

    def negLogLikelihood(self, x, y, dx=None):
        self.__x[...] = x
        self.__y[...] = y
        if dx is None:
            return np.float64(self.__lik2(self.__obj, self.__x_p, self.__y_p))
        self.__dx[...] = dx
        res = self.__lik1(self.__obj, self.__x_p, self.__y_p, self.__dx_p, self.__grad_p)
        return np.float64(res), self.__grad.astype(np.float64)

    @property
    def num_output(self):
        return self.__num_output

    @property
    def num_latent(self):
        return self.__num_latent

    @property
    def igp_dim(self):
        return self.__igp_dim

    @property
    def num_param(self):
        return self.__num_param

    @property
    def num_igp_param(self):
        return self.__num_igp_param

    @property
    def params(self):
        self.__get_params(self.__obj, self.__params_p)
        return self.__params

    @property
    def covariance(self):
        # pywrapper.py:256-270
        params = self.params.copy()
        U = np.reshape(params[:self.num_output * self.num_latent], (self.num_output, self.num_latent))
        sqrtS = np.diag(np.sqrt(params[self.num_output * self.num_latent:(self.num_output + 1) * self.num_latent]))
        B = []
        igp_params = np.reshape(params[-self.num_latent * 3:], (self.num_latent, 3))
        for magnitude, lengthscale, _ in igp_params:
            B.append(magnitude ** 0.5 * (3 ** 0.5 / lengthscale ** 0.5) ** 1.5)
        B = np.diag(B)
        return U @ sqrtS @ B @ sqrtS @ U.T
