"""Run the REFERENCE's own source files for the hot path inside this container.

The reference cannot be imported as shipped (paramz / pathos / cma / matplotlib are absent and
GPy/__init__.py needs numpy.testing.Tester).  This harness loads the real
files of the path straight from /root/reference with stand-ins for exactly those missing pieces:

  * paramz (parameter bookkeeping / caching only, no arithmetic on the path): Param = ndarray
    subclass, Parameterized/Model = plain object bases, Cache_this = identity decorator;
  * the package __init__ files of GPy / GPyOpt (they import every model family, plotting, ...) are
    replaced by empty packages whose __path__ points at the real directories, so submodule imports
    resolve to the reference's real files;
  * GPy.kern.src.kern.Kern (a paramz Parameterized with a slicing metaclass that is the identity when
    all input dimensions are active), GPy.likelihoods.likelihood.Likelihood, link functions: minimal bases;
  * GPy.kern.src.stationary_cython: ctypes binding of the reference's C routine compiled from
    GPy/kern/src/stationary_utils.c into oracle/_ref (the Cython .pyx is a 1:1 wrapper of it);
  * pathos ProcessingPool: a serial map (the pool computes the same numbers, quirk q1).

Everything that does arithmetic is the reference's code, unmodified: GPy/util/linalg.py,
GPy/kern/src/{stationary,rbf,se}.py, GPy/inference/latent_function_inference/{posterior,
exact_gaussian_inference}.py, GPy/core/gp.py, GPy/likelihoods/gaussian.py, GPy/util/normalizer.py,
GPy/models/gp_regression.py, GPyOpt/models/{gpmodel,gpmodel_fixed_hyps}.py, GPyOpt/acquisitions/base.py,
multi_outputGP.py, uEI_noiseless.py, uPI.py, maEI.py, maPI.py, EI.py, PI.py, utility.py,
parameter_distribution.py.

Only tests/golden/make_golden.py uses this (to produce fixtures); nothing here runs on the GPU box.
"""
import configparser
import ctypes
import importlib.util
import os
import sys
import types

import numpy as np

REF = os.environ.get("BOCF_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _pkg(name, relpath=None, **attrs):
    m = types.ModuleType(name)
    m.__path__ = [os.path.join(REF, relpath)] if relpath else []
    m.__package__ = name
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], child, m)
    return m


def _mod(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], child, m)
    return m


def _load(name, relpath):
    """Execute the reference's real file `relpath` as module `name`."""
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], child, m)
    spec.loader.exec_module(m)
    return m


# ---- paramz stand-ins ------------------------------------------------------------------------------------
class Param(np.ndarray):
    def __new__(cls, name, input_array, default_constraint=None, *a, **kw):
        obj = np.atleast_1d(np.array(input_array, dtype=np.float64)).view(cls)
        obj.name = name
        obj.gradient = np.zeros(obj.shape)
        return obj

    def __array_finalize__(self, obj):
        self.name = getattr(obj, "name", None)
        self.gradient = getattr(obj, "gradient", None)

    @property
    def values(self):
        return np.asarray(self)

    def constrain_fixed(self, *a, **kw):
        pass

    def constrain_positive(self, *a, **kw):
        pass

    def set_prior(self, *a, **kw):
        pass


class ObsAr(np.ndarray):
    def __new__(cls, input_array):
        return np.atleast_1d(np.asarray(input_array, dtype=np.float64)).view(cls)

    def copy(self):
        return np.ndarray.copy(self).view(ObsAr)


class Parameterized(object):
    def __init__(self, name=None, *a, **kw):
        self.name = name
        self.parameters = []

    def link_parameter(self, p, index=None):
        self.parameters.append(p)

    def link_parameters(self, *ps):
        for p in ps:
            self.parameters.append(p)

    def unlink_parameter(self, p):
        pass

    def update_model(self, flag):
        # paramz re-runs parameters_changed when updates are switched back on (GP.set_XY relies on it)
        if flag and hasattr(self, "parameters_changed") and getattr(self, "posterior", "unset") != "unset":
            self.parameters_changed()


def Cache_this(limit=5, ignore_args=(), force_kwargs=()):
    def deco(f):
        return f
    return deco


class Logexp(object):
    pass


_installed = False


def install():
    """Create the stand-in modules and load the reference's real modules.  Idempotent."""
    global _installed
    if _installed:
        return sys.modules["_bocf_ref_ns"]
    if not os.path.isdir(REF):
        raise RuntimeError("reference checkout not found at %s" % REF)

    pz = _pkg("paramz", Param=Param, ObsAr=ObsAr, Parameterized=Parameterized, Model=Parameterized)
    _mod("paramz.caching", Cache_this=Cache_this)
    _mod("paramz.transformations", Logexp=Logexp)

    # ---- GPy ------------------------------------------------------------------------------------------
    _pkg("GPy", "GPy")
    _pkg("GPy.core", "GPy/core")
    class VariationalPosterior(object):
        pass
    _pkg("GPy.core.parameterization", "GPy/core/parameterization", Param=Param, Parameterized=Parameterized)
    _mod("GPy.core.parameterization.param", Param=Param)
    _mod("GPy.core.parameterization.parameterized", Parameterized=Parameterized)
    _mod("GPy.core.parameterization.variational", VariationalPosterior=VariationalPosterior)
    sys.modules["GPy.core"].Param = Param
    _mod("GPy.core.model", Model=Parameterized)
    class Mapping(object):
        pass
    _mod("GPy.core.mapping", Mapping=Mapping)

    util = _pkg("GPy.util", "GPy/util")
    cfg = configparser.ConfigParser()
    cfg.add_section("cython")
    cfg.set("cython", "working", "True")
    _mod("GPy.util.config", config=cfg)
    # GPy/util/linalg_cython.pyx:9-21 symmetrify is pure data movement (copy one triangle onto the other)
    def _symmetrify(A, upper):
        tri = np.triu_indices_from(A, k=1)
        if upper:
            A.T[tri] = A[tri]
        else:
            A[tri] = A.T[tri]
    _mod("GPy.util.linalg_cython", symmetrify=_symmetrify)
    _load("GPy.util.diag", "GPy/util/diag.py")
    _load("GPy.util.linalg", "GPy/util/linalg.py")
    cfg.set("cython", "working", "True")     # stationary.py consults the same flag for _grad_X (C path below)
    _load("GPy.util.normalizer", "GPy/util/normalizer.py")

    # kernels
    _pkg("GPy.kern", "GPy/kern")
    _pkg("GPy.kern.src", "GPy/kern/src")

    class Kern(Parameterized):
        _support_GPU = False

        def __init__(self, input_dim, active_dims, name, useGPU=False, *a, **kw):
            super(Kern, self).__init__(name=name)
            self.input_dim = int(input_dim)
            if active_dims is None:
                active_dims = np.arange(input_dim, dtype=np.int_)
            self.active_dims = np.atleast_1d(np.asarray(active_dims, np.int_))
            self._all_dims_active = self.active_dims
            self.useGPU = False

        @property
        def _effective_input_dim(self):
            return np.size(self._all_dims_active)

        def parameters_changed(self):
            pass

        def set_prior(self, *a, **kw):
            pass
    _mod("GPy.kern.src.kern", Kern=Kern)
    sys.modules["GPy.kern"].Kern = Kern
    class _Dummy(object):
        def __init__(self, *a, **kw):
            pass
    _mod("GPy.kern.src.psi_comp", PSICOMP_RBF=_Dummy, PSICOMP_RBF_GPU=_Dummy, PSICOMP_GH=_Dummy)
    _mod("GPy.kern.src.grid_kerns", GridRBF=_Dummy)

    # the reference's C routine, compiled from its own source (oracle/Makefile)
    so = os.path.join(ROOT, "oracle", "_ref", "libstationary_utils.so")
    if not os.path.exists(so):
        raise RuntimeError("run `make -C oracle` first (builds the reference's stationary_utils.c)")
    clib = ctypes.CDLL(so)
    dp = ctypes.POINTER(ctypes.c_double)
    clib._grad_X.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, dp, dp, dp]
    clib._lengthscale_grads.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, dp, dp, dp]

    def grad_X(N, D, M, X, X2, tmp, grad):          # GPy/kern/src/stationary_cython.pyx:17-27
        for a in (X, X2, tmp, grad):
            assert a.flags["C_CONTIGUOUS"] and a.dtype == np.float64
        clib._grad_X(N, D, M, X.ctypes.data_as(dp), X2.ctypes.data_as(dp), tmp.ctypes.data_as(dp),
                     grad.ctypes.data_as(dp))

    def lengthscale_grads(N, M, Q, tmp, X, X2, grad):   # stationary_cython.pyx:51-60
        tmp = np.ascontiguousarray(tmp)
        clib._lengthscale_grads(N, M, Q, tmp.ctypes.data_as(dp), X.ctypes.data_as(dp), X2.ctypes.data_as(dp),
                                grad.ctypes.data_as(dp))
    _mod("GPy.kern.src.stationary_cython", grad_X=grad_X, lengthscale_grads=lengthscale_grads)

    st = _load("GPy.kern.src.stationary", "GPy/kern/src/stationary.py")
    rbf = _load("GPy.kern.src.rbf", "GPy/kern/src/rbf.py")
    se = _load("GPy.kern.src.se", "GPy/kern/src/se.py")
    kern = sys.modules["GPy.kern"]
    kern.RBF, kern.SE, kern.Matern52, kern.Matern32 = rbf.RBF, se.SE, st.Matern52, st.Matern32

    # likelihood
    _pkg("GPy.likelihoods", "GPy/likelihoods")
    class GPTransformation(object):
        pass
    class Identity(GPTransformation):
        pass
    _mod("GPy.likelihoods.link_functions", GPTransformation=GPTransformation, Identity=Identity)

    class Likelihood(Parameterized):
        def __init__(self, gp_link, name):
            super(Likelihood, self).__init__(name)
            self.gp_link = gp_link
            self.log_concave = False
    _mod("GPy.likelihoods.likelihood", Likelihood=Likelihood)
    ga = _load("GPy.likelihoods.gaussian", "GPy/likelihoods/gaussian.py")
    lk = sys.modules["GPy.likelihoods"]
    lk.Likelihood, lk.Gaussian = Likelihood, ga.Gaussian
    lk.MixedNoise = type("MixedNoise", (), {})

    # inference
    _pkg("GPy.inference", "GPy/inference")
    class LatentFunctionInference(object):
        pass
    _pkg("GPy.inference.latent_function_inference", "GPy/inference/latent_function_inference",
         LatentFunctionInference=LatentFunctionInference)
    _load("GPy.inference.latent_function_inference.posterior", "GPy/inference/latent_function_inference/posterior.py")
    _load("GPy.inference.latent_function_inference.exact_gaussian_inference",
          "GPy/inference/latent_function_inference/exact_gaussian_inference.py")
    _mod("GPy.inference.latent_function_inference.expectation_propagation", EP=_Dummy)

    gp = _load("GPy.core.gp", "GPy/core/gp.py")
    sys.modules["GPy.core"].GP = gp.GP
    _pkg("GPy.models", "GPy/models")
    gpr = _load("GPy.models.gp_regression", "GPy/models/gp_regression.py")
    sys.modules["GPy.models"].GPRegression = gpr.GPRegression

    # ---- GPyOpt -----------------------------------------------------------------------------------------
    _pkg("GPyOpt", "GPyOpt")
    _pkg("GPyOpt.models", "GPyOpt/models")
    _load("GPyOpt.models.base", "GPyOpt/models/base.py")
    gm = _load("GPyOpt.models.gpmodel", "GPyOpt/models/gpmodel.py")
    gf = _load("GPyOpt.models.gpmodel_fixed_hyps", "GPyOpt/models/gpmodel_fixed_hyps.py")
    sys.modules["GPyOpt.models"].GPModel = gm.GPModel
    sys.modules["GPyOpt.models"].GPModelFixedHyps = gf.GPModelFixedHyps
    _pkg("GPyOpt.core", "GPyOpt/core")
    _pkg("GPyOpt.core.task", "GPyOpt/core/task")

    def constant_cost_withGradients(x):           # GPyOpt/core/task/cost.py:73-80
        return np.ones(x.shape[0])[:, None], np.zeros(x.shape)
    _mod("GPyOpt.core.task.cost", constant_cost_withGradients=constant_cost_withGradients)
    _pkg("GPyOpt.acquisitions", "GPyOpt/acquisitions")
    _load("GPyOpt.acquisitions.base", "GPyOpt/acquisitions/base.py")

    class ProcessingPool(object):
        def __init__(self, *a, **kw):
            pass

        def map(self, f, xs):
            return [f(x) for x in xs]
    _pkg("pathos")
    _mod("pathos.multiprocessing", ProcessingPool=ProcessingPool)

    ns = types.ModuleType("_bocf_ref_ns")
    for name in ("utility", "parameter_distribution", "multi_outputGP", "uEI_noiseless", "uPI", "maEI", "maPI",
                 "EI", "PI"):
        setattr(ns, name, _load(name, name + ".py"))
    ns.GPy = sys.modules["GPy"]
    ns.GPyOpt = sys.modules["GPyOpt"]
    ns.Param = Param
    sys.modules["_bocf_ref_ns"] = ns
    _installed = True
    return ns


# ---- building reference models with explicit hyper-samples ---------------------------------------------------
def make_kernel(ns, kind, d, variance, lengthscale):
    K = {"se": ns.GPy.kern.SE, "rbf": ns.GPy.kern.RBF, "matern52": ns.GPy.kern.Matern52,
         "matern32": ns.GPy.kern.Matern32}[kind]
    return K(d, variance=variance, lengthscale=np.asarray(lengthscale, dtype=float), ARD=True)


def make_reference_model(ns, kind, X, Y_list, variance, lengthscale, noise):
    """The reference's multi_outputGP with GPModel outputs whose hyper-sample instances are filled explicitly
    (stands in for GPModel.updateModel's ML-II + HMC, which is out of scope): instance i of output j is the
    reference's GPRegression(X, Y_j, kernel_ij, noise_var_ij), factorised by its own parameters_changed()."""
    H, m = variance.shape
    d = X.shape[1]
    model = ns.multi_outputGP.multi_outputGP(output_dim=m, n_samples=H, fixed_hyps=False)
    for j in range(m):
        out = model.output[j]                      # GPyOpt.models.GPModel (real class)
        out.model_instances = []
        for h in range(H):
            k = make_kernel(ns, kind, d, variance[h, j], lengthscale[h, j])
            g = ns.GPy.models.GPRegression(X, Y_list[j], kernel=k, noise_var=noise[h, j])
            g.parameters_changed()                 # paramz would trigger this at the end of __init__
            out.model_instances.append(g)
        out.model = out.model_instances[0]
        out.set_hyperparameters(0)
    return model
