"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Import shim that lets the *unmodified* reference sources under /root/reference be
imported in a container that has neither ``gpytorch`` nor ``pyro`` installed
(SURVEY.md Appendix C).  It is used by ``tests/golden/make_golden.py`` (run once,
in the build container, to produce the committed fixtures) and by the optional
``-m "not gpu"`` cross-checks that are skipped when /root/reference is absent
(e.g. on the GPU box).

What it does
  * registers a namespace stub for ``meta_learn`` so that
    ``meta_learn/__init__.py`` (which imports gpytorch at line 1) never runs;
  * registers inert ``gpytorch.*`` modules whose attributes are empty
    ``torch.nn.Module`` subclasses, enough for the ``class X(gpytorch...)``
    statements in meta_learn/models.py:406-601 to evaluate;
  * maps ``pyro.distributions.Normal(...).to_event(1)`` onto
    ``torch.distributions.Independent(Normal, 1)`` (random_gp.py:6,131-151,248);
  * replaces the one method whose arithmetic lives in gpytorch,
    ``VectorizedGP.forward`` (random_gp.py:54-89), by the dense-Cholesky
    restatement from ``oracle/pacoh_oracle.py``.

Nothing under /root/reference is modified or copied.
"""
import importlib
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("PACOH_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "meta_learn"))


class _StubModule(types.ModuleType):
    """Module whose every missing attribute is an inert nn.Module subclass."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (torch.nn.Module,), {"__module__": self.__name__})
        setattr(self, name, cls)
        return cls


def _install_gpytorch_stub():
    names = ["gpytorch", "gpytorch.means", "gpytorch.kernels", "gpytorch.functions", "gpytorch.utils",
             "gpytorch.utils.broadcasting", "gpytorch.likelihoods", "gpytorch.likelihoods.noise_models",
             "gpytorch.models", "gpytorch.models.approximate_gp", "gpytorch.variational",
             "gpytorch.distributions", "gpytorch.mlls", "gpytorch.settings"]
    mods = {}
    for n in names:
        m = _StubModule(n)
        m.__path__ = []
        mods[n] = m
        sys.modules[n] = m
    for n, m in mods.items():
        if "." in n:
            parent, child = n.rsplit(".", 1)
            setattr(mods[parent], child, m)
    mods["gpytorch.utils.broadcasting"]._mul_broadcast_shape = torch.broadcast_shapes


def _install_pyro_stub():
    import torch.distributions as td

    class Normal(td.Normal):
        def to_event(self, n):
            return td.Independent(self, n)

    class LogNormal(td.LogNormal):
        def to_event(self, n):
            return td.Independent(self, n)

    pyro = types.ModuleType("pyro")
    pyro.__path__ = []
    dist = types.ModuleType("pyro.distributions")
    dist.Normal, dist.LogNormal, dist.Independent = Normal, LogNormal, td.Independent
    pyro.distributions = dist
    sys.modules["pyro"] = pyro
    sys.modules["pyro.distributions"] = dist


_installed = False


def install():
    """Make ``meta_learn.{svgd,util,abstract,models,random_gp}``, ``config`` and
    ``experiments.data_sim`` importable from the read-only reference tree."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    try:
        import gpytorch  # noqa: F401  (a real install wins)
    except ImportError:
        _install_gpytorch_stub()
    try:
        import pyro  # noqa: F401
    except ImportError:
        _install_pyro_stub()
    pkg = types.ModuleType("meta_learn")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "meta_learn")]
    sys.modules["meta_learn"] = pkg
    _installed = True


def load_reference():
    """Returns a namespace with the live reference modules, VectorizedGP.forward patched."""
    install()
    ns = types.SimpleNamespace()
    ns.svgd = importlib.import_module("meta_learn.svgd")
    ns.util = importlib.import_module("meta_learn.util")
    ns.models = importlib.import_module("meta_learn.models")
    ns.random_gp = importlib.import_module("meta_learn.random_gp")
    ns.data_sim = importlib.import_module("experiments.data_sim")

    from oracle import pacoh_oracle as orc

    def forward(self, x_data, y_data, train=True, prior=False):
        # dense restatement of random_gp.py:54-89 (train-mode value only; eval mode is
        # covered by oracle.gp_posterior and is not reachable through gpytorch stubs)
        assert x_data.ndim == 3 and train and not prior
        mean = self.mean_nn(x_data).squeeze(-1) if self.mean_module_str == 'NN' \
            else self.constant_mean.expand(x_data.shape[:-1])
        feat = self.kernel_nn(x_data) if self.covar_module_str == 'NN' else x_data
        ls = torch.nn.functional.softplus(self.lengthscale_raw)
        noise = torch.nn.functional.softplus(self.noise_raw)
        mll = orc.mvn_mll(mean, orc.se_gram(feat, ls), noise.reshape(-1), y_data)
        return None, mll

    ns.random_gp.VectorizedGP.forward = forward
    return ns
