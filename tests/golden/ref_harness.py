"""Harness that makes the UNMODIFIED reference (compops/qnmh-sysid2018) importable in the
build container, so that golden vectors can be generated from it.

Only usable where /root/reference exists (the build container).  Nothing under tests/ that is
marked `gpu`, nor smoke()/bench.py, imports this module.  It never copies reference sources into
the repository: a scratch copy goes to /tmp, where the four dtype aliases that no longer exist
in numpy 2 are patched in `resampling.pyx` (shim 1) and the module is
cythonized.  The remaining shims (stub modules for quandl/matplotlib/palettable, `np.mat`,
import order against `warnings.filterwarnings("error")`) are applied in-process.
"""
import os
import shutil
import subprocess
import sys
import types
import warnings

import numpy as np

REFERENCE_ROOT = "/root/reference"
SCRATCH = "/tmp/qnmh_ref_scratch"


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "python", "state"))


def _build_scratch_copy():
    pydir = os.path.join(SCRATCH, "python")
    marker = os.path.join(SCRATCH, ".built")
    if os.path.exists(marker):
        return pydir
    if os.path.isdir(SCRATCH):
        shutil.rmtree(SCRATCH)
    os.makedirs(SCRATCH)
    shutil.copytree(os.path.join(REFERENCE_ROOT, "python"), pydir)
    shutil.copytree(os.path.join(REFERENCE_ROOT, "data"), os.path.join(SCRATCH, "data"))
    pyx = os.path.join(pydir, "state", "particle_methods", "resampling.pyx")
    src = open(pyx).read()
    # numpy>=1.24 removed np.int / np.float and the matching C typedefs (resampling.pyx:27-31).
    src = src.replace("DTYPE_int = np.int\n", "DTYPE_int = np.int64\n")
    src = src.replace("ctypedef np.int_t DTYPE_int_t", "ctypedef np.int64_t DTYPE_int_t")
    src = src.replace("DTYPE_float = np.float\n", "DTYPE_float = np.float64\n")
    src = src.replace("ctypedef np.float_t DTYPE_float_t", "ctypedef np.float64_t DTYPE_float_t")
    open(pyx, "w").write(src)
    setup = (
        "from setuptools import setup\n"
        "from Cython.Build import cythonize\n"
        "import numpy\n"
        "setup(ext_modules=cythonize(['state/particle_methods/resampling.pyx',"
        "'state/particle_methods/cython_lgss_helper.pyx',"
        "'state/particle_methods/cython_sv_helper.pyx',"
        "'state/particle_methods/cython_sv_leverage_helper.pyx',"
        "'state/kalman_methods/cython_helper.pyx'], language_level=2),"
        " include_dirs=[numpy.get_include()])\n"
    )
    open(os.path.join(pydir, "_setup_golden.py"), "w").write(setup)
    subprocess.run([sys.executable, "_setup_golden.py", "build_ext", "--inplace"],
                   cwd=pydir, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    open(marker, "w").write("ok\n")
    return pydir


def _install_stubs():
    def stub(name, **attrs):
        mod = types.ModuleType(name)
        for key, val in attrs.items():
            setattr(mod, key, val)
        sys.modules.setdefault(name, mod)
        return sys.modules[name]

    stub("quandl")
    stub("matplotlib")
    stub("matplotlib.pylab")
    qual = stub("palettable.colorbrewer.qualitative", Dark2_8=None)
    brewer = stub("palettable.colorbrewer", qualitative=qual)
    stub("palettable", colorbrewer=brewer)
    if not hasattr(np, "mat"):
        np.mat = lambda a: np.atleast_2d(np.asarray(a))


_loaded = {}


def load_reference():
    """Returns a dict of the reference classes used for golden generation."""
    if _loaded:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference not mounted at " + REFERENCE_ROOT)
    pydir = _build_scratch_copy()
    _install_stubs()
    if pydir not in sys.path:
        sys.path.insert(0, pydir)
    os.chdir(pydir)  # the reference uses data paths relative to python/
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        # import order matters: modules with `is "literal"` must be byte-compiled before
        # metropolis_hastings.py:42 turns warnings into errors (shim 4)
        from state.particle_methods.standard import ParticleMethods
        from state.particle_methods import resampling
        from state.kalman_methods.standard import KalmanMethods
        import parameter.mcmc.hessian_estimation  # noqa: F401
        import parameter.mcmc.quasi_newton.main  # noqa: F401
        import parameter.mcmc.quasi_newton.bfgs  # noqa: F401
        import parameter.mcmc.output  # noqa: F401
        from parameter.mcmc.metropolis_hastings import MetropolisHastings
        from models.linear_gaussian_model import LinearGaussianModel
        from models.stochastic_volatility_model import StochasticVolatilityModel
        from models.stochastic_volatility_model_leverage import StochasticVolatilityModelLeverage
        from state.particle_methods.cython_lgss import ParticleMethodsCythonLGSS
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    warnings.filterwarnings("ignore", category=PendingDeprecationWarning)
    _loaded.update(ParticleMethods=ParticleMethods, KalmanMethods=KalmanMethods,
                   MetropolisHastings=MetropolisHastings, resampling=resampling,
                   LinearGaussianModel=LinearGaussianModel,
                   StochasticVolatilityModel=StochasticVolatilityModel,
                   StochasticVolatilityModelLeverage=StochasticVolatilityModelLeverage,
                   ParticleMethodsCythonLGSS=ParticleMethodsCythonLGSS,
                   data_dir=os.path.join(SCRATCH, "data", "linear_gaussian_model"))
    return _loaded
