"""The C-ABI library builds, loads and exports every symbol include/pfb200.h declares (no compute
calls: there is no GPU in the CPU test tier), and refuses to run without a device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from qnmh_sysid2018_b200 import _lib
    return _lib.load()


def test_every_header_symbol_is_exported_and_bound(lib):
    from qnmh_sysid2018_b200 import _lib
    header = open(os.path.join(ROOT, "include", "pfb200.h")).read()
    declared = set(re.findall(r"\b(pfb200_[a-z0-9_]+)\s*\(", header))
    declared.discard("pfb200_handle")
    assert declared, "no declarations parsed from the header"
    for name in sorted(declared):
        assert hasattr(lib, name), "libpfb200.so does not export " + name
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))


def test_abi_version_and_model_table(lib):
    assert lib.pfb200_abi_version() == 1
    assert [lib.pfb200_model_no_params(m) for m in (0, 1, 2)] == [4, 3, 4]
    assert lib.pfb200_model_no_params(7) < 0


def test_no_cpu_fallback(lib):
    """Without a CUDA device the product path must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    rc = lib.pfb200_create(-1, 0, 0, ctypes.byref(h))
    assert rc == -2  # PFB200_ENODEVICE
    assert b"no CPU fallback" in lib.pfb200_last_error()
    from qnmh_sysid2018_b200 import ParticleMethodsB200, LinearGaussianModel
    m = LinearGaussianModel()
    m.set_observations([0.0, 0.1, -0.2])
    with pytest.raises(Exception) as err:
        ParticleMethodsB200({"no_particles": 16}).filter(m)
    assert "no CUDA device" in str(err.value)


def test_product_package_never_imports_the_oracle():
    """Only tests/, smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(ROOT, "qnmh_sysid2018_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pf_oracle" not in text and "import oracle" not in text, f


def test_integration_doc_quotes_the_executed_binding():
    """INTEGRATION.md section 2 shows examples/flps_binding.py verbatim (the GPU tier executes that file)."""
    snippet = open(os.path.join(ROOT, "examples", "flps_binding.py")).read()
    body = snippet.split("# --- snippet begin\n")[1].split("# --- snippet end")[0]
    assert body.strip() in open(os.path.join(ROOT, "INTEGRATION.md")).read()


def test_product_defaults_to_the_cuda_engine():
    """The `engine=` / `engine_factory=` seams exist for the CPU test tier only: left alone, both classes
    construct the CUDA engine (and fail loudly without a device) -- never an oracle-backed one."""
    import inspect
    from qnmh_sysid2018_b200 import batch, particle_methods
    src = inspect.getsource(particle_methods.ParticleMethodsB200._get_engine)
    assert "from .engine import Engine" in src and "oracle" not in src
    src = inspect.getsource(batch.BatchedParticleFilter.__init__)
    assert "from .engine import Engine" in src and "oracle" not in src
    assert particle_methods.ParticleMethodsB200()._engine is None
