"""CPU suite: the C-ABI library loads, exports every symbol include/prost_b200.h declares, and
refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "prost_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(pb_[a-z0-9_]+)\s*\(", src))
    names -= {"pb_stopping_cb", "pb_interm_cb"}
    return sorted(names)


def test_header_declares_the_boundary():
    names = declared_symbols()
    assert len(names) > 70
    for must in ("pb_pdhg_create", "pb_backend_iterate", "pb_linop_eval", "pb_prox_eval", "pb_solver_solve"):
        assert must in names


def test_library_exports_every_declared_symbol():
    import prost_b200._capi as capi
    dll = ctypes.CDLL(capi.LIB_PATH)
    missing = [n for n in declared_symbols() if not hasattr(dll, n)]
    assert not missing, missing
    # and the ctypes table covers the header
    untyped = [n for n in declared_symbols() if n not in capi.SIGNATURES]
    assert not untyped, untyped


def test_options_defaults_match_reference_frontend():
    """matlab/+prost/+backend/pdhg.m:3-14, options.m:3-14, +backend/admm.m:3-13."""
    import prost_b200 as pb
    o = pb.pdhg_options()
    assert (o.tau0, o.sigma0, o.residual_iter, o.scale_steps_operator) == (1.0, 1.0, 1, 1)
    assert o.stepsize_variant == 4 and abs(o.arb_delta - 1.05) < 1e-6 and abs(o.arb_tau - 0.8) < 1e-6
    assert abs(o.arg_alpha0 - 0.5) < 1e-6 and abs(o.arg_nu - 0.95) < 1e-6 and abs(o.arg_delta - 1.5) < 1e-6
    s = pb.solver_options()
    assert s.max_iters == 1000 and s.num_cback_calls == 10 and abs(s.tol_rel_primal - 1e-4) < 1e-9
    a = pb.admm_options()
    assert a.rho0 == 1 and a.alpha == 1.7 and a.cg_max_iter == 10 and a.cg_tol_pow == 1.3


def test_function_names_follow_the_mex_registry():
    import prost_b200 as pb
    from prost_b200.api import FUNCTIONS_1D, function_id
    assert [function_id(n) for n in FUNCTIONS_1D] == list(range(14))
    with pytest.raises(pb.ProstError):
        function_id("nope")


def test_no_cpu_fallback_without_gpu():
    import prost_b200 as pb
    if pb.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pb.ProstError) as e:
        pb.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_public_headers_compile_standalone():
    """include/prost_b200.h is plain C (the drop-in boundary binds from C, cgo-style FFI and ctypes alike) and
    every include/prost/**/*.hpp shim compiles on its own with a host C++14 compiler (no nvcc, no thrust)."""
    import glob
    import shutil
    import subprocess
    inc = os.path.join(ROOT, "include")
    gcc, gxx = shutil.which("gcc"), shutil.which("g++")
    if not gcc or not gxx:
        pytest.skip("no host compiler")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c",
                        os.path.join(inc, "prost_b200.h")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for hpp in sorted(glob.glob(os.path.join(inc, "prost", "**", "*.hpp"), recursive=True)):
        rel = os.path.relpath(hpp, inc)
        r = subprocess.run([gxx, "-std=c++14", "-fsyntax-only", "-x", "c++", "-I", inc, "-"],
                           input=f'#include "{rel}"\nint main() {{ return 0; }}\n', capture_output=True, text=True)
        assert r.returncode == 0, f"{rel}:\n{r.stderr[:2000]}"
