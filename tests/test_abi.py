"""The C-ABI library loads and exports every symbol include/pwicp.h declares (no GPU needed)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b((?:pwicp_|PiecewiseICP_)\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import pwicp_b200
    lib = pwicp_b200.load_library()
    syms = declared_symbols("pwicp.h")
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"libpwicp.so does not export {s}"
    assert sorted(pwicp_b200.EXPORTS) == syms


def test_no_cpu_fallback_without_device():
    import pwicp_b200
    lib = pwicp_b200.load_library()
    if lib.pwicp_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(pwicp_b200.PwicpError) as e:
        pwicp_b200.Context(0)
    assert e.value.status == -1
    assert "no CUDA device" in str(e.value)


def test_missing_extension_fails_loudly(tmp_path):
    import pwicp_b200
    with pytest.raises(ImportError):
        pwicp_b200.load_library(str(tmp_path / "libpwicp.so"))


def test_product_never_touches_the_oracle():
    """The product tree must not import, link or mention anything under oracle/."""
    pkg = os.path.join(ROOT, "piecewise-icp_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".cu", ".cuh", ".cpp", ".h", ".hpp", ".py", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle_py" not in txt and "liboracle" not in txt and "pwicp_oracle" not in txt, f


def test_host_helpers_match_oracle(oracle):
    import numpy as np
    import pwicp_b200 as P
    from pwicp_b200 import synth
    rng = np.random.default_rng(3)
    for _ in range(20):
        T = synth.rigid_matrix(*rng.uniform(-0.5, 0.5, 3), *rng.uniform(-2, 2, 3)).astype(np.float32)
        assert np.array_equal(P.matrix2angle(T), oracle.matrix2angle(T))
        bb = np.sort(rng.uniform(-10, 10, (2, 3)), axis=0).reshape(6)
        assert P.bbox_corner_change(bb, T) == oracle.bbox_corner_change(bb, T)
        B = synth.rigid_matrix(*rng.uniform(-0.5, 0.5, 6)).astype(np.float32)
        assert np.array_equal(P.mat4_mul(T, B), oracle.mat4_mul(T, B))
    # gimbal branch of matrix2angle (src/CommonFunc.cpp:388-399)
    G = np.eye(4, dtype=np.float32); G[:3, :3] = [[0, 0, 1], [0, 1, 0], [-1, 0, 0]]
    assert np.array_equal(P.matrix2angle(G), oracle.matrix2angle(G))


def test_drivers_have_no_cpu_fallback():
    """The reference-shaped drivers (host/Registration.cpp) reach the device for every hot-path and pre-processing step: the
    host-only statements kept for device-less tools (PCpreprocessingHost / SORfilterHost, c_hooks.cpp) are never called
    from them, and a missing CUDA device is fatal (pwicpHostContext), not a switch to CPU code."""
    reg = open(os.path.join(ROOT, "piecewise-icp_b200", "host", "Registration.cpp")).read()
    assert "PCpreprocessingHost" not in reg and "SORfilterHost" not in reg
    assert "PCpreprocessing(" in reg and "pwicp_piecewise_icp" in reg
    cf = open(os.path.join(ROOT, "piecewise-icp_b200", "host", "CommonFunc.cpp")).read()
    i = cf.index("pwicp_ctx* pwicpHostContext()")
    body = cf[i:i + 900]
    assert "pwicp_ctx_create" in body and ("exit(EXIT_FAILURE)" in body or "std::exit" in body)
