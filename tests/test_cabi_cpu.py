"""No-GPU checks of the drop-in boundary: the C-ABI library loads, exports every
symbol include/fast_rnnt_b200.h declares, validates arguments before touching
CUDA, and the product never imports the oracle."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fast_rnnt_b200.h")
PKG = os.path.join(ROOT, "tf-fast-rnnt_b200")
LIB = os.path.join(PKG, "lib", "libfast_rnnt_b200.so")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(frn_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(PKG, "csrc")])
    return ctypes.CDLL(LIB)


def test_library_exports_every_declared_symbol(lib):
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"


def test_python_binding_covers_the_header():
    sys.path.insert(0, PKG)
    from tf_fast_rnnt import _lib
    assert sorted(_lib.EXPORTS) == declared_functions()


def test_version_and_status_strings(lib):
    assert lib.frn_version() == 100
    lib.frn_status_string.restype = ctypes.c_char_p
    assert lib.frn_status_string(0) == b"ok"
    assert b"invalid" in lib.frn_status_string(-1)
    assert b"workspace" in lib.frn_status_string(-2)


def test_workspace_queries(lib):
    for f in ("frn_mi_workspace_bytes", "frn_simple_loss_workspace_bytes", "frn_pruned_loss_workspace_bytes",
              "frn_simple_logprobs_workspace_bytes", "frn_pruned_logprobs_workspace_bytes"):
        getattr(lib, f).restype = ctypes.c_size_t
    B, S, T, C, R = 32, 100, 500, 500, 5
    mi = lib.frn_mi_workspace_bytes(B, S, T, T + 1)
    # diagonal-major planes of [B][Dn][P] cells: arcs (16 B), alpha (8 B), backward operands (16 B)
    P, Dn = 128, 608
    assert mi == (16 + 8 + 16) * B * Dn * P
    assert lib.frn_simple_loss_workspace_bytes(B, S, T, C) > mi
    assert lib.frn_pruned_loss_workspace_bytes(B, S, T, R) > mi
    assert lib.frn_mi_workspace_bytes(0, S, T, T + 1) == 0
    assert lib.frn_prune_ranges_width(100, 5) == 5
    assert lib.frn_prune_ranges_width(3, 5) == 4      # s_range > S -> S + 1 (rnnt_loss.py:710-711)


def test_argument_validation_needs_no_gpu(lib):
    null = ctypes.c_void_p(0)
    EINVAL, EWORKSPACE = -1, -2
    # null tensors / bad shapes are rejected before any CUDA call
    assert lib.frn_mi_fwd_bwd(null, null, null, 2, 3, 4, 5, 1, null, null, null, null, 0, null) == EINVAL
    assert lib.frn_mi_fwd_bwd(null, null, null, 2, 3, 4, 7, 1, null, null, null, null, 0, null) == EINVAL
    one = ctypes.c_void_p(256)
    assert lib.frn_mi_fwd_bwd(one, one, one, 2, 3, 4, 5, 0, one, null, null, null, 0, null) == EWORKSPACE
    assert lib.frn_prune_ranges(one, one, one, 1, 0, 4, 5, 2, one, one, ctypes.c_size_t(1 << 20), null) == EINVAL
    assert lib.frn_reduce(one, 4, 7, ctypes.c_float(0), one, null) == EINVAL
    assert lib.frn_simple_loss(one, one, one, one, 1, 4, 8, 16, 99, 0, 0, ctypes.c_float(0), ctypes.c_float(0),
                               ctypes.c_float(0), 0, one, null, null, one, ctypes.c_size_t(1 << 30), null) == EINVAL


def test_product_does_not_reach_the_oracle_or_a_cpu_fallback():
    bad = re.compile(r"^\s*(from|import)\s+oracle\b|rnnt_oracle|liborc", re.M)
    for base, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                text = open(os.path.join(base, f)).read()
                assert not bad.search(text), f"{f} references the oracle"
    src = open(os.path.join(PKG, "tf_fast_rnnt", "_lib.py")).read()
    assert "no CPU fallback" in src and "raise ImportError" in src


def test_public_api_names_match_the_reference():
    sys.path.insert(0, PKG)
    import tf_fast_rnnt
    # tf_fast_rnnt/python/tf_fast_rnnt/__init__.py:24-36,42,151
    for n in ("do_rnnt_pruning", "get_rnnt_logprobs", "get_rnnt_logprobs_joint", "get_rnnt_logprobs_pruned",
              "get_rnnt_logprobs_smoothed", "get_rnnt_prune_ranges", "rnnt_loss", "rnnt_loss_pruned",
              "rnnt_loss_simple", "rnnt_loss_smoothed", "mutual_information_recursion", "cummin"):
        assert callable(getattr(tf_fast_rnnt, n)), n
    assert tf_fast_rnnt.__version__ == "1.2"
    import inspect
    sig = inspect.signature(tf_fast_rnnt.rnnt_loss_simple)
    assert list(sig.parameters)[:9] == ["lm", "am", "symbols", "termination_symbol", "boundary", "rnnt_type",
                                        "delay_penalty", "reduction", "calc_gradients"]
    sig = inspect.signature(tf_fast_rnnt.rnnt_loss_pruned)
    assert list(sig.parameters)[:9] == ["logits", "symbols", "ranges", "termination_symbol", "boundary",
                                        "rnnt_type", "delay_penalty", "reduction", "calc_gradients"]
    sig = inspect.signature(tf_fast_rnnt.rnnt_loss_smoothed)
    assert sig.parameters["lm_only_scale"].default == 0.1 and sig.parameters["am_only_scale"].default == 0.1


def test_tf_shim_type_checks_against_the_c_header():
    """The TensorFlow op shim cannot be built here (no TensorFlow).  It is parsed and type-checked
    with g++ -fsyntax-only against a minimal mock of the TF op API (tests/tf_mock) and the REAL
    include/fast_rnnt_b200.h, so every frn_* call in it has the right arity and argument types."""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    assert gxx, "g++ is part of the image"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cuda_inc = "/usr/local/cuda/include"
    cmd = [gxx, "-std=c++17", "-fsyntax-only", "-Wall", "-Wno-comment", "-I", os.path.join(root, "include"),
           "-I", os.path.join(root, "tests", "tf_mock"), "-I", cuda_inc,
           os.path.join(root, "tf-fast-rnnt_b200", "tf_shim", "tf_fast_rnnt_b200_ops.cc")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]


def test_tf_frontend_uses_only_registered_ops():
    """Every _ops.<name> the TensorFlow front-end calls is an op the shim registers
    (TensorFlow exposes REGISTER_OP("FastRnntFooBar") as fast_rnnt_foo_bar)."""
    import ast
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    shim = open(os.path.join(root, "tf-fast-rnnt_b200", "tf_shim", "tf_fast_rnnt_b200_ops.cc")).read()
    front = open(os.path.join(root, "tf-fast-rnnt_b200", "tf_shim", "tf_frontend.py")).read()
    ast.parse(front)

    def snake(name):
        s = re.sub(r"([A-Z]+)([A-Z][a-z])", r"\1_\2", name)
        return re.sub(r"([a-z0-9])([A-Z])", r"\1_\2", s).lower()

    registered = {snake(n) for n in re.findall(r'REGISTER_OP\("(\w+)"\)', shim)}
    used = set(re.findall(r"_ops\.(\w+)\(", front))
    assert used and used <= registered, sorted(used - registered)
    # keyword arguments of every call are attributes of that op, positional arguments do not exceed its inputs
    ops_def = {}
    for m in re.finditer(r'REGISTER_OP\("(\w+)"\)(.*?);', shim, re.S):
        body = m.group(2)
        ops_def[snake(m.group(1))] = (len(re.findall(r'\.Input\("', body)),
                                      set(re.findall(r'\.Attr\("(\w+):', body)))
    for node in ast.walk(ast.parse(front)):
        if (isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute)
                and isinstance(node.func.value, ast.Name) and node.func.value.id == "_ops"):
            n_in, attrs = ops_def[node.func.attr]
            assert len(node.args) <= n_in, (node.func.attr, len(node.args), n_in)
            for kw in node.keywords:
                if kw.arg is not None:
                    assert kw.arg in attrs, (node.func.attr, kw.arg)
    # the reference's public names (tf_fast_rnnt/__init__.py:24-33, 42, 151) all exist in the front-end
    for name in ("do_rnnt_pruning", "get_rnnt_logprobs", "get_rnnt_logprobs_joint", "get_rnnt_logprobs_pruned",
                 "get_rnnt_logprobs_smoothed", "get_rnnt_prune_ranges", "rnnt_loss", "rnnt_loss_pruned",
                 "rnnt_loss_simple", "rnnt_loss_smoothed", "mutual_information_recursion", "cummin"):
        assert re.search(rf"^def {name}\(", front, re.M), name


def test_product_library_never_reads_the_environment(lib):
    """The FRN_* overrides (force one of two implementations of a stage) live only in the
    -DFRN_DEBUG_HOOKS build; the product library does not import getenv and carries none of the names."""
    dbg = LIB[:-3] + "_dbg.so"
    assert os.path.exists(dbg), "make builds both libraries"
    und = subprocess.run(["nm", "-D", "--undefined-only", LIB], capture_output=True, text=True, check=True).stdout
    assert "getenv" not in und
    blob = open(LIB, "rb").read()
    for name in (b"FRN_DP_CHAIN", b"FRN_DP_SCAN", b"FRN_BAND_DENSE", b"FRN_SIMPLE_SIMT", b"FRN_SCAN_K", b"FRN_RPL"):
        assert name not in blob, name
    und = subprocess.run(["nm", "-D", "--undefined-only", dbg], capture_output=True, text=True, check=True).stdout
    assert "getenv" in und


def test_nvtx_ranges_only_in_the_debug_hooks_build(lib):
    """every entry point that enqueues work opens an NVTX range - in the -DFRN_NVTX build
    (the debug-hooks library); the product library carries no tracing code."""
    dbg = LIB[:-3] + "_dbg.so"
    assert b"NVTX_INJECTION64_PATH" in open(dbg, "rb").read()          # nvtx3's lazy injection loader
    assert b"NVTX_INJECTION64_PATH" not in open(LIB, "rb").read()
