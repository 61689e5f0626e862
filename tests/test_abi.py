"""The C-ABI library loads and exports every symbol include/ta_b200.h declares (no compute calls)."""

import ctypes
import os
import re

import pytest

from track_analyser_b200 import _native


def header_symbols(root):
    text = open(os.path.join(root, "include", "ta_b200.h")).read()
    return sorted(set(re.findall(r"TA_API\s+[\w\s\*]+?\b(ta_\w+)\s*\(", text)))


def test_library_is_built_and_exports_header_symbols(repo_root):
    assert os.path.exists(_native.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(_native.LIB_PATH)
    names = header_symbols(repo_root)
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), f"{name} declared in ta_b200.h but not exported"
    assert set(names) == set(_native.SYMBOLS), "binding table and header disagree"


def test_abi_version_and_error_string():
    lib = _native.load()
    assert lib.ta_abi_version() == _native.TA_ABI_VERSION
    assert isinstance(lib.ta_last_error(), (bytes, type(None)))


def test_struct_layouts_match_header(repo_root, tmp_path):
    """sizeof / offsetof of every ABI struct as gcc sees the header == the ctypes mirror in _native.py."""
    import subprocess

    structs = {"ta_plan_desc": _native.PlanDesc, "ta_batch": _native.Batch, "ta_frontend_out": _native.FrontendOut}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "ta_b200.h"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["return 0; }"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(repo_root, "include"), "-o", str(exe), str(src)], check=True)
    got = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = _native.load()
    desc = _native.PlanDesc(0, 44100, 2048, 512, 128, 12, 384, 0, 0.0, 0.0, 0.85, 0.4)
    handle = ctypes.c_void_p()
    rc = lib.ta_plan_create(ctypes.byref(desc), ctypes.byref(handle))
    assert rc == _native.TA_ERR_CUDA and not handle.value
    assert b"no CPU fallback" in lib.ta_last_error()
    from track_analyser_b200 import engine

    with pytest.raises(RuntimeError):
        engine.Plan(44100)


def test_invalid_plan_arguments_are_rejected():
    lib = _native.load()
    handle = ctypes.c_void_p()
    bad = _native.PlanDesc(0, 44100, 1000, 512, 128, 12, 384, 0, 0.0, 0.0, 0.85, 0.4)
    assert lib.ta_plan_create(ctypes.byref(bad), ctypes.byref(handle)) == _native.TA_ERR_INVALID
    assert b"n_fft" in lib.ta_last_error()
    bad = _native.PlanDesc(0, 44100, 2048, 0, 128, 12, 384, 0, 0.0, 0.0, 0.85, 0.4)
    assert lib.ta_plan_create(ctypes.byref(bad), ctypes.byref(handle)) == _native.TA_ERR_INVALID
    assert b"hop" in lib.ta_last_error()
    bad = _native.PlanDesc(0, 44100, 8192, 512, 128, 12, 384, 0, 0.0, 0.0, 0.85, 0.4)
    assert lib.ta_plan_create(ctypes.byref(bad), ctypes.byref(handle)) == _native.TA_ERR_INVALID
