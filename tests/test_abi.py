"""CPU-side checks of the drop-in boundary: the library builds for sm_100a, loads, exports every symbol that
include/wgbs_b200.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "wgbs_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wgbs_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_something():
    syms = declared_symbols()
    assert "wgbs_create" in syms and "wgbs_pat2beta" in syms and len(syms) >= 15


def test_library_exports_every_declared_symbol(built_lib):
    L = ctypes.CDLL(built_lib)
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert not missing, f"declared in include/wgbs_b200.h but not exported: {missing}"


def test_python_binding_covers_header(built_lib):
    from wgbs_tools_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_built_for_sm100a_only(built_lib):
    out = subprocess.run(["cuobjdump", "-lelf", built_lib], stdout=subprocess.PIPE, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback(built_lib):
    """Without a GPU the product must fail loudly, never silently compute on the host."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    from wgbs_tools_b200.api import Context
    from wgbs_tools_b200._lib import WgbsError
    with pytest.raises(WgbsError, match="no usable CUDA device|no CPU path"):
        Context(0)


def test_product_never_touches_oracle():
    """wgbs_tools_b200/ must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "wgbs_tools_b200")
    bad = re.compile(r"(^|\W)(import|from)\s+oracle\b|oracle/_ref|liboracle|oracle\.harness|oracle_port")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert not bad.search(txt), f"{f} reaches into oracle/"


def test_bindings_and_library_agree_on_the_abi_version():
    """a stale libwgbs_b200.so (older struct layouts) must not load silently"""
    import re
    from wgbs_tools_b200 import _lib
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "wgbs_b200.h")).read()
    v = int(re.search(r"#define\s+WGBS_B200_ABI_VERSION\s+(\d+)", hdr).group(1))
    assert v == _lib.ABI_VERSION == _lib.lib.wgbs_abi_version()


def test_ctypes_structs_have_the_layout_of_the_header(tmp_path):
    """ViewOpts / PileupOpts (wgbs_tools_b200/_lib.py) field by field against offsetof() in include/wgbs_b200.h (gcc)"""
    import ctypes as C
    from wgbs_tools_b200._lib import PileupOpts, ViewOpts
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    prog = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{root}/include/wgbs_b200.h"', 'int main(void) {']
    for cname, T in (("wgbs_view_opts", ViewOpts), ("wgbs_pileup_opts", PileupOpts)):
        prog.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for f, _ in T._fields_:
            prog.append(f'printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
    prog += ['return 0; }']
    src = tmp_path / "layout.c"; src.write_text("\n".join(prog))
    exe = str(tmp_path / "layout")
    r = subprocess.run(["gcc", str(src), "-o", exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    got = dict(l.split() for l in subprocess.run([exe], stdout=subprocess.PIPE, text=True).stdout.splitlines())
    for cname, T in (("wgbs_view_opts", ViewOpts), ("wgbs_pileup_opts", PileupOpts)):
        assert int(got[cname]) == C.sizeof(T), cname
        for f, _ in T._fields_:
            assert int(got[f"{cname}.{f}"]) == getattr(T, f).offset, (cname, f)
