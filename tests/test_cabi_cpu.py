"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol the headers declare,
the ctypes prototype table matches the headers, and compute entry points fail loudly without a device."""
import ctypes as C
import os
import re

import pytest

import fibergen_b200 as fb
from fibergen_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    return sorted(set(re.findall(r"\b(fg(?:b|ls)_[A-Za-z0-9_]+)\s*\(", src)) - {"fgls_callback"})


def test_library_is_built_in_tree():
    assert os.path.exists(L.LIB_PATH), "run `make` or __graft_entry__.build()"


@pytest.mark.parametrize("header", ["fgb200.h", "fgb200_lssolver.h"])
def test_every_declared_symbol_is_exported(header):
    lib = C.CDLL(L.LIB_PATH)
    names = declared(header)
    assert len(names) > 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_prototype_table_covers_the_headers():
    names = set(declared("fgb200.h")) | set(declared("fgb200_lssolver.h"))
    assert names == set(L.PROTOTYPES), (names ^ set(L.PROTOTYPES))


def test_version_and_no_cpu_fallback():
    lib = L.load()
    assert b"sm_100a" in lib.fgb_version()
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    # without a device creation must fail loudly (FGB_ENODEV), never fall back to a CPU path
    with pytest.raises(fb.FgbError) as e:
        fb.Context(4, 4, 4)
    assert e.value.code == L.FGB_ENODEV
    s = fb.LSSolver(4, 4, 4)
    s.add_material("m", "iso", 1.0, 1.0)
    with pytest.raises(fb.FgbError):
        s.init()


def test_host_side_settings_errors_without_device():
    s = fb.LSSolver(4, 4, 4)
    s.set("tol", 1e-6)
    s.set("method", "basic")
    s.set("loadsteps", 4)
    with pytest.raises(fb.FgbError, match="Unknown solver setting"):
        s.set("bogus", 1)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fibergen_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read().replace("the CPU oracle", ""), f


@pytest.mark.parametrize("header", ["fgb200.h", "fgb200_lssolver.h"])
def test_headers_are_plain_c(header):
    """the boundary is a C ABI: both headers must compile as C99 on their own (cgo / JNI / ctypes-generators consume them)"""
    import subprocess
    p = subprocess.run(["gcc", "-std=c99", "-Wall", "-pedantic", "-fsyntax-only", "-x", "c", "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "include", header)], capture_output=True, text=True)
    assert p.returncode == 0 and not p.stderr.strip(), p.stderr


def test_c_host_links_and_fails_loudly_without_device(tmp_path):
    """a host written in plain C against include/fgb200.h links with libfgb200.so; without an sm_100 device fgb_create returns
    FGB_ENODEV with a message (no CPU path behind the ABI)"""
    import subprocess
    import torch
    src = tmp_path / "host.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "fgb200.h"
#include "fgb200_lssolver.h"
int main(void) {
    fgb_ctx* ctx = NULL;
    int rc;
    if (!strstr(fgb_version(), "sm_100a")) return 2;
    rc = fgb_create(&ctx, 8, 8, 8, 1.0, 1.0, 1.0, FGB_MODE_ELASTICITY, FGB_GAMMA_STAGGERED, -1, 0, 1);
    printf("rc=%d msg=%s\n", rc, fgb_last_error(ctx));
    if (rc == FGB_OK) { fgb_destroy(ctx); return 0; }
    return (rc == FGB_ENODEV && ctx == NULL && strlen(fgb_last_error(NULL)) > 0) ? 10 : 3;
}
''')
    exe = tmp_path / "host"
    libdir = os.path.dirname(L.LIB_PATH)
    p = subprocess.run(["gcc", "-std=c99", "-Wall", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L" + libdir,
                        "-lfgb200", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == (0 if torch.cuda.is_available() else 10), (r.returncode, r.stdout, r.stderr)
