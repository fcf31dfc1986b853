"""CPU: the C-ABI library loads and exports exactly what include/pldepth_b200.h declares
(no compute calls -- there is no GPU here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pldepth_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pld_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("pld_ctx_create", "pld_mask_compact", "pld_sample_lists_philox", "pld_sample_lists_fed",
                 "pld_sample_lists_mt", "pld_mt19937_generate", "pld_score_lists", "pld_select_top",
                 "pld_listmle_fwd_bwd", "pld_fused_sample_loss_bwd", "pld_fused_step", "pld_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol_and_binding_is_complete():
    from pldepth_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        pytest.fail("libpldepth_b200.so is not built; run __graft_entry__.build()")
    lib = _lib.load_library()
    syms = declared_symbols()
    for s in syms:
        assert hasattr(lib, s), "library does not export %s" % s
    assert sorted(_lib.SIGNATURES) == syms, "ctypes SIGNATURES and the header disagree"
    assert lib.pld_version() >= 100
    assert lib.pld_launch_count() == 0


def test_header_cites_reference_lines():
    src = open(HEADER).read()
    for cite in ("sampling.py:135", "sampling.py:110-145", "depth_utils.py:39-61", "nll_loss.py:43-62",
                 "sampling.py:161-169", "depth_utils.py:5-21"):
        assert cite in src


def test_missing_library_fails_loudly(tmp_path):
    from pldepth_b200 import _lib
    with pytest.raises(_lib.PLDError):
        _lib.load_library(str(tmp_path / "nope.so"))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pldepth_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_header_is_valid_c_and_links_from_plain_c(tmp_path):
    """The boundary is a C ABI: the header must compile as C99 and a plain-C program must link against the
    library and call it (no GPU needed for pld_version / pld_last_error)."""
    import shutil
    import subprocess
    from pldepth_b200 import _lib
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    src = tmp_path / "use_abi.c"
    src.write_text('#include "pldepth_b200.h"\n#include <stdio.h>\n'
                   'int main(void) { pld_ctx* c = 0; (void)c; printf("%d %s", pld_version(), pld_last_error()); '
                   'return pld_ctx_destroy(0); }\n')
    exe = tmp_path / "use_abi"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-l:libpldepth_b200.so", "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    assert out.split()[0] == "100"


def test_ctypes_signatures_have_the_header_arity():
    """Every prototype of the header and its ctypes binding take the same number of arguments (a drifted binding would
    pass garbage pointers to the kernels)."""
    from pldepth_b200 import _lib
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = re.findall(r"\b(pld_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S)
    assert len(protos) == len(_lib.SIGNATURES)
    for name, args in protos:
        args = " ".join(args.split())
        n = 0 if args in ("", "void") else args.count(",") + 1
        assert n == len(_lib.SIGNATURES[name][1]), "%s: header has %d arguments, ctypes binding %d" % (
            name, n, len(_lib.SIGNATURES[name][1]))
