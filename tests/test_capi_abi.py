"""CPU: the C-ABI library loads, exports every symbol include/jclip_b200.h declares, the ctypes binding
covers exactly those symbols, and -- there being no GPU here -- the product fails loudly instead of
falling back to anything."""
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "jclip_b200.h")


def _header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(jcb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(jb):
    lib_path = str(jb._capi.LIB_PATH)
    assert os.path.exists(lib_path), "run __graft_entry__.build() first"
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (jcb_[a-z0-9_]+)", out))
    declared = _header_functions()
    assert len(declared) >= 30
    missing = [f for f in declared if f not in exported]
    assert not missing, f"declared in the header but not exported: {missing}"
    extra = sorted(exported - set(declared))
    assert not extra, f"exported but not declared in include/jclip_b200.h: {extra}"


def test_binding_matches_header(jb):
    assert sorted(jb._capi.PROTOTYPES) == _header_functions()
    lib = jb.load_library()
    assert lib.jcb_abi_version() == jb._capi.JCB_ABI_VERSION == 4


def test_header_is_plain_c(tmp_path):
    """extern "C", plain pointers and sizes, no C++ / torch types: the header compiles as C11."""
    c = tmp_path / "t.c"
    c.write_text('#include "jclip_b200.h"\nint main(void){ jcb_mta_params p; (void)p; return JCB_ABI_VERSION - 4; }\n')
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(c)],
                   check=True)


def test_sass_is_blackwell_native(jb):
    """tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, TMA -> UTMALDG must be in the shipped cubin."""
    out = subprocess.run(["cuobjdump", "-sass", str(jb._capi.LIB_PATH)], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnemonic in out.stdout, mnemonic
    assert "sm_100a" in out.stdout


@pytest.mark.skipif(torch.cuda.is_available(), reason="this check is for GPU-less hosts")
def test_no_cpu_fallback(jb):
    from ctypes import byref, c_void_p
    lib = jb.load_library()
    h = c_void_p()
    assert lib.jcb_ctx_create(0, byref(h)) == jb._capi.JCB_E_NO_DEVICE and not h.value
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        jb.get_context()
    sd = jb.synth.make_vit_state_dict(seed=0, layers=1)
    model = jb.jclip.build_model(sd)
    with pytest.raises(RuntimeError):
        model.encode_image(torch.zeros(1, 3, 224, 224))
    with pytest.raises(RuntimeError):
        jb.solve_mta(torch.zeros(3, 512), torch.zeros(512, 403))


def test_missing_library_fails_loudly(jb, tmp_path):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        jb._capi.load_library(tmp_path / "libjclip_b200.so")


def test_product_does_not_import_oracle():
    """Only tests/, smoke() and bench.py may touch oracle/."""
    pkg = os.path.join(ROOT, "jittor-clip-fewshot_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, f)
