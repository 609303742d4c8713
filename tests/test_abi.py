"""The C-ABI shared library builds, loads and exports every symbol include/mips_b200.h declares.
No compute is attempted here (no GPU in the CPU tier); compute entry points must FAIL LOUDLY
without a device — there is no CPU fallback to route through."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from retrieval_augmented_mds_b200 import _lib, build

ROOT = Path(__file__).resolve().parent.parent


def test_library_builds_and_exports_declared_symbols():
    path = build.build()
    assert path.exists() and path.suffix == ".so" and ROOT in path.parents  # in-tree, not site-packages
    L = _lib.lib()
    names = _lib.declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in mips_b200.h but not exported: {missing}"


def test_header_is_plain_c_and_cites_reference():
    text = (ROOT / "include" / "mips_b200.h").read_text()
    assert 'extern "C"' in text and "torch" not in text.replace("no torch types", "")
    assert len(re.findall(r"mips\.py:\d+", text)) >= 8  # every entry point cites what it replaces


def test_sass_contains_blackwell_tensor_path():
    """tcgen05.mma / tcgen05.ld / TMA must be in the shipped binary (B200_PROFILING.md table)."""
    import shutil, subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", str(build.LIB_PATH)], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "STTM", "UTMALDG"):
        assert mnemonic in sass, mnemonic
    assert "sm_100a" in sass


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_entry_points_fail_loudly_without_gpu():
    L = _lib.lib()
    h = C.c_void_p()
    rc = L.mips_create(C.byref(h), 64, 0, 1, 0, 0)
    assert rc != 0 and not h.value
    assert len(L.mips_last_error()) > 0
    import retrieval_augmented_mds_b200 as m
    with pytest.raises(RuntimeError):
        m.B200FlatIndex(64)
    with pytest.raises((RuntimeError, ValueError)):
        m.normalize_L2(np.ones((2, 4), dtype=np.float32))


def test_argument_validation_without_device():
    L = _lib.lib()
    h = C.c_void_p()
    assert L.mips_create(C.byref(h), 0, 0, 1, 0, 0) == -1  # d <= 0
    assert b"d must be" in L.mips_last_error()
    assert L.mips_create(C.byref(h), 64, 7, 1, 0, 0) == -1  # bad metric
    assert L.mips_create(C.byref(h), 64, 0, 9, 0, 0) == -1  # bad dtype
    assert L.mips_ntotal(None) == -1
    assert L.mips_merge(None, None, None, 1, 4, 0, 4, 0, 0, 0.0, None, None, None, None, None, None,
                        1.0, 0.0, None, 0, None) == -1
    with pytest.raises(ValueError):
        _lib.check(-1)
    with pytest.raises(_lib.MipsError):
        _lib.check(-2)


def test_c_host_example_compiles_and_links_against_the_abi(tmp_path):
    """include/mips_b200.h is plain C99 and examples/c_abi_demo.c links against the in-tree library with
    gcc alone. Without a GPU the program must fail loudly (mips_create reports the missing device)."""
    import shutil, subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    build.build()
    exe = tmp_path / "c_abi_demo"
    cmd = [gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", str(ROOT / "include"),
           str(ROOT / "examples" / "c_abi_demo.c"), "-L", str(build.PKG_DIR), "-lmips_b200",
           f"-Wl,-rpath,{build.PKG_DIR}", "-o", str(exe)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    if not torch.cuda.is_available():
        run = subprocess.run([str(exe)], capture_output=True, text=True)
        assert run.returncode == 1 and "mips_b200:" in run.stderr    # no device: loud failure, no CPU path
