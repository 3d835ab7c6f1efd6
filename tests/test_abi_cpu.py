"""The C-ABI shared library builds for sm_100a without a GPU, loads, and exports exactly the entry points
include/gyre_b200.h declares; the ctypes table in gyre_b200/_native.py covers every one of them.
No compute call is made here."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "gyre_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gyre_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from gyre_b200 import _native, build
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = header_functions()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in gyre_b200.h but not exported: {missing}"
    not_bound = [n for n in names if n not in _native.SIGNATURES]
    assert not not_bound, f"declared in gyre_b200.h but missing from the ctypes table: {not_bound}"
    extra = [n for n in _native.SIGNATURES if n not in names]
    assert not extra, f"bound in _native.py but not declared in the header: {extra}"
    assert _native.load(build_if_missing=False).gyre_b200_abi_version() == 2


def test_error_channel_without_gpu():
    """A failing call returns a negative status and leaves a message; nothing throws across the ABI."""
    from gyre_b200 import _native as N
    lib = N.load()
    n = ctypes.c_size_t()
    rc = lib.gyre_b200_tome_workspace_bytes(0, 0, 0, ctypes.byref(n))
    assert rc < 0
    assert "tome" in N.last_error()
    assert lib.gyre_b200_conv3x3_packed_elems(4, 320) == 320 * 9 * 64


def test_kernels_are_blackwell_native():
    """SASS evidence (B200_PROFILING.md): tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG/UTMASTG."""
    import shutil
    import subprocess
    from gyre_b200 import build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        import pytest
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", build.LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG"):
        assert mnemonic in sass, f"{mnemonic} not found in the SASS of libgyre_b200.so"
    assert "HMMA." not in sass.replace("UTCHMMA", ""), "legacy mma.sync path present"


def test_groupnorm_statistics_planning_is_host_logic():
    """gyre_b200_conv3x3_gn_parts / gyre_b200_groupnorm_pre_ok plan on the host (no launch): which convolutions of the
    headline configuration leave GroupNorm statistics, and what is refused."""
    from gyre_b200 import _native as N
    lib = N.load()
    G = 32
    # SD1.5, CFG batch 16: the 64x64 and 32x32 levels (one partial per 128-pixel tile of a sample), downsamplers included
    assert lib.gyre_b200_conv3x3_gn_parts(16, 64, 64, 320, 1, 1, G) == 32
    assert lib.gyre_b200_conv3x3_gn_parts(16, 32, 32, 640, 1, 1, G) == 8
    assert lib.gyre_b200_conv3x3_gn_parts(16, 64, 64, 320, 2, 1, G) == 8          # -> 32x32
    assert lib.gyre_b200_conv3x3_gn_parts(8, 64, 64, 320, 1, 1, G) == 32           # the CFG shared prefix runs at half the batch
    # VAE decoder widths at batch 8 (2048 partials per sample at 512x512: the fold launch)
    assert lib.gyre_b200_conv3x3_gn_parts(8, 512, 512, 128, 1, 1, G) == 2048
    assert lib.gyre_b200_conv3x3_gn_parts(8, 256, 256, 256, 1, 1, G) == 512
    assert lib.gyre_b200_conv3x3_gn_parts(8, 64, 64, 512, 1, 1, G) == 32
    # refused: a tile spans samples, tiles overhang the map, odd-width groups, too few channels
    assert lib.gyre_b200_conv3x3_gn_parts(16, 8, 8, 1280, 1, 1, G) == 0
    assert lib.gyre_b200_conv3x3_gn_parts(1, 12, 20, 320, 1, 1, G) == 0
    assert lib.gyre_b200_conv3x3_gn_parts(16, 64, 64, 96, 1, 1, G) == 0
    assert lib.gyre_b200_conv3x3_gn_parts(16, 64, 64, 4, 1, 1, G) == 0
    # the tunable switches it off everywhere
    old = N.get_tunable("GN_FUSE")
    N.set_tunable("GN_FUSE", 0)
    try:
        assert lib.gyre_b200_conv3x3_gn_parts(16, 64, 64, 320, 1, 1, G) == 0
    finally:
        N.set_tunable("GN_FUSE", old)
    # small maps are normalised by the single-pass kernel: no precomputed statistics there
    assert lib.gyre_b200_groupnorm_pre_ok(320, 4096, G) == 1
    assert lib.gyre_b200_groupnorm_pre_ok(1280, 256, G) == 0
    assert lib.gyre_b200_groupnorm_pre_ok(1280, 64, G) == 0
    assert lib.gyre_b200_groupnorm_pre_ok(100, 4096, G) == 0                       # C not divisible into 32 groups
