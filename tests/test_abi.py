"""CPU: the C-ABI library builds for sm_100a, loads without a GPU and exports exactly what
include/hippo_b200.h declares.  No compute call is made here."""
import ctypes
import os
import re
import subprocess

from hippomm_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "hippo_b200.h")


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)          # drop comments
    names = re.findall(r"\b(hippo_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_header_and_binding_table_agree():
    declared = _declared_functions()
    assert len(declared) >= 17
    assert declared == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol(lib):
    for name in _declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/hippo_b200.h but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (hippo_\w+)", out))
    assert exported == set(_declared_functions())


def test_library_identifies_itself(lib):
    assert lib.hippo_abi_version() == _lib.ABI_VERSION
    assert isinstance(lib.hippo_last_error(), bytes)
    # pure size queries work without a device
    # two bf16 images of the rows plus the bit matrix of ONE band -- far below the n^2/8 bytes of a full bit matrix
    assert 2 * 100_000 * 1024 * 2 < lib.hippo_consolidate_workspace_bytes(100_000, 1024) < 100_000 * 100_000 // 8
    assert lib.hippo_topk_batched_workspace_bytes(10_000_000, 1024, 4096, 10) > 4096 * 1024 * 2
    assert lib.hippo_frame_pairs_workspace_bytes(3600, 224, 224, 3599) >= 3600 * 224 * 224
    assert lib.hippo_topk_single_workspace_bytes(10_000_000, 1024, 10) >= 148 * 2 * 10 * 8


def test_stream_desc_layout_matches_header():
    # hippo_stream_desc: 3 pointers/int64 ... 96 bytes, naturally aligned
    assert ctypes.sizeof(_lib.StreamDesc) == 96
    assert _lib.StreamDesc.sample_rate.offset == 64
    assert _lib.StreamDesc.out_bounds.offset == 72
    assert _lib.StreamDesc.max_segments.offset == 88


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    """The similarity kernel must be a real tcgen05 / TMA kernel (UTCHMMA, UTMALDG, LDTM in SASS)."""
    r = subprocess.run(["cuobjdump", "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True)
    if r.returncode != 0:
        import pytest

        pytest.skip("cuobjdump unavailable")
    sass = r.stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR"):
        assert mnemonic in sass, f"{mnemonic} missing from SASS"
    assert "HMMA.16816" not in sass, "legacy mma.sync path present"


def test_sass_of_the_streaming_kernels_uses_the_intended_instructions():
    """The secondary kernels, by SASS: packed fp32 pairs and byte dot products in the SSIM kernel, a warp-level
    REDUX + one 32-bit shared atomic (no 64-bit compare-and-swap loop) in the boundary kernel, 128-bit non-allocating
    loads in the GEMV kernels."""
    import pytest

    def sass_of(pattern):
        r = subprocess.run(["cuobjdump", "-sass", "-fun", pattern, str(_lib.LIB_PATH)], capture_output=True, text=True)
        if r.returncode != 0 or "Function" not in r.stdout:
            pytest.skip("cuobjdump unavailable or function not found")
        return r.stdout

    names = subprocess.run(["cuobjdump", "-elf", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    import re

    def mangled(fragment):
        m = re.search(r"\.text\.(_Z\w*%s\w*)" % fragment, names)
        if not m:
            pytest.skip(f"{fragment} not found in the ELF listing")
        return m.group(1)

    for kernel in ("ssim_pair_kernel", "ssim_pair7_kernel"):
        ssim = sass_of(mangled(kernel))
        # x^2 + y^2 and 2xy of a column are one 16 x 8-bit dot product each (IDP.2A), the squared error a byte dot
        # product (IDP.4A), the ratio runs on packed fp32 pairs
        for mnemonic in ("FMUL2", "FADD2", "IDP.4A", "IDP.2A.LO", "IDP.2A.HI", "SHFL.DOWN", "MUFU.RCP"):
            assert mnemonic in ssim, f"{mnemonic} missing from {kernel}"
        assert "LDL" not in ssim and "STL" not in ssim, f"{kernel} spills"
    seg = sass_of(mangled("segment_kernel"))
    assert "REDUX" in seg and "ATOMS.MAX" in seg, "boundary kernel: warp reduce + native shared atomic expected"
    assert "ATOMS.CAST.SPIN.64" not in seg, "boundary kernel: 64-bit compare-and-swap loop is back"
    few = sass_of(mangled("topk_few_kernel"))
    assert "LDG.E.NA.128" in few or "LDG.E.128" in few, "two-query GEMV: 128-bit loads expected"
