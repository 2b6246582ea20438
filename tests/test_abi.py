"""The C-ABI shared library loads without a GPU and exports exactly the symbols include/fitclip_b200.h declares
(no compute calls here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fitclip_b200.h")


@pytest.fixture(scope="module")
def lib():
    from fitclip_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib.load()


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"FC_API\s+[\w\s\*]+?\b(fc_\w+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(lib):
    from fitclip_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes signatures out of sync with the header"


def test_no_undeclared_exports():
    from fitclip_b200 import _lib
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = sorted({line.split()[-1] for line in out.splitlines() if " T " in line and "fc_" in line.split()[-1]})
    assert exported == declared_symbols()


def test_version_and_error_buffer(lib):
    assert lib.fc_version() >= 100
    buf = ctypes.create_string_buffer(64)
    assert lib.fc_last_error(buf, 64) >= 0
    assert lib.fc_launch_count() >= 0
    assert lib.fc_sim_workspace_bytes(1000, 500, 512, 3) == 1000 * 1536 * 2 + 500 * 1536 * 2


def test_only_sm100a_code_and_tcgen05_tma_present():
    from fitclip_b200 import _lib
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True)
    if sass.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in sass.stdout and "sm_90" not in sass.stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM"):  # tcgen05.mma, TMA load/store, tcgen05.ld
        assert mnemonic in sass.stdout, mnemonic


def test_config_struct_layout_matches_header():
    from fitclip_b200 import _lib
    fields = re.search(r"typedef struct fc_config \{(.*?)\} fc_config;", open(HEADER).read(), re.S).group(1)
    names = re.findall(r"int32_t\s+(\w+);", fields)
    assert names == [n for n, _ in _lib.fc_config._fields_]
    assert ctypes.sizeof(_lib.fc_config) == 4 * len(names)
    fields = re.search(r"typedef struct fc_profile_record \{(.*?)\} fc_profile_record;", open(HEADER).read(), re.S).group(1)
    assert ctypes.sizeof(_lib.fc_profile_record) == 8 + 8 * 7
