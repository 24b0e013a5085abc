"""CPU: the C-ABI shared library builds for sm_100a, loads, exports every symbol include/cw_b200.h declares,
and rejects bad arguments before touching the device.  No compute is launched here."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cw_b200.h")


@pytest.fixture(scope="module")
def lib():
    from gym_craftingworld_b200 import _lib
    return _lib.load()


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cw_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree():
    from gym_craftingworld_b200 import _lib
    assert declared_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), f"libcw_b200.so does not export {name}"
    assert lib.cw_abi_version() == 4
    assert lib.cw_error_string(0) == b"ok"
    assert b"NULL" in lib.cw_error_string(-2)


def test_library_is_native_sm100a_with_tma_bulk_store():
    """The shipped cubin targets sm_100a and the render path really uses the TMA bulk-copy unit (SASS UBLKCP)."""
    from gym_craftingworld_b200 import build
    so = build.build()
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass
    for kernel in ("cw_env_kernel", "cw_step_kernel", "cw_onehot_kernel"):
        assert kernel in sass


def test_struct_layout_matches_header(lib):
    from gym_craftingworld_b200 import _lib
    assert C.sizeof(_lib.CwConfig) == 8 * 4 + 16
    assert C.sizeof(_lib.CwState) == 6 * 8 + 3 * 8 + 3 * 8 + 3 * 8 + 2 * 8


def test_argument_errors_are_codes_not_crashes(lib):
    from gym_craftingworld_b200 import _lib
    from gym_craftingworld_b200.env import make_config
    cfg = make_config()
    st = _lib.CwState()
    st.n = 4                                    # pointers all NULL
    assert lib.cw_step(None, C.byref(st), None, None, None, None, 0, None) == -2          # CW_E_NULLPTR
    assert lib.cw_step(C.byref(cfg), C.byref(st), None, None, None, None, 0, None) == -2
    bad = make_config()
    bad.cell_stride = 441
    assert lib.cw_render(C.byref(bad), None, None, None, 1, None) == -1                    # CW_E_BADCONFIG
    bad = make_config()
    bad.number_of_tasks = 12
    assert lib.cw_reset(C.byref(bad), C.byref(st), None, None, None, None, None) == -1
    st.n = 0
    assert lib.cw_step(C.byref(cfg), C.byref(st), None, None, None, None, 7, None) == -3   # CW_E_BADFLAGS
    assert lib.cw_step(C.byref(cfg), C.byref(st), None, None, None, None, 1, None) == 0    # empty batch: no-op
    assert lib.cw_host_step(None, None, None, None, None) == -4                            # CW_E_BADHANDLE
    # the newer step entry points validate the same way
    st.n = 4
    assert lib.cw_step_render_chained(C.byref(cfg), C.byref(st), None, None, None, None, None, None, None, 0, None, 0, 1, None) == -2
    assert lib.cw_step_render_edit(C.byref(cfg), C.byref(st), None, None, None, None, None, None, None, 0, None, None) == -2
    assert lib.cw_step_chained(C.byref(cfg), C.byref(st), None, None, None, None, 0, None, 0, None) == -2
    st.n = 0
    assert lib.cw_step_chained(C.byref(cfg), C.byref(st), None, None, None, None, 0, None, 1024, None) == -1     # position out of range
    assert lib.cw_step_chained(C.byref(cfg), C.byref(st), None, None, None, None, 6, None, 0, None) == -3
    assert lib.cw_step_chained(C.byref(cfg), C.byref(st), None, None, None, None, 1, None, 5, None) == 0         # empty batch: no-op
    assert lib.cw_step_render_chained(C.byref(cfg), C.byref(st), None, None, None, None, None, None, None, 0, None, 1024, 1, None) == -1
    assert lib.cw_step_render_chained(C.byref(cfg), C.byref(st), None, None, None, None, None, None, None, 0, None, 0, 0, None) == -1
    assert lib.cw_step_render_edit(C.byref(cfg), C.byref(st), None, None, None, None, None, None, None, 2, None, None) == -3
    assert lib.cw_step_delta(C.byref(cfg), C.byref(st), None, None, None, None, 0, 64, None) == 0   # empty batch wins over the tag check
    tiny = make_config()
    tiny.H = tiny.W = 1; tiny.cell_stride = 16
    assert lib.cw_render(C.byref(tiny), None, None, None, 1, None) == -1                   # sides < 2 are rejected
    with pytest.raises(_lib.CwError):
        _lib.check(-1, "unit test")
