"""fortran/m_b200_bindings.f90 is shipped uncompiled (no Fortran compiler in this image), so
its consistency with include/mfc_b200.h is checked textually: every bind(C) name must be a
symbol the library exports, and the bind(C) derived type must list the members of
mfc_b200_params_t in the same order as the ctypes mirror the tests use."""
import os
import re

from microfc_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "fortran", "m_b200_bindings.f90")).read()


def test_every_bound_name_is_exported():
    names = re.findall(r"bind\(C,\s*name='(\w+)'\)", SRC)
    lib_names = [n for n in names if n.startswith("mfc_b200_")]
    assert len(lib_names) >= 12
    for n in lib_names:
        assert n in abi.EXPORTED_SYMBOLS, n
    # the host-facing part of the ABI is bound completely
    host_api = [s for s in abi.EXPORTED_SYMBOLS if s not in (
        "mfc_b200_get_weno_coefficients", "mfc_b200_kernel_launches", "mfc_b200_state_snapshot",
        "mfc_b200_state_restore", "mfc_b200_timer_start", "mfc_b200_timer_stop", "mfc_b200_profile_enable",
        "mfc_b200_profile_get", "mfc_b200_kernel_name", "mfc_b200_debug_fill_ghosts")]
    for s in host_api:
        assert s in lib_names, s


def test_params_type_member_order_matches_ctypes_mirror():
    body = re.search(r"type, bind\(C\) :: mfc_b200_params_t(.*?)end type", SRC, re.S).group(1)
    members = []
    for line in body.splitlines():
        line = line.split("!")[0]
        if "::" not in line:
            continue
        for m in line.split("::")[1].split(","):
            m = m.strip()
            if m and not m[0].isdigit() and not m.startswith("MFC_"):
                members.append(re.sub(r"\(.*", "", m))
    members = [m for m in members if m and m != ")"]
    assert members == [f[0] for f in abi.Params._fields_], members


def _members(type_name):
    body = re.search(r"type, bind\(C\) :: %s(.*?)end type" % type_name, SRC, re.S).group(1)
    members = []
    for line in body.splitlines():
        line = line.split("!")[0]
        if "::" not in line:
            continue
        for m in re.sub(r"\([^)]*\)", "", line.split("::")[1]).split(","):
            if m.strip():
                members.append(m.strip())
    return members


def test_patch_type_member_order_matches_ctypes_mirror():
    assert _members("mfc_b200_patch_t") == [f[0] for f in abi.Patch._fields_]
