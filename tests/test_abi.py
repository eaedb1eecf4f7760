"""The C-ABI boundary without a GPU: libmfc_b200.so loads, exports every entry point
include/mfc_b200.h declares, mirrors the struct layout, validates parameters with the
reference's messages, and FAILS LOUDLY (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from microfc_b200 import abi, cases, pre_process
from microfc_b200.domain import ghosted_metrics, rank_layout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mfc_b200.h")


def _declared():
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(mfc_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built_and_exports_every_declared_symbol():
    L = abi.lib()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/mfc_b200.h but not exported"
    assert sorted(abi.EXPORTED_SYMBOLS) == names


def test_only_the_abi_is_exported():
    out = subprocess.run(["nm", "-D", "--defined-only", abi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    syms = [l.split()[-1] for l in out.splitlines() if " T " in l]
    ours = [s for s in syms if not s.startswith(("_init", "_fini", "__cuda", "__cudart", "cuda", "__nv"))]
    assert all(s.startswith("mfc_b200_") for s in ours), [s for s in ours if not s.startswith("mfc_b200_")]


def test_library_does_not_link_the_oracle():
    out = subprocess.run(["ldd", abi.LIB_PATH], capture_output=True, text=True).stdout
    assert "liborc" not in out
    strings = subprocess.run(["strings", abi.LIB_PATH], capture_output=True, text=True).stdout
    assert "orc_step" not in strings and "liborc" not in strings


def test_params_struct_layout_matches_the_header():
    """sizeof / offsets computed by gcc from the header == the ctypes mirror."""
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "mfc_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(mfc_b200_params_t), offsetof(mfc_b200_params_t, weno_eps),
         offsetof(mfc_b200_params_t, bc), offsetof(mfc_b200_params_t, gammas), offsetof(mfc_b200_params_t, Re),
         offsetof(mfc_b200_params_t, cb), offsetof(mfc_b200_params_t, ds), offsetof(mfc_b200_params_t, strict_math));
  return 0; }'''
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        src, exe = os.path.join(td, "t.c"), os.path.join(td, "t")
        open(src, "w").write(prog)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        got = list(map(int, subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()))
    P = abi.Params
    want = [C.sizeof(P), P.weno_eps.offset, P.bc.offset, P.gammas.offset, P.Re.offset, P.cb.offset, P.ds.offset,
            P.strict_math.offset]
    assert got == want


def test_patch_struct_layout_matches_the_header():
    """mfc_b200_patch_t (device-side initial condition): header == ctypes mirror."""
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "mfc_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(mfc_b200_patch_t), offsetof(mfc_b200_patch_t, alter_patch),
         offsetof(mfc_b200_patch_t, x_centroid), offsetof(mfc_b200_patch_t, radii), offsetof(mfc_b200_patch_t, smooth_coeff),
         offsetof(mfc_b200_patch_t, pres), offsetof(mfc_b200_patch_t, alpha));
  return 0; }'''
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        src, exe = os.path.join(td, "t.c"), os.path.join(td, "t")
        open(src, "w").write(prog)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        got = list(map(int, subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()))
    P = abi.Patch
    want = [C.sizeof(P), P.alter_patch.offset, P.x_centroid.offset, P.radii.offset, P.smooth_coeff.offset, P.pres.offset,
            P.alpha.offset]
    assert got == want and abi.MAX_PATCHES == 10


def _params(cfg, cb):
    lay = rank_layout(0, 1, cfg)
    met = ghosted_metrics(lay, cfg, cb, [lay])
    p = abi.Params()
    p.abi_version = abi.ABI_VERSION
    p.m, p.n, p.p = lay.N
    p.m_glb, p.n_glb, p.p_glb = cfg.m, cfg.n, cfg.p
    p.num_dims, p.num_fluids, p.sys_size, p.buff_size = cfg.num_dims, cfg.num_fluids, cfg.sys_size, cfg.buff_size
    p.weno_order, p.weno_eps, p.time_stepper = cfg.weno_order, cfg.weno_eps, cfg.time_stepper
    p.t_step_start, p.t_step_stop = 0, 10
    for d in range(3):
        p.bc[2 * d], p.bc[2 * d + 1] = lay.bc[d]
    p.proc_rank, p.num_procs = 0, 1
    for i in range(cfg.num_fluids):
        p.gammas[i], p.pi_infs[i] = cfg.gamma[i], cfg.pi_inf[i]
        p.Re[i][0], p.Re[i][1] = cfg.Re[i][0], cfg.Re[i][1]
    keep = []
    for d in range(cfg.num_dims):
        for name, arrs in (("cb", met.cb), ("cc", met.cc), ("ds", met.ds)):
            a = np.ascontiguousarray(arrs[d])
            keep.append(a)
            getattr(p, name)[d] = a.ctypes.data_as(abi.c_double_p)
    p.device = -1
    return p, keep


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.parametrize("mutate,msg", [
    (lambda p: setattr(p, "abi_version", 99), "ABI version"),
    (lambda p: setattr(p, "num_fluids", 0), "num_fluids"),
    (lambda p: setattr(p, "sys_size", 5), "sys_size"),
    (lambda p: setattr(p, "m", 0), "value of m"),
    (lambda p: setattr(p, "weno_order", 4), "weno_order"),
    (lambda p: setattr(p, "weno_eps", 0.0), "weno_eps"),
    (lambda p: setattr(p, "time_stepper", 0), "time_stepper"),
    (lambda p: setattr(p, "buff_size", 3), "buff_size"),
    (lambda p: p.gammas.__setitem__(0, -1.0), "gamma"),
])
def test_init_rejects_bad_parameters_like_s_check_input_file(mutate, msg):
    cfg = cases.config(cases.sod_1d())
    p, keep = _params(cfg, pre_process.generate_grid(cfg))
    mutate(p)
    L = abi.lib()
    rc = L.mfc_b200_init(C.byref(p))
    assert rc == -1
    assert msg in L.mfc_b200_last_error().decode()


def test_more_fluids_than_the_kernels_are_built_for_fail_at_init():
    """MFC_B200_MAX_FLUIDS (4) is the array extent of the ABI; the sweep kernels are instantiated
    for 1..MFC_B200_BUILT_FLUIDS (3).  Four fluids must be refused by mfc_b200_init, not by the
    first step after a multi-gigabyte upload."""
    cfg = cases.config(cases.sod_1d())
    p, keep = _params(cfg, pre_process.generate_grid(cfg))
    p.num_fluids, p.sys_size = 4, 2 * 4 + 1 + 1
    for i in range(4):
        p.gammas[i], p.pi_infs[i] = 2.5, 0.0
    L = abi.lib()
    assert L.mfc_b200_init(C.byref(p)) == -7               # MFC_B200_EUNSUPPORTED
    assert b"instantiated" in L.mfc_b200_last_error()


def test_null_params_is_einval():
    assert abi.lib().mfc_b200_init(None) == -1


def test_call_order_is_enforced():
    L = abi.lib()
    L.mfc_b200_finalize()
    assert L.mfc_b200_upload(None) == -4                      # MFC_B200_ESTATE
    assert L.mfc_b200_step(0, 1e-3, None, None) == -4
    assert L.mfc_b200_download(None) == -4
    assert b"before" in L.mfc_b200_last_error()
    assert L.mfc_b200_generate_initial_condition(1, None, None, 1.0) == -4
    assert b"before mfc_b200_init" in L.mfc_b200_last_error()


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful on a host without a CUDA device")
def test_no_cuda_device_fails_loudly_no_cpu_fallback():
    cfg = cases.config(cases.sod_1d())
    p, keep = _params(cfg, pre_process.generate_grid(cfg))
    L = abi.lib()
    rc = L.mfc_b200_init(C.byref(p))
    assert rc == -2, "MFC_B200_ENODEVICE expected: the hot path has no CPU fallback"
    assert b"no CPU fallback" in L.mfc_b200_last_error()
    from microfc_b200.simulation import Simulation
    with pytest.raises(abi.MfcB200Error) as e:
        Simulation(cfg, pre_process.generate_grid(cfg))
    assert e.value.code == -2


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "microfc_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".inc")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle_lib" not in txt and "liborc" not in txt and "mfc_oracle" not in txt, os.path.join(dp, f)
