"""Case description front-end.

The reference keeps its inputs in ``case.py`` scripts that print one flat JSON dictionary
(``examples/*/case.py``); the toolchain filters the keys per target and writes Fortran
namelists (``toolchain/mfc/run/case_dicts.py:5-120``, ``toolchain/mfc/run/input.py:14-32``).
Those files stay unchanged: :func:`load_case_file` runs one and :func:`parse_case` turns the
flat dictionary into a :class:`CaseConfig` with the defaults of
``s_assign_default_values_to_user_inputs`` (``src/simulation/m_global_parameters.fpp:229-280``,
``src/pre_process/m_global_parameters.fpp:150-260``).

3-D keys (``p``, ``z_domain%*``, ``bc_z%*``, ``vel(3)``, ``z_centroid``, ``length_z``,
geometries 8 = sphere / 9 = cuboid) are an EXTENSION: the reference is 1-D/2-D only
(``src/simulation/m_global_parameters.fpp:50,403``).
"""
from __future__ import annotations

import json
import re
import subprocess
import sys
from dataclasses import dataclass, field
from typing import Dict, List, Optional

DFLT_REAL = -1.0e6   # m_global_parameters.fpp:37
DFLT_INT = -100      # m_global_parameters.fpp:38
MAX_FLUIDS = 4       # MFC_B200_MAX_FLUIDS in include/mfc_b200.h


def _logical(v) -> bool:
    if isinstance(v, str):
        return v.strip().strip("'\"").upper() in ("T", ".TRUE.", "TRUE")
    return bool(v)


@dataclass
class Patch:
    """patch_icpp(i), src/pre_process/m_global_parameters.fpp:196-230 (defaults)."""
    geometry: int = DFLT_INT
    x_centroid: float = DFLT_REAL
    y_centroid: float = DFLT_REAL
    z_centroid: float = DFLT_REAL
    length_x: float = DFLT_REAL
    length_y: float = DFLT_REAL
    length_z: float = DFLT_REAL
    radius: float = DFLT_REAL
    radii: List[float] = field(default_factory=lambda: [DFLT_REAL] * 3)
    normal: List[float] = field(default_factory=lambda: [DFLT_REAL] * 3)
    epsilon: float = DFLT_REAL
    alter_patch: Dict[int, bool] = field(default_factory=lambda: {0: True})
    smoothen: bool = False
    smooth_patch_id: int = 0          # default is the patch's own id, set in parse_case
    smooth_coeff: float = DFLT_REAL
    vel: List[float] = field(default_factory=lambda: [0.0] * 3)
    pres: float = DFLT_REAL
    alpha_rho: List[float] = field(default_factory=lambda: [0.0] * MAX_FLUIDS)
    alpha: List[float] = field(default_factory=lambda: [0.0] * MAX_FLUIDS)


@dataclass
class CaseConfig:
    # computational domain (global)
    m: int = DFLT_INT
    n: int = 0
    p: int = 0
    domain: List[List[float]] = field(default_factory=lambda: [[DFLT_REAL, DFLT_REAL] for _ in range(3)])
    stretch: List[bool] = field(default_factory=lambda: [False] * 3)
    a_s: List[float] = field(default_factory=lambda: [DFLT_REAL] * 3)     # a_x, a_y
    s_a: List[float] = field(default_factory=lambda: [DFLT_REAL] * 3)     # x_a, y_a
    s_b: List[float] = field(default_factory=lambda: [DFLT_REAL] * 3)     # x_b, y_b
    loops: List[int] = field(default_factory=lambda: [1] * 3)
    dt: float = DFLT_REAL
    t_step_start: int = DFLT_INT
    t_step_stop: int = DFLT_INT
    t_step_save: int = DFLT_INT
    # algorithm
    num_fluids: int = DFLT_INT
    time_stepper: int = DFLT_INT
    weno_order: int = DFLT_INT
    weno_eps: float = DFLT_REAL
    weno_Re_flux: bool = False
    run_time_info: bool = False
    parallel_io: bool = False
    bc: List[List[int]] = field(default_factory=lambda: [[DFLT_INT, DFLT_INT] for _ in range(3)])
    # fluids
    gamma: List[float] = field(default_factory=lambda: [DFLT_REAL] * MAX_FLUIDS)
    pi_inf: List[float] = field(default_factory=lambda: [DFLT_REAL] * MAX_FLUIDS)
    Re: List[List[float]] = field(default_factory=lambda: [[DFLT_REAL, DFLT_REAL] for _ in range(MAX_FLUIDS)])
    # initial condition
    num_patches: int = 0
    patches: List[Patch] = field(default_factory=list)

    # ---- derived, as in s_initialize_global_parameters_module (m_global_parameters.fpp:285-396)
    @property
    def num_dims(self) -> int:
        return 1 + min(1, self.n) + (min(1, self.p) if self.n > 0 else 0)

    @property
    def sys_size(self) -> int:
        return 2 * self.num_fluids + self.num_dims + 1

    @property
    def weno_polyn(self) -> int:
        return (self.weno_order - 1) // 2

    @property
    def viscous(self) -> bool:
        return any(self.Re[i][j] > 0 for i in range(self.num_fluids) for j in range(2))

    @property
    def buff_size(self) -> int:
        return 2 * self.weno_polyn + 2 if self.viscous else self.weno_polyn + 2

    @property
    def shape_glb(self):
        """(Nz, Ny, Nx) numbers of cells."""
        return (self.p + 1, self.n + 1, self.m + 1)

    def check(self) -> None:
        """The subset of s_check_input_file (src/simulation/m_start_up.fpp:117-290) that
        guards the hot path; raises ValueError with the reference's message."""
        def bad(what):
            raise ValueError(f"Unsupported value of {what}. Exiting ...")
        if self.m <= 0: bad("m")
        if self.n < 0: bad("n")
        if self.p < 0 or (self.p > 0 and self.n == 0): bad("p")
        if self.dt <= 0: bad("dt")
        if self.t_step_start < 0: bad("t_step_start")
        if self.t_step_stop <= self.t_step_start: bad("t_step_start and t_step_stop")
        if not (1 <= self.num_fluids <= MAX_FLUIDS): bad("num_fluids")
        if self.weno_order not in (1, 3, 5): bad("weno_order")
        if self.time_stepper not in (1, 2, 3): bad("time_stepper")
        if self.m + 1 < 5 * self.weno_order: bad("m and weno_order")
        if self.n > 0 and self.n + 1 < 5 * self.weno_order: bad("n and weno_order")
        if self.p > 0 and self.p + 1 < 5 * self.weno_order: bad("p and weno_order")
        if self.weno_eps <= 0.0 or self.weno_eps > 1e-6: bad("weno_eps")
        for d in range(self.num_dims):
            for s in range(2):
                if self.bc[d][s] < -12 or self.bc[d][s] > -1: bad(f"bc_{'xyz'[d]}")
            if (self.bc[d][0] == -1) != (self.bc[d][1] == -1): bad(f"bc_{'xyz'[d]}%beg and %end")
        for i in range(self.num_fluids):
            if self.gamma[i] <= 0.0: bad(f"fluid_pp({i + 1})%gamma")
            if self.pi_inf[i] < 0.0: bad(f"fluid_pp({i + 1})%pi_inf")


_IDX = re.compile(r"^(\w+)\((\d+)\)%(\w+)(?:\((\d+)\))?$")
_DIR = {"x": 0, "y": 1, "z": 2}


def parse_case(d: dict) -> CaseConfig:
    """Flat MFC case dictionary -> CaseConfig."""
    c = CaseConfig()
    npatch = int(d.get("num_patches", 0))
    c.num_patches = npatch
    c.patches = [Patch(smooth_patch_id=i + 1) for i in range(npatch)]
    for key, val in d.items():
        key = key.strip()
        if key in ("m", "n", "p", "t_step_start", "t_step_stop", "t_step_save", "num_fluids",
                   "time_stepper", "weno_order"):
            setattr(c, key, int(val))
        elif key in ("dt", "weno_eps"):
            setattr(c, key, float(val))
        elif key in ("run_time_info", "weno_Re_flux", "parallel_io"):
            setattr(c, key, _logical(val))
        elif re.match(r"^[xyz]_domain%(beg|end)$", key):
            c.domain[_DIR[key[0]]][0 if key.endswith("beg") else 1] = float(val)
        elif re.match(r"^bc_[xyz]%(beg|end)$", key):
            c.bc[_DIR[key[3]]][0 if key.endswith("beg") else 1] = int(val)
        elif re.match(r"^stretch_[xyz]$", key):
            c.stretch[_DIR[key[-1]]] = _logical(val)
        elif re.match(r"^a_[xyz]$", key):
            c.a_s[_DIR[key[-1]]] = float(val)
        elif re.match(r"^[xyz]_a$", key):
            c.s_a[_DIR[key[0]]] = float(val)
        elif re.match(r"^[xyz]_b$", key):
            c.s_b[_DIR[key[0]]] = float(val)
        elif re.match(r"^loops_[xyz]$", key):
            c.loops[_DIR[key[-1]]] = int(val)
        else:
            mt = _IDX.match(key)
            if not mt:
                continue      # post_process / formatting keys: not on this path
            grp, i, attr, j = mt.group(1), int(mt.group(2)), mt.group(3), mt.group(4)
            if grp == "fluid_pp" and i <= MAX_FLUIDS:
                if attr == "gamma": c.gamma[i - 1] = float(val)
                elif attr == "pi_inf": c.pi_inf[i - 1] = float(val)
                elif attr == "Re": c.Re[i - 1][int(j) - 1] = float(val)
            elif grp == "patch_icpp" and i <= npatch:
                pt = c.patches[i - 1]
                if attr in ("vel", "alpha", "alpha_rho", "radii", "normal"):
                    getattr(pt, attr)[int(j) - 1] = float(val)
                elif attr == "alter_patch":
                    pt.alter_patch[int(j)] = _logical(val)
                elif attr == "smoothen":
                    pt.smoothen = _logical(val)
                elif attr in ("geometry", "smooth_patch_id"):
                    setattr(pt, attr, int(val))
                elif hasattr(pt, attr):
                    setattr(pt, attr, float(val))
    return c


def load_case_file(path: str) -> CaseConfig:
    """Run an unchanged reference case script and parse the JSON it prints
    (toolchain/mfc/run/input.py:100-112)."""
    out = subprocess.run([sys.executable, path], check=True, capture_output=True, text=True).stdout
    return parse_case(json.loads(out))
