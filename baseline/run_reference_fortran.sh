#!/usr/bin/env bash
# Time the UNMODIFIED Fortran+MPI reference (CPU build) on one case, the way BASELINE.md 4.2
# describes -- for hosts that have the toolchain.  This image has none (no gfortran/nvfortran,
# no fypp, no MPI, no network), so here the script prints why and exits 0; bench.py's
# `--impl reference` arm then times the C++ restatement under oracle/ instead
# (cpu_baseline.kind = "port").
#
#   baseline/run_reference_fortran.sh <reference-tree> <case.py> [ranks] [--fixtures]
#
# --fixtures: also turn the run's restart files into tests/golden/ref_<case>.npz
# (baseline/make_reference_fixtures.py), which pins the oracle and the CUDA path against the
# reference itself (tests/test_reference_fixtures.py).
#
# Output: one line "ranks seconds_per_step grind_ns_per_cell_eq_rhs" from time_data.dat
# (mean cpu_time per RK step over steps >= 4, max over ranks: m_time_steppers.fpp:352-358,
# p_main.fpp:261-270).
set -euo pipefail
REF=${1:-/root/reference}
CASE=${2:-$REF/examples/2D_advection/case.py}
RANKS=${3:-$(nproc)}
FIXTURES=${4:-}
HERE=$(cd "$(dirname "$0")" && pwd)
missing=()
command -v gfortran >/dev/null 2>&1 || command -v nvfortran >/dev/null 2>&1 || missing+=("Fortran compiler (gfortran >= 5 or nvfortran)")
command -v mpirun >/dev/null 2>&1 || missing+=("MPI (mpirun + mpif90)")
command -v cmake >/dev/null 2>&1 || missing+=("cmake >= 3.18")
python3 -c 'import fypp' >/dev/null 2>&1 || command -v fypp >/dev/null 2>&1 || missing+=("fypp")
[ -d "$REF" ] || missing+=("reference tree $REF")
if [ ${#missing[@]} -gt 0 ]; then
    echo "SKIP: cannot build the Fortran reference here: missing ${missing[*]}" >&2
    exit 0
fi
WORK=$(mktemp -d)
cp -r "$REF" "$WORK/ref"                      # the reference tree is read-only; build in a copy
cd "$WORK/ref"
./mfc.sh build -t pre_process simulation -j "$(nproc)"
./mfc.sh run "$CASE" -n "$RANKS" -t pre_process simulation
TD=$(dirname "$CASE")/time_data.dat
[ -f "$TD" ] || TD=$(find . -name time_data.dat | head -1)
python3 - "$TD" "$CASE" <<'PY'
import json, subprocess, sys
ranks, secs = open(sys.argv[1]).read().split()[-2:]
case = json.loads(subprocess.run([sys.executable, sys.argv[2]], capture_output=True, text=True, check=True).stdout)
cells = (case["m"] + 1) * (case.get("n", 0) + 1) * (case.get("p", 0) + 1)
nd = 1 + (case.get("n", 0) > 0) + (case.get("p", 0) > 0)
E = 2 * case["num_fluids"] + nd + 1
print(ranks, secs, float(secs) / (cells * E * 3) * 1e9)
PY
if [ "$FIXTURES" = "--fixtures" ]; then
    python3 "$HERE/make_reference_fixtures.py" "$CASE" --case-dir "$(dirname "$TD")"
fi
