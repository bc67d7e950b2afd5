"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/lnaphylodyn_b200.h declares; compute calls fail loudly without a GPU."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "lnaphylodyn_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lna_[A-Za-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    from lnaphylodyn_b200 import _lib
    L = _lib.lib()
    decl = declared_symbols()
    assert len(decl) >= 30
    for name in decl:
        assert hasattr(L, name), f"{name} declared in the header but not exported by the .so"
    assert set(decl) == set(_lib.exported_symbols()), set(decl) ^ set(_lib.exported_symbols())
    assert L.lna_abi_version() == 1


def test_shard_helper():
    import ctypes as C
    from lnaphylodyn_b200 import _lib
    L = _lib.lib()
    covered = []
    for r in range(8):
        b, e = C.c_int64(), C.c_int64()
        assert L.lna_shard(8192 + 5, r, 8, C.byref(b), C.byref(e)) == 0
        covered.append((b.value, e.value))
    assert covered[0][0] == 0 and covered[-1][1] == 8197
    assert all(covered[i][1] == covered[i + 1][0] for i in range(7))
    assert max(e - b for b, e in covered) - min(e - b for b, e in covered) <= 1


def test_no_cpu_fallback(changepoint):
    """Without a CUDA device every compute entry point must fail loudly (never silently run on CPU)."""
    from lnaphylodyn_b200 import _lib, api
    if _lib.lib().lna_device_count() > 0:
        pytest.skip("a GPU is present")
    g = changepoint
    with pytest.raises(_lib.LnaError, match="no CUDA device"):
        api.problem_for(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under lnaphylodyn_b200/ may reference it."""
    pkg = os.path.join(ROOT, "lnaphylodyn_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.lower() or f == "__init__.py" and False, f"{f} mentions the oracle"


HOT_PATH_EXPORTS = """betaTs param_transform ODE_rk45 ODE_rk45_stop SigmaF KF_param KF_param_chol log_like_traj_general2
log_like_traj_general_adjust Traj_sim_general3 Traj_sim_general_noncentral TransformTraj Traj_sim_ezG2 Traj_sim_ezG_NC
coal_loglik3 coal_loglik volz_loglik_nh2 volz_loglik_nh3 Ode_Coarse_Slicer New_Param_List Update_Param Param_Slice_update
Param_Slice_update_all2 ESlice_change_points ESlice_par_General ESlice_general_NC ESlice_general2""".split()


def test_reference_side_binding_covers_the_export_list():
    """integration/rcpp_shim.cpp defines every hot-path export of SURVEY.md section 8(b) (R/RcppExports.R names), and the
    drop-in build of it (the reference's harness over the binding instead of the reference's bodies) links with no
    unresolved symbol and exports the harness entry points the reference build exports."""
    import ctypes as C
    import subprocess
    from oracle import reference as ref
    src = open(os.path.join(ROOT, "integration", "rcpp_shim.cpp")).read()
    for name in HOT_PATH_EXPORTS:
        assert re.search(r"^[A-Za-z:<> ]+\b%s\(" % name, src, flags=re.M), f"{name} is not defined by the binding"
    if not ref.dropin_available():
        pytest.skip("oracle/_ref/liblna_dropin.so not built and /root/reference absent")
    so = os.path.join(ROOT, "oracle", "_ref", "liblna_dropin.so")
    C.CDLL(so)  # resolves liblnaphylodyn_b200.so through its rpath
    def syms(path):
        out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True).stdout
        return {l.split()[-1] for l in out.splitlines() if " T ref_" in l}
    mine = syms(so)
    assert len(mine) >= 30
    if ref.available():
        theirs = syms(os.path.join(ROOT, "oracle", "_ref", "liblna_ref.so"))
        assert theirs - mine <= {"ref_inv2", "ref_chols"}, theirs - mine  # basic_util.cpp stays the reference's own file


def test_header_is_plain_c_and_links(tmp_path):
    """The drop-in boundary is a C ABI: include/lnaphylodyn_b200.h must compile as C99 (no C++ in the signatures) and a
    plain C program must link against the library and call into it."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    src = tmp_path / "t.c"
    src.write_text('#include <stdio.h>\n#include "lnaphylodyn_b200.h"\n'
                   "int main(void) { lna_cfg c; lna_mcmc_opts o; lna_mcmc2_opts o2; long long b, e;\n"
                   "  (void)c; (void)o; (void)o2;\n"
                   "  if (lna_shard(10, 1, 3, (int64_t *)&b, (int64_t *)&e)) return 2;\n"
                   '  printf("%d %lld %lld\\n", lna_abi_version(), b, e); return 0; }\n')
    libdir = os.path.join(ROOT, "lnaphylodyn_b200")
    from lnaphylodyn_b200 import _lib
    _lib.lib()  # built
    exe = tmp_path / "t"
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                           "-L", libdir, "-llnaphylodyn_b200", "-Wl,-rpath," + libdir, "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.split() == ["1", "4", "7"], (out.stdout, out.stderr)
