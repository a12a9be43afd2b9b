"""The C-ABI library loads and exports every symbol include/mpcb200.h declares; no compute calls (no GPU here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mpcb200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mpcb200_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from mpc_b200 import _capi
    lib = _capi.load()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_capi.EXPORTS.keys()) == names     # the ctypes table covers the header, nothing more


def test_config_struct_matches_header_layout(tmp_path):
    """The ctypes mirror of mpcb200_config against the C compiler's view of include/mpcb200.h: size and every field offset."""
    import subprocess
    from mpc_b200 import _capi
    cfg = _capi.default_config(30, _capi.F32)
    assert cfg.abi_version == _capi.ABI_VERSION == 4 and cfg.N == 30
    names = [n for n, _ in _capi.Config._fields_]
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "mpcb200.h"\nint main(void){printf("%zu\\n", sizeof(mpcb200_config));' + \
           "".join('printf("%%zu\\n", offsetof(mpcb200_config, %s));' % n for n in names) + "return 0;}"
    src = tmp_path / "layout.c"
    src.write_text(prog)
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(tmp_path / "layout"), str(src)])
    out = [int(x) for x in subprocess.check_output([str(tmp_path / "layout")]).split()]
    assert out[0] == C.sizeof(_capi.Config)
    assert out[1:] == [getattr(_capi.Config, n).offset for n in names]
    assert abs(cfg.l_wb - 2.5789128) < 1e-12 and cfg.l_fric == 2.578
    assert (cfg.deltav_min, cfg.deltav_max, cfg.a_max, cfg.v_min, cfg.v_max) == (-0.4, 0.4, 11.5, 0.0, 50.8)
    c64 = _capi.default_config(50, _capi.F64)
    assert c64.mu_min < cfg.mu_min and c64.tol_step < cfg.tol_step
    # the host emulator's mirror (test tooling) must be the same struct
    import hostsim
    assert [(n, t) for n, t in hostsim.Config._fields_] == [(n, t) for n, t in _capi.Config._fields_]


def test_no_cpu_fallback_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mpc_b200 import _capi
    with pytest.raises(_capi.Mpcb200Error):
        _capi.Handle(_capi.default_config(30))
    from mpc_b200.optimizer import B200Optimizer, make_configuration, init_values_from_state
    import mpc_b200
    sc = mpc_b200.load_scenario("ZAM_Over-1_1_LF")
    with pytest.raises(_capi.Mpcb200Error):
        B200Optimizer(make_configuration(sc, 30), init_values_from_state(sc.x0), 30)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "motion-planning-for-autonomous-driving-with-mpc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "host_sim" not in txt.replace("tests/host_sim", ""), f
