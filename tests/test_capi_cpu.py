"""The C-ABI library loads on a CPU-only box, exports every symbol the header declares, and the
compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from sihmpc_b200 import _capi, scenarios as sc
from sihmpc_b200.ihmpccontroller import IHMPCController

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "sihmpc_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sihmpc_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = _capi.load()
    names = _declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_capi.EXPORTS) == names
    assert lib.sihmpc_abi_version() == _capi.ABI_VERSION


def test_struct_layouts_match_header_field_order():
    txt = open(os.path.join(ROOT, "include", "sihmpc_b200.h")).read()
    body = txt[txt.index("typedef struct sihmpc_io {"):txt.index("} sihmpc_io;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"[\*\s]([A-Za-z_][A-Za-z0-9_]*)\s*[;,]", body)
    assert fields == [f[0] for f in _capi.IO._fields_]
    body = txt[txt.index("typedef struct sihmpc_plan_desc {"):txt.index("} sihmpc_plan_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"[\*\s]([A-Za-z_][A-Za-z0-9_]*)\s*[;,]", body)
    assert fields == [f[0] for f in _capi.PlanDesc._fields_]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _capi.load()
    assert lib.sihmpc_device_count() == 0
    c = sc.declare_distillation(IHMPCController)
    desc, keep = _capi.make_desc(c.compile())
    h = C.c_void_p()
    rc = lib.sihmpc_plan_create(C.byref(desc), C.byref(h))
    assert rc == -2 and b"no CPU fallback" in lib.sihmpc_last_error()
    with pytest.raises(_capi.SihmpcError):
        c.engine(4)
    with pytest.raises(_capi.SihmpcError):
        c.mpc(x0=np.zeros(22), ySP=[0, 0], w0=[], u0=[0, 0], pesos=[])


def test_product_never_imports_oracle():
    bad = []
    for base in ("sihmpc_b200", "sihmpc", "opom"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    t = open(os.path.join(dp, f)).read()
                    if re.search(r"^\s*(from|import)\s+oracle\b", t, flags=re.M) or "oracle/" in t:
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
