"""The C-ABI library loads without a GPU, exports every symbol include/tde_b200.h declares, and
fails loudly (never falls back) when no CUDA device is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "tde_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tde_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_are_exported(cuda_lib):
    from torchdriveenv_b200 import _capi
    declared = _declared_symbols()
    assert len(declared) >= 25
    raw = C.CDLL(_capi.library_path())
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in include/tde_b200.h but not exported"
    assert sorted(_capi.EXPORTS) == declared, "torchdriveenv_b200/_capi.py and the header disagree"


def test_version_and_default_config(cuda_lib):
    from torchdriveenv_b200 import _capi
    assert cuda_lib.tde_version() == 100
    c = _capi.TdeConfig()
    assert cuda_lib.tde_default_config(C.byref(c)) == 0
    d = _capi.default_config()
    for name, _ in _capi.TdeConfig._fields_:
        if name != "reserved":
            assert getattr(c, name) == getattr(d, name), name
    # the reference's EnvConfig defaults (gym_env.py:34-54)
    assert (c.max_environment_steps, c.waypoint_bonus, c.heading_penalty, c.distance_bonus, c.distance_cutoff) == (200, 100.0, 25.0, 1.0, 0.5)
    assert c.terminated_at_infraction == 1 and c.left_handed_coordinates == 1


def test_config_struct_layout_matches_header():
    from torchdriveenv_b200 import _capi
    # int32 x2, int64, int32 x6, float x11, int32 x8
    assert C.sizeof(_capi.TdeConfig) == 8 + 8 + 24 + 44 + 32 + 4  # + tail padding to 8
    assert _capi.TdeConfig.env_index_offset.offset == 8 and _capi.TdeConfig.dt.offset == 40


def test_create_rejects_bad_arguments_and_missing_gpu(cuda_lib):
    from torchdriveenv_b200 import _capi
    import torch
    h = C.c_void_p()
    cfg = _capi.default_config(num_envs=0)
    assert cuda_lib.tde_create(C.byref(cfg), C.byref(h)) == -1
    assert b"num_envs" in cuda_lib.tde_last_error(None)
    cfg = _capi.default_config(num_envs=4, max_agents=65)
    assert cuda_lib.tde_create(C.byref(cfg), C.byref(h)) == -3
    if not torch.cuda.is_available():
        cfg = _capi.default_config(num_envs=4, max_agents=4)
        rc = cuda_lib.tde_create(C.byref(cfg), C.byref(h))
        assert rc == -2 and b"no CPU fallback" in cuda_lib.tde_last_error(None)
    assert cuda_lib.tde_step(None, None, None, None, None, None, None, None) == -1


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "torchdriveenv_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("is test infrastructure", "").replace("The oracle under oracle/", "").replace("the oracle", "").replace("CPU oracle", "") \
                    or "import" not in "".join(l for l in text.splitlines() if "oracle" in l and not l.strip().startswith(("#", "//", "*", '"'))), f
                for line in text.splitlines():
                    s = line.strip()
                    if s.startswith(("import ", "from ", "#include")):
                        assert "oracle" not in s, f"{f}: {s}"
