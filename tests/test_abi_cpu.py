"""CPU tests of the drop-in boundary: the C-ABI library loads (no GPU needed) and exports every symbol that
include/claymore_b200.h declares; host-side logic (config, samplers, scene schema)."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "claymore_b200.h")).read()
    return sorted(set(re.findall(r"CB200_API[^;(]*?\b(cb200_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import claymore_b200 as cb
    cb.build_library()
    lib = C.CDLL(cb.lib_path())
    names = _declared_symbols()
    assert len(names) >= 50
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/claymore_b200.h but not exported: {missing}"
    assert lib.cb200_version is not None


def test_binding_signatures_cover_header():
    from claymore_b200 import _capi
    bound = set(_capi._SIGNATURES) | {"cb200_version", "cb200_error_string", "cb200_sim_launch_count", "cb200_default_material"}
    assert set(_declared_symbols()) <= bound, sorted(set(_declared_symbols()) - bound)


def test_struct_layouts_match_header():
    """ctypes mirrors of the by-value structs: sizes as the C compiler lays them out."""
    import subprocess
    import tempfile
    from claymore_b200 import _capi
    prog = '#include <stdio.h>\n#include "claymore_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(cb200_config), sizeof(cb200_particle_buffer), sizeof(cb200_partition), sizeof(cb200_sim_desc), sizeof(cb200_sim_stats));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(prog)
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(d, "t")]).split()]
    assert sizes == [C.sizeof(_capi.Config), C.sizeof(_capi.ParticleBuffer), C.sizeof(_capi.Partition), C.sizeof(_capi.SimDesc), C.sizeof(_capi.SimStats)]


def test_config_derived_quantities():
    import claymore_b200 as cb
    c = cb.Config(domain_bits=9, max_ppc=64)
    assert c.grid_size == 128 and c.ppb == 4096 and abs(c.dx - 1 / 512) < 1e-12


def test_samplers_match_reference_lattice():
    from claymore_b200 import samplers
    dx = 1.0 / 128
    p = samplers.uniform_box(dx, (51, 51, 51), (77, 77, 77))
    assert p.shape == (140608, 3)                      # 26^3 cells x 8 (BASELINE config 1)
    cells = np.round(p / dx).astype(int)
    assert cells.min() == 51 and cells.max() == 76
    frac = p / dx - cells
    assert np.allclose(np.abs(frac), 0.25, atol=1e-4)  # +-0.25 dx around i*dx (GeometrySampler.h:22-29)
    s = samplers.sphere(1.0 / 64, (0.5, 0.5, 0.5), 0.1)
    assert np.all(((s - 0.5) ** 2).sum(1) <= 0.1 ** 2 + 1e-9) and 2000 < len(s) < 12000
    parts = samplers.split_slabs(s, 4)
    assert sum(len(x) for x in parts) == len(s) and max(len(x) for x in parts) - min(len(x) for x in parts) <= 1
    assert all(parts[i][:, 0].max() <= parts[i + 1][:, 0].min() + 1e-9 for i in range(3))


def test_scene_schema_mapping():
    from claymore_b200 import scene
    import claymore_b200 as cb
    assert scene.CONSTITUTIVE == {"jfluid": cb.J_FLUID, "fixed_corotated": cb.FIXED_COROTATED, "sand": cb.SAND, "nacc": cb.NACC}
    cfg = cb.Config(domain_bits=6)
    pos = scene._positions({"file": "box", "offset": [0.25, 0.25, 0.25], "span": [0.125, 0.125, 0.125]}, cfg, ".")
    assert pos.shape == (8 * 8 * 8 * 8, 3)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from claymore_b200 import _capi
    monkeypatch.setattr(_capi, "_lib", None)
    monkeypatch.setattr(_capi, "_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(_capi.CB200Error):
        _capi.lib()
