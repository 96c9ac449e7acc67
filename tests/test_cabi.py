"""The C-ABI shared library: loads without a GPU, exports every symbol include/pgi.h and include/pgb.h declare, has
no CPU compute path (entry points fail loudly without a device) and does not link anything from oracle/."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pg[ib]_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported():
    from pose_graph_initialization_b200 import builder, engine

    lib = engine.load_library()
    builder._host_lib()
    for header in ("pgi.h", "pgb.h"):
        names = _declared(header)
        assert len(names) >= 10
        for n in names:
            assert hasattr(lib, n), f"{n} declared in include/{header} but not exported"
    assert set(engine.EXPORTS) == set(_declared("pgi.h"))
    assert set(builder.PGB_EXPORTS) == set(_declared("pgb.h"))


def test_struct_layouts_match_the_header():
    from pose_graph_initialization_b200 import builder, engine

    assert engine.VERDICT_DTYPE.itemsize == 160  # pgi_verdict, also the all-gather element
    assert builder.ITEM_DTYPE.itemsize == 72 and builder.EDGE_DTYPE.itemsize == 88 and builder.LOG_DTYPE.itemsize == 232
    assert builder.RECORD_DTYPE.itemsize == 224 and engine.ADJ_ENTRY_DTYPE.itemsize == 24 and engine.SEARCH_RESULT_DTYPE.itemsize == 32


def test_no_cpu_fallback_without_a_device():
    import torch
    from pose_graph_initialization_b200 import engine

    lib = engine.load_library()
    assert lib.pgi_version().startswith(b"pgi")
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the -m gpu tests")
    assert lib.pgi_device_count() == 0
    cfg = engine.PgiConfig(0, 20, 5, 1000, 1.5, 64, 0)
    h = C.c_void_p()
    assert lib.pgi_create(C.byref(cfg), C.byref(h)) == -2  # PGI_ERR_CUDA: nothing to fall back to
    with pytest.raises(engine.PgiError):
        engine.Engine()


def test_product_library_does_not_link_the_oracle():
    from pose_graph_initialization_b200 import engine

    out = subprocess.run(["ldd", engine.library_path()], capture_output=True, text=True).stdout
    assert "pgo_oracle" not in out
    # and no product source mentions the oracle directory
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pose_graph_initialization_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pgo_oracle" not in txt and "import oracle" not in txt and "from oracle" not in txt, f


def test_scene_container_round_trip(tmp_path):
    from pose_graph_initialization_b200 import scene as S

    sc = S.make_scene(n_views=5, n_corr=60, outlier_ratio=0.3, seed=3, n_points=200)
    p = str(tmp_path / "scene.pgi")
    S.save_scene(p, sc)
    back = S.load_scene(p)
    for k in ("focal", "size", "sim", "kp_offset", "kp", "pair_views", "m_offset", "matches"):
        assert np.array_equal(np.asarray(sc[k]), back[k]), k
    assert np.allclose(np.diag(sc["sim"]), 1.0) and sc["sim"][np.triu_indices(5, 1)].max() < 1.0


def test_pyposegraphbuilder_module_is_importable_and_checks_shapes():
    import numpy as np
    import pyposegraphbuilder as ppg

    assert ppg.PoseGraphBuilder is not None and callable(ppg.estimate_pose) and callable(ppg.test_pose) and callable(ppg.guided_match)
    with pytest.raises(ValueError):
        ppg.estimate_pose(np.zeros((5, 3)), 1e-3)
    with pytest.raises(ValueError):
        ppg.test_pose(np.zeros((5, 4)), 1e-3, np.zeros(6))
