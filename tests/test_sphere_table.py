"""Sphere-growing table generator against the reference's shipped
tables (by SHA-256 -- the tables themselves stay in the reference) and against the literal
shell sweep of CI.getSpherePix (oracle/sphere_oracle.py)."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle.sphere_oracle import sphere_table_sweep
from vent_analysis_b200 import CI
from vent_analysis_b200.sphere_table import generate, get_sphere_pix, table_file_name


def test_generator_reproduces_shipped_tables_byte_for_byte():
    meta = json.load(open(os.path.join(GOLDEN, "sphere_tables.json")))
    for name, m in meta.items():
        t = generate(m["vox"], m["radius"])
        assert list(t.shape) == m["shape"] and t.dtype == np.float64
        assert hashlib.sha256(t.tobytes()).hexdigest() == m["sha256"], name
        assert table_file_name(m["vox"], m["radius"]) == name
        # the 4 offsets (+-32,0,0),(0,+-32,0) sit in two consecutive shells
        _, cnt = np.unique(t[:, 1:], axis=0, return_counts=True)
        assert (cnt > 1).sum() == 4


@pytest.mark.parametrize("vox", [[3.9, 3.9, 3.9], [1.0, 2.5, 3.3], [2.0, 2.0, 11.5], [1.5, 1.5, 10.0]])
def test_generator_equals_literal_sweep(vox):
    assert np.array_equal(generate(vox, 7), sphere_table_sweep(vox, 7))


def test_get_sphere_pix_loads_from_cwd_like_the_reference(tmp_path, monkeypatch):
    t = generate([1.0, 1.0, 2.0], 3)
    fake = t.copy()
    fake[1, 0] = 123.0
    np.save(tmp_path / table_file_name([1.0, 1.0, 2.0], 3), fake)
    monkeypatch.chdir(tmp_path)
    assert CI.getSpherePix([1.0, 1.0, 2.0], 3)[1, 0] == 123.0       # CI.py:44-46
    assert np.array_equal(get_sphere_pix([1.0, 1.0, 2.0], 3, search_cwd=False), t)


def test_radii_indices_and_px2vec_mirror_reference():
    t = generate([1.5, 1.5, 10.0], 5)
    b = CI.getRadiiIndices(t)
    assert b[0] == 1 and b[1] == 5 and b[2] == 9
    assert CI.px2vec(3, 4, 5, (10, 20, 30)) == 3 + 3 * 10 + 4 * 200
    assert CI.vec2px(CI.px2vec(3, 4, 5, (10, 20, 30)), (10, 20, 30)) == (3, 4, 5)
