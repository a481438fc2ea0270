"""Header-only C++ mirror of the reference's `class IS3D` (include/iS3D_b200.hpp): compiles and links against the C ABI
(CPU), and on a GPU gives byte-identical result files to the file-driven entry."""
import ctypes as C
import filecmp
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

from common import load_golden, surface_columns
from is3d_b200 import api, workdir

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "class_api_example.cpp")


def build_example(out):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    libdir = os.path.dirname(api.LIB_PATH)
    cmd = [gxx, "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), SRC, "-L" + libdir, "-lis3d_b200",
           "-Wl,-rpath," + libdir, "-Wl,--allow-shlib-undefined", "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def test_class_header_compiles_and_links(tmp_path):
    exe = build_example(str(tmp_path / "class_api_example"))
    r = subprocess.run([exe, "/nonexistent/surface.dat"], capture_output=True, text=True)
    assert r.returncode == 2 and "cannot open" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("operation", [1, 0])
def test_class_api_matches_file_entry(fx, tmp_path, operation):
    gold = load_golden("s3_df4")                     # df_mode 4: the lambda/z tables must be rebuilt from the in-memory surface
    exe = build_example(str(tmp_path / "class_api_example"))
    cols = surface_columns(gold["recipe"], fx)
    wd_file, wd_mem = str(tmp_path / "file"), str(tmp_path / "mem")
    for wd in (wd_file, wd_mem):
        workdir.materialize(wd, surface_columns=cols, chosen=gold["recipe"]["chosen"], fixture=fx, operation=operation, mode=1,
                            **gold["recipe"]["params"])
    st = api.Stats()
    assert api.lib().is3d_b200_run_workdir(wd_file.encode(), None, C.c_int64(0), None, 0, C.byref(st)) == 0
    shutil.move(os.path.join(wd_mem, "input", "surface.dat"), str(tmp_path / "surface_elsewhere.dat"))
    r = subprocess.run([exe, str(tmp_path / "surface_elsewhere.dat")], cwd=wd_mem, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "cells %d skipped 0" % len(cols) in r.stdout
    sub = "results" if operation == 1 else os.path.join("results", "spacetime_distribution")
    names = sorted(f for f in os.listdir(os.path.join(wd_file, sub)) if f.endswith(".dat"))
    assert names and names == sorted(f for f in os.listdir(os.path.join(wd_mem, sub)) if f.endswith(".dat"))
    for f in names:
        assert filecmp.cmp(os.path.join(wd_file, sub, f), os.path.join(wd_mem, sub, f), shallow=False), f
