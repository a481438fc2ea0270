#!/usr/bin/env python
"""Build is3d_b200/data/is3d_tables.npz from the reference checkout's DATA files (not sources).

Run in the build container, where /root/reference exists:
    python tests/golden/make_tables_fixture.py [/root/reference]

The fixture carries the *numeric content* of the particle lists, quadrature tables and delta-f coefficient
tables that the smooth Cooper-Frye path consumes, so that the GPU box (which has no /root/reference) can
re-create an iS3D-style working directory (is3d_b200.workdir.materialize) for the parity tests, the
reference arm of bench.py and the smoke test.  Every number is parsed with Python's correctly rounded
float(), i.e. to the same double the reference's `istream >> double` / `fscanf("%lf")` produce.

Files read (all relative to the reference root):
  PDG/pdg-urqmd_v3.3+.dat, PDG/pdg_smash.dat      12-column header + decay lines (readindata.cpp:1440-1568)
  PDG/pdg_box.dat                                  SMASH box format (readindata.cpp:1571-1684)
  PDG/chosen_particles*.dat                        one MC id per line
  tables/pT_gauss_legendre_table.dat, phi_gauss_legendre_table.dat, y_trapezoid_table_21pt.dat,
  tables/eta/eta_trapezoid_table_241pt.dat, tables/gla_roots_weights_32_points.txt,
  tables/gauss_legendre_48pts.dat                  (iS3D.cpp:161-167, emissionfunction.cpp:1311,1315)
  deltaf_coefficients/vh/{urqmd,smash,smash_box}/*.dat   muB = 0 block only (deltafReader.cpp:65-219)
  deltaf_coefficients/vah/c{0..4}_vah1.dat         (src/cuda/deltafReader.cu:192-277)
  input/surface.dat                                the shipped one-cell surface
"""
import os
import sys

import numpy as np


def read_table(path):
    rows = []
    with open(path) as f:
        for line in f:
            t = line.split()
            if t:
                rows.append([float(v) for v in t])
    return np.array(rows, dtype=np.float64)


def read_pdg_conventional(path):
    tok = open(path).read().split()
    i = 0
    out = {k: [] for k in ("mcid", "name", "mass", "width", "gspin", "baryon", "strange", "charm", "bottom",
                           "gisospin", "charge", "decays")}
    dec_n, dec_br, dec_parts, dec_owner = [], [], [], []
    while i < len(tok):
        mcid = int(tok[i]); name = tok[i + 1]
        mass, width = float(tok[i + 2]), float(tok[i + 3])
        ints = [int(v) for v in tok[i + 4:i + 12]]
        i += 12
        for k, v in zip(("gspin", "baryon", "strange", "charm", "bottom", "gisospin", "charge", "decays"), ints):
            out[k].append(v)
        out["mcid"].append(mcid); out["name"].append(name); out["mass"].append(mass); out["width"].append(width)
        for _ in range(ints[7]):
            # mcid, n_daughters, branching ratio, 5 daughter ids
            dec_owner.append(len(out["mcid"]) - 1)
            dec_n.append(int(tok[i + 1])); dec_br.append(float(tok[i + 2]))
            dec_parts.append([int(v) for v in tok[i + 3:i + 8]])
            i += 8
    res = {k: np.array(v) for k, v in out.items()}
    res["dec_owner"] = np.array(dec_owner, dtype=np.int64)
    res["dec_n"] = np.array(dec_n, dtype=np.int64)
    res["dec_br"] = np.array(dec_br, dtype=np.float64)
    res["dec_parts"] = np.array(dec_parts, dtype=np.int64)
    return res


def read_pdg_box(path):
    names, mass, width, parity, ids = [], [], [], [], []
    for line in open(path, encoding="utf-8"):
        s = line.split("#")[0].split()
        if len(s) < 5:
            continue
        names.append(s[0]); mass.append(float(s[1])); width.append(float(s[2])); parity.append(s[3])
        row = [int(v) for v in s[4:8]] + [0, 0, 0, 0]
        ids.append(row[:4])
    return dict(name=np.array(names), mass=np.array(mass), width=np.array(width), parity=np.array(parity),
                mcid=np.array(ids, dtype=np.int64))


def read_df_vh(dirname):
    out = {}
    for name in ("c0", "c1", "c2", "c3", "c4", "F", "G", "betabulk", "betaV", "betapi"):
        with open(os.path.join(dirname, name + ".dat")) as f:
            nT = int(f.readline()); nB = int(f.readline()); f.readline()
            rows = [f.readline().split() for _ in range(nT)]          # first (muB = 0) block
        out["T"] = np.array([float(r[0]) for r in rows])
        out["muB0"] = np.array([float(r[1]) for r in rows])
        out[name] = np.array([float(r[2]) for r in rows])
        out["points_T"] = np.array(nT); out["points_muB"] = np.array(nB)
    return out


def read_df_vah(dirname):
    out = {}
    for k in range(5):
        with open(os.path.join(dirname, "c%d_vah1.dat" % k)) as f:
            nL = int(f.readline()); naL = int(f.readline()); f.readline()
            a = np.array([[float(v) for v in f.readline().split()] for _ in range(nL * naL)])
        out["nL"] = np.array(nL); out["naL"] = np.array(naL)
        # keep the file's row order (whatever it is) so the workdir writer reproduces it
        out["L_col"] = a[:, 0]; out["aL_col"] = a[:, 1]; out["c%d" % k] = a[:, 2]
    return out


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    here = os.path.dirname(os.path.abspath(__file__))
    fx = {}
    for key, fn in (("pdg_urqmd", "PDG/pdg-urqmd_v3.3+.dat"), ("pdg_smash", "PDG/pdg_smash.dat")):
        for k, v in read_pdg_conventional(os.path.join(ref, fn)).items():
            fx["%s/%s" % (key, k)] = v
    for k, v in read_pdg_box(os.path.join(ref, "PDG/pdg_box.dat")).items():
        fx["pdg_box/%s" % k] = v
    for key, fn in (("chosen_pikp", "chosen_particles_pikp.dat"), ("chosen_urqmd", "chosen_particles_urqmd_v3.3+.dat"),
                    ("chosen_smash", "chosen_particles_smash.dat"), ("chosen_box", "chosen_particles_box.dat")):
        fx[key] = read_table(os.path.join(ref, "PDG", fn))[:, 0].astype(np.int64)
    fx["pT_tab"] = read_table(os.path.join(ref, "tables/pT_gauss_legendre_table.dat"))
    fx["phi_tab"] = read_table(os.path.join(ref, "tables/phi_gauss_legendre_table.dat"))
    fx["y_tab"] = read_table(os.path.join(ref, "tables/y_trapezoid_table_21pt.dat"))
    fx["eta_tab"] = read_table(os.path.join(ref, "tables/eta/eta_trapezoid_table_241pt.dat"))
    gla = open(os.path.join(ref, "tables/gla_roots_weights_32_points.txt")).read().split()
    na, npts = int(gla[0]), int(gla[1])
    body = np.array([float(v) for v in gla[2:2 + 3 * na * npts]]).reshape(na, npts, 3)
    fx["gla_root"] = body[:, :, 1].copy(); fx["gla_weight"] = body[:, :, 2].copy()
    gl48 = open(os.path.join(ref, "tables/gauss_legendre_48pts.dat")).read().split()
    fx["legendre48"] = np.array([float(v) for v in gl48[1:1 + 2 * int(gl48[0])]]).reshape(-1, 2)
    for eos in ("urqmd", "smash", "smash_box"):
        for k, v in read_df_vh(os.path.join(ref, "deltaf_coefficients/vh", eos)).items():
            fx["df_%s/%s" % (eos, k)] = v
    for k, v in read_df_vah(os.path.join(ref, "deltaf_coefficients/vah")).items():
        fx["df_vah/%s" % k] = v
    fx["toy_surface"] = read_table(os.path.join(ref, "input/surface.dat"))
    out = os.path.join(os.path.dirname(os.path.dirname(here)), "is3d_b200", "data", "is3d_tables.npz")
    np.savez_compressed(out, **fx)
    print("wrote", out, os.path.getsize(out), "bytes;", len(fx), "arrays")


if __name__ == "__main__":
    main()
