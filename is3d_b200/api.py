"""ctypes binding of the C ABI in include/is3d_b200.h (is3d_b200/libis3d_b200.so).

This is the host-side mirror used by tests, bench.py and the torch.distributed driver: it only marshals pointers.
The library is built in-tree by `is3d_b200.build`; if it is missing, or no CUDA device is present, calls fail loudly
-- there is no CPU fallback anywhere in this package.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libis3d_b200.so")
_D = C.POINTER(C.c_double)

SURFACE_FIELDS = ("tau", "eta", "dat", "dax", "day", "dan", "ux", "uy", "un", "T", "P", "E",
                  "pixx", "pixy", "pixn", "piyy", "piyn", "bulkPi", "muB", "nB", "Vx", "Vy", "Vn",
                  "pitt", "pitx", "pity", "pitn", "pinn", "Wx", "Wy", "Lambda", "aL", "c0", "c1", "c2", "c3", "c4", "x", "y")
DF_FIELDS = ("T", "c0", "c1", "c2", "c3", "c4", "F", "G", "betabulk", "betaV", "betapi")

ERRORS = {0: "IS3D_OK", 1: "IS3D_ERR_ARGUMENT", 2: "IS3D_ERR_UNSUPPORTED", 3: "IS3D_ERR_TABLE_RANGE", 4: "IS3D_ERR_CUDA",
          5: "IS3D_ERR_NO_DEVICE", 6: "IS3D_ERR_IO", 7: "IS3D_ERR_NCCL"}


class Surface(C.Structure):
    _fields_ = [("n_cells", C.c_int64)] + [(k, _D) for k in SURFACE_FIELDS]


class Species(C.Structure):
    _fields_ = [("n", C.c_int32), ("mass", _D), ("sign", _D), ("degeneracy", _D), ("baryon", _D)]


class Grid(C.Structure):
    _fields_ = [("n_pT", C.c_int32), ("n_phi", C.c_int32), ("n_y", C.c_int32), ("n_eta", C.c_int32),
                ("pT", _D), ("phi", _D), ("y", _D), ("eta", _D), ("eta_weight", _D)]


class Flags(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("mode", "df_mode", "dimension", "include_baryon", "include_bulk_deltaf",
                                         "include_shear_deltaf", "include_baryondiff_deltaf", "regulate_deltaf", "outflow")] + \
               [("deta_min", C.c_double), ("mass_pion0", C.c_double)]


class DfTables(C.Structure):
    _fields_ = [("n_T", C.c_int32)] + [(k, _D) for k in DF_FIELDS] + \
               [("n_jonah", C.c_int32), ("jonah_x", _D), ("jonah_lambda2", _D), ("jonah_z", _D),
                ("bulkPi_over_Peq_max", C.c_double)]


class Laguerre(C.Structure):
    _fields_ = [("n_points", C.c_int32), ("root1", _D), ("weight1", _D), ("root2", _D), ("weight2", _D)]


class Options(C.Structure):
    _fields_ = [("memory", C.c_int32), ("stream", C.c_void_p), ("n_chunks", C.c_int32), ("tile_variant", C.c_int32),
                ("reserved", C.c_int32 * 4)]


class Stats(C.Structure):
    _fields_ = [("cells_skipped_udsigma", C.c_int64), ("cells_feqmod_breakdown", C.c_int64), ("evaluations", C.c_int64),
                ("h2d_ms", C.c_double), ("prepare_ms", C.c_double), ("kernel_ms", C.c_double), ("reduce_ms", C.c_double),
                ("d2h_ms", C.c_double), ("total_ms", C.c_double), ("gpu_launches", C.c_int32), ("n_chunks", C.c_int32),
                ("tile_variant", C.c_int32), ("n_gpus", C.c_int32), ("allreduce_ms", C.c_double),
                ("n_chunks_wanted", C.c_int32), ("reserved", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Is3dError(RuntimeError):
    def __init__(self, code, text):
        super().__init__("%s (%d): %s" % (ERRORS.get(code, "?"), code, text))
        self.code = code


_LIB = None


def lib():
    """Load libis3d_b200.so; raises if it has not been built (no fallback)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError("%s is missing: run `python -m is3d_b200.build` (needs nvcc)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.is3d_b200_strerror.restype = C.c_char_p
        L.is3d_b200_last_error.restype = C.c_char_p
        L.is3d_b200_smooth_spectra.restype = C.c_int
        L.is3d_b200_smooth_spectra.argtypes = [C.POINTER(Flags), C.POINTER(Surface), C.POINTER(Species), C.POINTER(Grid),
                                               C.POINTER(DfTables), C.POINTER(Laguerre), C.POINTER(Options), C.c_void_p,
                                               C.POINTER(Stats)]
        L.is3d_b200_measure_fp64_peak.argtypes = [_D, _D]
        _LIB = L
    return _LIB


def _check(rc):
    if rc != 0:
        raise Is3dError(rc, lib().is3d_b200_last_error().decode() or lib().is3d_b200_strerror(rc).decode())


def init():
    _check(lib().is3d_b200_init())


def shutdown():
    _check(lib().is3d_b200_shutdown())


def init_devices(n_gpus=0):
    """One process, several GPUs: select the first n_gpus visible devices (0: all / IS3D_B200_GPUS) for the *_multi entry points."""
    _check(lib().is3d_b200_init_devices(int(n_gpus)))
    return lib().is3d_b200_device_count()


def measure_fp64_peak():
    t = C.c_double(0); ms = C.c_double(0)
    _check(lib().is3d_b200_measure_fp64_peak(C.byref(t), C.byref(ms)))
    return t.value, ms.value


def measure_fp64_sustained(seconds=3.0):
    t = C.c_double(0)
    _check(lib().is3d_b200_measure_fp64_sustained(C.c_double(seconds), C.byref(t)))
    return t.value


def surface_averages(cells):
    """T, E, P, muB, nB surface averages after the reference's 15-digit text round trip (host arrays)."""
    m = _Marshal(False)
    sf = Surface(); sf.n_cells = len(cells["tau"])
    for k in SURFACE_FIELDS:
        if k in cells and cells[k] is not None:
            setattr(sf, k, m.cells(cells[k]))
    out = np.zeros(5)
    _check(lib().is3d_b200_surface_averages(C.byref(sf), out.ctypes.data_as(_D)))
    return out


def jonah_tables(pdg_mass, pdg_degeneracy, pdg_sign, T_avg, laguerre):
    """lambda^2(Pi/P), z(Pi/P) tables for df_mode 4 (host computation in the C++ layer)."""
    m = _Marshal(False)
    x = np.zeros(301); l2 = np.zeros(301); z = np.zeros(301); mx = C.c_double(0)
    f = lib().is3d_b200_jonah_tables
    f.argtypes = [C.c_int32, _D, _D, _D, C.c_double, C.c_int32, _D, _D, _D, _D, _D, C.POINTER(C.c_double)]
    _check(f(len(pdg_mass), m.host(pdg_mass), m.host(pdg_degeneracy), m.host(pdg_sign), float(T_avg), len(laguerre["root2"]),
             m.host(laguerre["root2"]), m.host(laguerre["weight2"]), x.ctypes.data_as(_D), l2.ctypes.data_as(_D),
             z.ctypes.data_as(_D), C.byref(mx)))
    return dict(jonah_x=x, jonah_lambda2=l2, jonah_z=z, bulkPi_over_Peq_max=mx.value)


def vah_cells(columns, fx):
    """Mode-2 (31 column) surface -> SoA dict incl. alpha_L, Lambda and per-cell c0..c4 (C++ host layer, in memory)."""
    from . import synthetic
    cells = synthetic.columns_to_cells(columns, 2)
    a = np.ascontiguousarray(columns, dtype=np.float64)
    n = len(a)
    m = _Marshal(False)
    aL = np.zeros(n); Lam = np.zeros(n)
    f = lib().is3d_b200_vah_anisotropy
    f.argtypes = [C.c_int64, _D, _D, _D, _D, _D]
    _check(f(n, m.host(a[:, 13]), m.host(a[:, 14]), m.host(a[:, 15]), aL.ctypes.data_as(_D), Lam.ctypes.data_as(_D)))
    out = vah_coefficients(fx, Lam, aL)
    cells["aL"] = aL; cells["Lambda"] = Lam
    for k in range(5):
        cells["c%d" % k] = out[k]
    return cells


def vah_coefficients(fx, Lambda_GeV, aL):
    """Per-cell c0..c4 [GeV units] of the anisotropic model from the (Lambda [fm^-1], alpha_L) tables: is3d_b200_vah_coefficients
    (C++ host layer).  Raises Is3dError(IS3D_ERR_TABLE_RANGE) if a cell lies outside the table (the reference leaves garbage)."""
    m = _Marshal(False)
    n = len(aL)
    nL, naL = int(fx["df_vah/nL"]), int(fx["df_vah/naL"])
    Lg = fx["df_vah/L_col"][:nL]; ag = fx["df_vah/aL_col"][::nL]
    tabs = [np.ascontiguousarray(fx["df_vah/c%d" % k].reshape(naL, nL).T) for k in range(5)]      # -> [iL][iaL]
    out = [np.zeros(n) for _ in range(5)]
    g = lib().is3d_b200_vah_coefficients
    g.argtypes = [C.c_int32, C.c_int32, _D, _D, _D, _D, _D, _D, _D, C.c_int64, _D, _D, _D, _D, _D, _D, _D]
    rc = g(nL, naL, m.host(Lg), m.host(ag), *[m.host(t) for t in tabs], n, m.host(Lambda_GeV), m.host(aL), *[o.ctypes.data_as(_D) for o in out])
    if rc:
        raise Is3dError(rc, lib().is3d_b200_host_error().decode() or lib().is3d_b200_strerror(rc).decode())
    return out


def _is_torch(x):
    return type(x).__module__.startswith("torch")


class _Marshal:
    """Turns numpy arrays / torch tensors into double* and keeps them alive."""

    def __init__(self, device):
        self.keep = []
        self.device = device

    def host(self, x):
        a = np.ascontiguousarray(x, dtype=np.float64)
        self.keep.append(a)
        return a.ctypes.data_as(_D)

    def cells(self, x):
        if self.device:
            if not _is_torch(x) or not x.is_cuda:
                raise TypeError("memory='device' needs CUDA tensors for the surface arrays")
            import torch
            t = x.contiguous()
            if t.dtype != torch.float64:
                raise TypeError("surface tensors must be float64")
            self.keep.append(t)
            return C.cast(C.c_void_p(t.data_ptr()), _D)
        if _is_torch(x):
            x = x.detach().cpu().numpy()
        return self.host(x)


def make_flags(flags, mode=1):
    f = Flags()
    f.mode = flags.get("mode", mode)
    f.df_mode = flags["df_mode"]; f.dimension = flags["dimension"]
    f.include_baryon = flags.get("include_baryon", 0)
    f.include_bulk_deltaf = flags.get("include_bulk", flags.get("include_bulk_deltaf", 1))
    f.include_shear_deltaf = flags.get("include_shear", flags.get("include_shear_deltaf", 1))
    f.include_baryondiff_deltaf = flags.get("include_diff", flags.get("include_baryondiff_deltaf", 0))
    f.regulate_deltaf = flags.get("regulate_deltaf", 1); f.outflow = flags.get("outflow", 1)
    f.deta_min = flags.get("deta_min", 1.0e-5); f.mass_pion0 = flags.get("mass_pion0", 0.138)
    return f


def _marshal_problem(m, cells, species, grid, df_tables, laguerre):
    sf = Surface()
    n = len(cells["tau"]) if not _is_torch(cells["tau"]) else int(cells["tau"].numel())
    sf.n_cells = n
    for k in SURFACE_FIELDS:
        if k in cells and cells[k] is not None:
            setattr(sf, k, m.cells(cells[k]))
    sp = Species(); sp.n = len(species["mass"])
    for k in ("mass", "sign", "degeneracy", "baryon"):
        setattr(sp, k, m.host(species[k]))
    g = Grid()
    g.n_pT, g.n_phi, g.n_y, g.n_eta = len(grid["pT"]), len(grid["phi"]), len(grid["y"]), len(grid["eta"])
    for k in ("pT", "phi", "y", "eta", "eta_weight"):
        setattr(g, k, m.host(grid[k]))
    dft = DfTables()
    if df_tables is not None:
        dft.n_T = len(df_tables["T"])
        for k in DF_FIELDS:
            if df_tables.get(k) is not None:
                setattr(dft, k, m.host(df_tables[k]))
        if df_tables.get("jonah_x") is not None:
            dft.n_jonah = len(df_tables["jonah_x"])
            dft.jonah_x = m.host(df_tables["jonah_x"]); dft.jonah_lambda2 = m.host(df_tables["jonah_lambda2"])
            dft.jonah_z = m.host(df_tables["jonah_z"]); dft.bulkPi_over_Peq_max = float(df_tables["bulkPi_over_Peq_max"])
    la = Laguerre()
    if laguerre is not None:
        la.n_points = len(laguerre["root1"])
        for k in ("root1", "weight1", "root2", "weight2"):
            setattr(la, k, m.host(laguerre[k]))
    return sf, sp, g, dft, la


def smooth_spectra(flags, cells, species, grid, df_tables=None, laguerre=None, out=None, memory="host",
                   stream=None, n_chunks=0, tile_variant=0, multi=False):
    """Call is3d_b200_smooth_spectra.

    cells: dict of per-cell arrays (numpy for memory='host', float64 CUDA tensors for memory='device').
    Returns (dN, stats dict); dN is flat [y][phi][pT][species] (species fastest), numpy or the CUDA tensor `out`.
    The result is ADDED into `out` when given.
    """
    device = (memory == "device")
    m = _Marshal(device)
    sf, sp, g, dft, la = _marshal_problem(m, cells, species, grid, df_tables, laguerre)
    n_bins = sp.n * g.n_pT * g.n_phi * g.n_y
    if device:
        import torch
        if out is None:
            out = torch.zeros(n_bins, dtype=torch.float64, device=cells["tau"].device)
        out_ptr = C.c_void_p(out.data_ptr())
        if stream is None:
            stream = torch.cuda.current_stream().cuda_stream
    else:
        if out is None:
            out = np.zeros(n_bins)
        assert out.dtype == np.float64 and out.flags["C_CONTIGUOUS"] and out.size == n_bins
        out_ptr = C.c_void_p(out.ctypes.data)
    opt = Options(); opt.memory = 1 if device else 0
    opt.stream = C.c_void_p(stream or 0); opt.n_chunks = n_chunks; opt.tile_variant = tile_variant
    st = Stats()
    fl = make_flags(flags)
    fn = lib().is3d_b200_smooth_spectra_multi if multi else lib().is3d_b200_smooth_spectra      # multi: host arrays, all devices of init_devices()
    fn.restype = C.c_int
    fn.argtypes = lib().is3d_b200_smooth_spectra.argtypes
    rc = fn(C.byref(fl), C.byref(sf), C.byref(sp), C.byref(g), C.byref(dft), C.byref(la), C.byref(opt), out_ptr, C.byref(st))
    _check(rc)
    return out, st.as_dict()


class SpacetimeBins(C.Structure):
    _fields_ = [("tau_min", C.c_double), ("tau_max", C.c_double), ("r_min", C.c_double), ("r_max", C.c_double),
                ("tau_bins", C.c_int32), ("r_bins", C.c_int32), ("pT_weight", _D), ("phi_weight", _D)]


class SpacetimeResult(C.Structure):
    _fields_ = [(k, _D) for k in ("dN_tau", "dN_r", "dN_taur", "dN_dydeta", "dN_dy")]


def spacetime_distributions(flags, cells, species, grid, df_tables, laguerre, bins, memory="host", stream=None,
                            n_chunks=0, tile_variant=0, multi=False):
    """Call is3d_b200_spacetime_distributions (operation = 0).

    cells must also carry the transverse positions "x", "y"; grid the quadrature weights "pT_weight", "phi_weight";
    bins = dict(tau_min, tau_max, tau_bins, r_min, r_max, r_bins).  Returns (dict of raw sums as numpy arrays:
    dN_tau [species][tau], dN_r [species][r], dN_taur [species][tau][r], dN_dydeta [species][eta_pts], dN_dy [species]; stats)."""
    device = (memory == "device")
    m = _Marshal(device)
    sf, sp, g, dft, la = _marshal_problem(m, cells, species, grid, df_tables, laguerre)
    b = SpacetimeBins()
    b.tau_min, b.tau_max, b.r_min, b.r_max = (float(bins[k]) for k in ("tau_min", "tau_max", "r_min", "r_max"))
    b.tau_bins, b.r_bins = int(bins["tau_bins"]), int(bins["r_bins"])
    if cells.get("x") is None or cells.get("y") is None:
        raise KeyError("spacetime_distributions needs the transverse cell positions cells['x'], cells['y']")
    b.pT_weight = m.host(grid["pT_weight"]); b.phi_weight = m.host(grid["phi_weight"])
    ns = sp.n; eta_pts = g.n_eta if flags["dimension"] == 2 else 1
    out = dict(dN_tau=np.zeros((ns, b.tau_bins)), dN_r=np.zeros((ns, b.r_bins)), dN_taur=np.zeros((ns, b.tau_bins, b.r_bins)),
               dN_dydeta=np.zeros((ns, eta_pts)), dN_dy=np.zeros(ns))
    res = SpacetimeResult()
    for k in out:
        setattr(res, k, out[k].ctypes.data_as(_D))
    if device and stream is None:
        import torch
        stream = torch.cuda.current_stream().cuda_stream
    opt = Options(); opt.memory = 1 if device else 0
    opt.stream = C.c_void_p(stream or 0); opt.n_chunks = n_chunks; opt.tile_variant = tile_variant
    st = Stats()
    fl = make_flags(flags)
    f = lib().is3d_b200_spacetime_distributions_multi if multi else lib().is3d_b200_spacetime_distributions
    f.restype = C.c_int
    f.argtypes = [C.POINTER(Flags), C.POINTER(Surface), C.POINTER(Species), C.POINTER(Grid), C.POINTER(DfTables),
                  C.POINTER(Laguerre), C.POINTER(SpacetimeBins), C.POINTER(Options), C.POINTER(SpacetimeResult), C.POINTER(Stats)]
    _check(f(C.byref(fl), C.byref(sf), C.byref(sp), C.byref(g), C.byref(dft), C.byref(la), C.byref(b), C.byref(opt),
             C.byref(res), C.byref(st)))
    return out, st.as_dict()


def particle_densities(species, avg5, df_mode, df_tables, laguerre3):
    """(n_eq, dn_bulk, dn_diff) per species at the surface averages avg5 = (T, E, P, muB, nB) (host computation in the C++ layer).

    laguerre3: dict root1, weight1, root2, weight2, root3, weight3 (alpha = 1, 2, 3 rows of the Gauss-Laguerre table)."""
    m = _Marshal(False)
    n = len(species["mass"])
    dft = DfTables(); dft.n_T = len(df_tables["T"])
    for k in DF_FIELDS:
        if df_tables.get(k) is not None:
            setattr(dft, k, m.host(df_tables[k]))
    neq = np.zeros(n); bulk = np.zeros(n); diff = np.zeros(n)
    f = lib().is3d_b200_particle_densities
    f.restype = C.c_int
    f.argtypes = [C.c_int32, _D, _D, _D, _D, _D, C.c_int32, C.POINTER(DfTables), C.c_int32, _D, _D, _D, _D, _D, _D, _D, _D, _D]
    rc = f(n, m.host(species["mass"]), m.host(species["degeneracy"]), m.host(species["baryon"]), m.host(species["sign"]), m.host(avg5),
           int(df_mode), C.byref(dft), len(laguerre3["root1"]), m.host(laguerre3["root1"]), m.host(laguerre3["weight1"]),
           m.host(laguerre3["root2"]), m.host(laguerre3["weight2"]), m.host(laguerre3["root3"]), m.host(laguerre3["weight3"]),
           neq.ctypes.data_as(_D), bulk.ctypes.data_as(_D), diff.ctypes.data_as(_D))
    if rc:
        raise Is3dError(rc, lib().is3d_b200_host_error().decode() or lib().is3d_b200_strerror(rc).decode())
    return neq, bulk, diff


def mean_yield(flags, cells, neq, dn_bulk, df_tables=None, y_cut=5.0, memory="host", stream=None):
    """Call is3d_b200_mean_yield (the sampler's total-yield estimate); returns (Ntot, stats dict)."""
    device = (memory == "device")
    m = _Marshal(device)
    sf = Surface()
    sf.n_cells = len(cells["tau"]) if not _is_torch(cells["tau"]) else int(cells["tau"].numel())
    for k in ("tau", "ux", "uy", "un", "dat", "dax", "day", "dan", "bulkPi", "P"):
        if k in cells and cells[k] is not None:
            setattr(sf, k, m.cells(cells[k]))
    dft = DfTables()
    if df_tables is not None and df_tables.get("jonah_x") is not None:
        dft.n_jonah = len(df_tables["jonah_x"])
        dft.jonah_x = m.host(df_tables["jonah_x"]); dft.jonah_lambda2 = m.host(df_tables["jonah_lambda2"])
        dft.jonah_z = m.host(df_tables["jonah_z"]); dft.bulkPi_over_Peq_max = float(df_tables["bulkPi_over_Peq_max"])
    if device and stream is None:
        import torch
        stream = torch.cuda.current_stream().cuda_stream
    opt = Options(); opt.memory = 1 if device else 0; opt.stream = C.c_void_p(stream or 0)
    st = Stats(); out = C.c_double(0.0)
    f = lib().is3d_b200_mean_yield
    f.restype = C.c_int
    f.argtypes = [C.POINTER(Flags), C.POINTER(Surface), C.c_int32, _D, _D, C.POINTER(DfTables), C.c_double, C.POINTER(Options),
                  C.POINTER(C.c_double), C.POINTER(Stats)]
    fl = make_flags(flags)
    _check(f(C.byref(fl), C.byref(sf), len(neq), m.host(neq), m.host(dn_bulk) if dn_bulk is not None else None, C.byref(dft),
             float(y_cut), C.byref(opt), C.byref(out), C.byref(st)))
    return out.value, st.as_dict()


class ParticleList(C.Structure):
    _I = C.POINTER(C.c_int32)
    _fields_ = [("n_particles", C.c_int32), ("mcid", _I), ("mass", _D), ("width", _D), ("stable", _I), ("decays", _I), ("dec_first", _I),
                ("dec_npart", _I), ("dec_br", _D), ("dec_part", _I)]


def resonance_decays(pdg, chosen_pdg_index, grid, dimension, dN, memory="host", stream=None):
    """Call is3d_b200_resonance_decays: feed-down of the unstable chosen species into their chosen daughters.

    pdg: dict of arrays mcid, mass, width, stable, decays, dec_first, dec_npart, dec_br, dec_part[rows, 5] (tables.pdg_decay_table);
    dN: numpy array (memory='host'; a copy is amended and returned) or float64 CUDA tensor (memory='device'; amended in place)."""
    keep = []

    def ia(x):
        a = np.ascontiguousarray(x, dtype=np.int32); keep.append(a)
        return a.ctypes.data_as(C.POINTER(C.c_int32))

    m = _Marshal(False)
    p = ParticleList()
    p.n_particles = len(pdg["mcid"])
    p.mcid = ia(pdg["mcid"]); p.mass = m.host(pdg["mass"]); p.width = m.host(pdg["width"]); p.stable = ia(pdg["stable"])
    p.decays = ia(pdg["decays"]); p.dec_first = ia(pdg["dec_first"]); p.dec_npart = ia(pdg["dec_npart"]); p.dec_br = m.host(pdg["dec_br"])
    p.dec_part = ia(np.asarray(pdg["dec_part"]).ravel())
    g = Grid()
    g.n_pT, g.n_phi, g.n_y, g.n_eta = len(grid["pT"]), len(grid["phi"]), len(grid["y"]), len(grid["eta"])
    for k in ("pT", "phi", "y", "eta", "eta_weight"):
        setattr(g, k, m.host(grid[k]))
    device = (memory == "device")
    if device:
        import torch
        out = dN
        ptr = C.c_void_p(out.data_ptr())
        if stream is None:
            stream = torch.cuda.current_stream().cuda_stream
    else:
        out = np.array(dN, dtype=np.float64, copy=True)
        ptr = C.c_void_p(out.ctypes.data)
    opt = Options(); opt.memory = 1 if device else 0; opt.stream = C.c_void_p(stream or 0)
    st = Stats()
    f = lib().is3d_b200_resonance_decays
    f.restype = C.c_int
    f.argtypes = [C.POINTER(ParticleList), C.c_int32, C.POINTER(C.c_int32), C.POINTER(Grid), C.c_int32, C.POINTER(Options), C.c_void_p, C.POINTER(Stats)]
    _check(f(C.byref(p), len(chosen_pdg_index), ia(chosen_pdg_index), C.byref(g), int(dimension), C.byref(opt), ptr, C.byref(st)))
    return out, st.as_dict()
