"""ctypes binding of the CPU oracle (oracle/liblcs_oracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, bench.py's cpu_baseline / --impl reference leg and __graft_entry__.smoke().
The product package never imports this module."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liblcs_oracle.so")

LEGACY_MATLAB = 1
F64 = 2


class Cell(C.Structure):
    """POD mirror of class Cell (reference include/common.h.in:101-129)."""
    _fields_ = [
        ("fc_requested", C.c_double), ("fc_programmed", C.c_double), ("pss_pow", C.c_double),
        ("ind", C.c_int32), ("freq", C.c_double), ("n_id_2", C.c_int32), ("n_id_1", C.c_int32),
        ("cp_type", C.c_int32), ("frame_start", C.c_double), ("freq_fine", C.c_double),
        ("freq_superfine", C.c_double), ("n_ports", C.c_int32), ("n_rb_dl", C.c_int32),
        ("phich_duration", C.c_int32), ("phich_resource", C.c_int32), ("sfn", C.c_int32),
    ]

    def n_id_cell(self):
        return self.n_id_2 + 3 * self.n_id_1 if (self.n_id_1 >= 0 and self.n_id_2 >= 0) else -1

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def build(force=False):
    src = [os.path.join(HERE, f) for f in ("lcs_oracle.cpp", "lcs_oracle_c.cpp", "lcs_oracle.hpp", "Makefile")]
    if (not force and os.path.exists(LIB_PATH)
            and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return LIB_PATH
    subprocess.check_call(["make", "-C", HERE, "-s"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.lcso_chi2cdf_inv.restype = C.c_double
        _lib.lcso_chi2cdf_inv.argtypes = [C.c_double, C.c_double]
        assert _lib.lcso_cell_sizeof() == C.sizeof(Cell)
    return _lib


def _p(a, t=C.c_void_p):
    return None if a is None else a.ctypes.data_as(t)


def new_cell(**kw):
    c = Cell()
    lib().lcso_cell_init(C.byref(c))
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def max_threads():
    return lib().lcso_max_threads()


def set_threads(n):
    lib().lcso_set_threads(int(n))


def chi2cdf_inv(p, k):
    return lib().lcso_chi2cdf_inv(float(p), float(k))


def lte_pn(c_init, n):
    out = np.zeros(n, np.uint8)
    lib().lcso_lte_pn(C.c_uint32(c_init), C.c_uint32(n), _p(out))
    return out


def pss_td(t):
    out = np.zeros(137, np.complex128)
    lib().lcso_pss_td(t, _p(out))
    return out


def pss_fd(t):
    out = np.zeros(62, np.complex128)
    lib().lcso_pss_fd(t, _p(out))
    return out


def sss_fd(n_id_1, n_id_2, slot):
    out = np.zeros(62, np.int32)
    lib().lcso_sss_fd(n_id_1, n_id_2, slot, _p(out))
    return out


def rs_dl(n_id_cell, cp_type):
    n_symb = 7 if cp_type == 1 else 6
    rs = np.zeros((20 * n_symb, 12), np.complex128)
    sh = np.zeros((20 * n_symb, 4), np.float64)
    lib().lcso_rs_dl(n_id_cell, cp_type, _p(rs), _p(sh))
    return rs, sh


def conv_encode(c):
    c = np.ascontiguousarray(c, np.uint8)
    d = np.zeros((3, c.size), np.uint8)
    lib().lcso_conv_encode(_p(c), c.size, _p(d))
    return d


def conv_decode(d_est):
    d_est = np.ascontiguousarray(d_est, np.float64)
    n_c = d_est.shape[1]
    c = np.zeros(n_c, np.uint8)
    lib().lcso_conv_decode(_p(d_est), n_c, _p(c))
    return c


def crc16(a):
    a = np.ascontiguousarray(a, np.uint8)
    p = np.zeros(16, np.uint8)
    lib().lcso_crc16(_p(a), a.size, _p(p))
    return p


def deratematch(e, n_c):
    e = np.ascontiguousarray(e, np.float64)
    d = np.zeros((3, n_c), np.float64)
    lib().lcso_deratematch(_p(e), e.size, n_c, _p(d))
    return d


def f_search_set(freq_start, ppm):
    n = C.c_int(0)
    lib().lcso_f_search_set(C.c_double(freq_start), C.c_double(ppm), None, C.byref(n))
    out = np.zeros(n.value, np.float64)
    lib().lcso_f_search_set(C.c_double(freq_start), C.c_double(ppm), _p(out), C.byref(n))
    return out


def xcorr_pss(capbuf, f_search_set, ds_comb_arm, fc_requested, fc_programmed, fs_programmed, flags=0,
              want_xc=False, want_sp=True):
    """searcher.h:22-41.  Returns a dict of the reference's outputs (single/incoherent as [3][9600][n_f])."""
    capbuf = np.ascontiguousarray(capbuf, np.complex128)
    f = np.ascontiguousarray(f_search_set, np.float64)
    n_cap, n_f = capbuf.size, f.size
    pw = np.zeros((3, 9600)); frq = np.zeros((3, 9600), np.int32)
    single = np.zeros((3, 9600, n_f)); inc = np.zeros((3, 9600, n_f)); spi = np.zeros(9600)
    xc = np.zeros((3, n_cap - 136, n_f), np.complex128) if want_xc else None
    sp = np.zeros(((n_cap - 273) // 9600) * 9600) if want_sp else None
    ncx, ncs = C.c_uint16(0), C.c_uint16(0)
    lib().lcso_xcorr_pss(_p(capbuf), C.c_uint32(n_cap), _p(f), n_f, int(ds_comb_arm), C.c_double(fc_requested),
                         C.c_double(fc_programmed), C.c_double(fs_programmed), C.c_uint32(flags), _p(pw), _p(frq),
                         _p(single), _p(inc), _p(spi), _p(xc), _p(sp), C.byref(ncx), C.byref(ncs))
    return dict(pow=pw, frq=frq, single=single, incoherent=inc, sp_incoherent=spi, xc=xc, sp=sp,
                n_comb_xc=ncx.value, n_comb_sp=ncs.value)


def calc_Z_th1(sp_incoherent, n_comb_xc, ds_comb_arm):
    s = np.ascontiguousarray(sp_incoherent, np.float64)
    z = np.zeros(9600)
    lib().lcso_calc_Z_th1(_p(s), int(n_comb_xc), int(ds_comb_arm), _p(z))
    return z


def peak_search(pw, frq, Z_th1, f_search_set, fc_requested, fc_programmed, single, ds_comb_arm, max_cells=256):
    pw = np.ascontiguousarray(pw, np.float64); frq = np.ascontiguousarray(frq, np.int32)
    z = np.ascontiguousarray(Z_th1, np.float64); f = np.ascontiguousarray(f_search_set, np.float64)
    single = np.ascontiguousarray(single, np.float64)
    cells = (Cell * max_cells)()
    n = lib().lcso_peak_search(_p(pw), _p(frq), _p(z), _p(f), f.size, C.c_double(fc_requested),
                               C.c_double(fc_programmed), _p(single), int(ds_comb_arm), cells, max_cells)
    return [cells[i] for i in range(min(n, max_cells))]


def _copy(c):
    o = Cell()
    C.memmove(C.byref(o), C.byref(c), C.sizeof(Cell))
    return o


def sss_detect(cell, capbuf, thresh2_n_sigma, fc_requested, fc_programmed, fs_programmed, flags=0):
    capbuf = np.ascontiguousarray(capbuf, np.complex128)
    out = Cell(); dbg = np.zeros(1292)
    lib().lcso_sss_detect(C.byref(cell), _p(capbuf), C.c_uint32(capbuf.size), C.c_double(thresh2_n_sigma),
                          C.c_double(fc_requested), C.c_double(fc_programmed), C.c_double(fs_programmed),
                          C.c_uint32(flags), C.byref(out), _p(dbg))
    d = dict(h1_np=dbg[0:62].copy(), h2_np=dbg[62:124].copy(),
             h1_nrm=dbg[124:248].view(np.complex128).copy(), h2_nrm=dbg[248:372].view(np.complex128).copy(),
             h1_ext=dbg[372:496].view(np.complex128).copy(), h2_ext=dbg[496:620].view(np.complex128).copy(),
             log_lik_nrm=dbg[620:956].reshape(168, 2).copy(), log_lik_ext=dbg[956:1292].reshape(168, 2).copy())
    return out, d


def pss_sss_foe(cell, capbuf, fc_requested, fc_programmed, fs_programmed, flags=0):
    capbuf = np.ascontiguousarray(capbuf, np.complex128)
    out = Cell()
    lib().lcso_pss_sss_foe(C.byref(cell), _p(capbuf), C.c_uint32(capbuf.size), C.c_double(fc_requested),
                           C.c_double(fc_programmed), C.c_double(fs_programmed), C.c_uint32(flags), C.byref(out))
    return out


def extract_tfg(cell, capbuf, fc_requested, fc_programmed, fs_programmed, flags=0):
    capbuf = np.ascontiguousarray(capbuf, np.complex128)
    tfg = np.zeros((854, 72), np.complex128); ts = np.zeros(854)
    n = lib().lcso_extract_tfg(C.byref(cell), _p(capbuf), C.c_uint32(capbuf.size), C.c_double(fc_requested),
                               C.c_double(fc_programmed), C.c_double(fs_programmed), C.c_uint32(flags), _p(tfg), _p(ts))
    return tfg.reshape(-1)[:n * 72].reshape(n, 72).copy(), ts[:n].copy()


def tfoec(cell, tfg, ts, fc_requested, fc_programmed, flags=0):
    tfg = np.ascontiguousarray(tfg, np.complex128); ts = np.ascontiguousarray(ts, np.float64)
    out = Cell(); tc = np.zeros_like(tfg); tsc = np.zeros_like(ts)
    lib().lcso_tfoec(C.byref(cell), _p(tfg), _p(ts), ts.size, C.c_double(fc_requested), C.c_double(fc_programmed),
                     C.c_uint32(flags), _p(tc), _p(tsc), C.byref(out))
    return out, tc, tsc


def chan_est(cell, tfg, port):
    tfg = np.ascontiguousarray(tfg, np.complex128)
    ce = np.zeros_like(tfg); npw = C.c_double(0)
    lib().lcso_chan_est(C.byref(cell), _p(tfg), tfg.shape[0], int(port), _p(ce), C.byref(npw))
    return ce, npw.value


def decode_mib(cell, tfg):
    tfg = np.ascontiguousarray(tfg, np.complex128)
    out = Cell(); bits = np.zeros(40, np.uint8); npv = np.zeros(4)
    g = lib().lcso_decode_mib(C.byref(cell), _p(tfg), tfg.shape[0], C.byref(out), _p(bits), _p(npv))
    return out, dict(frame_timing_guess=g, c_est=bits, np_v=npv)


def dedup(cells):
    n = len(cells)
    arr = (Cell * max(n, 1))(*cells)
    out = (Cell * max(n, 1))()
    m = lib().lcso_dedup(arr, n, out)
    return [_copy(out[i]) for i in range(m)]


def cell_search_one(capbuf, f_search_set, fc_requested, fc_programmed, fs_programmed, flags=0, max_cells=64):
    """One centre frequency of CellSearch's main loop (CellSearch.cpp:471-569).  Returns (cells, peaks)."""
    capbuf = np.ascontiguousarray(capbuf, np.complex128)
    f = np.ascontiguousarray(f_search_set, np.float64)
    cells = (Cell * max_cells)(); peaks = (Cell * max_cells)(); npk = C.c_int(0)
    n = lib().lcso_cell_search_one(_p(capbuf), C.c_uint32(capbuf.size), _p(f), f.size, C.c_double(fc_requested),
                                   C.c_double(fc_programmed), C.c_double(fs_programmed), C.c_uint32(flags), cells,
                                   max_cells, peaks, C.byref(npk))
    return [_copy(cells[i]) for i in range(min(n, max_cells))], [_copy(peaks[i]) for i in range(min(npk.value, max_cells))]
