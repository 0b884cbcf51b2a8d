"""BASELINE configs 2 and 5 built with the ORACLE only (oracle/qo100ref.c synthesis + literal specs).

TEST INFRASTRUCTURE ONLY.  bench.py's `--impl reference` arm and `cpu_baseline` leg take their networks
from here, so the reference arm never maps the product library: element values come from
ref_cheby_g / ref_ladder_lpf / ref_add_parasitics, grids from ref_grid_*.
tests/test_oracle.py checks that these bundles equal qo100net.workloads' element for element.

Reference artefacts: pcb/generic-filter/README.md:13 (11th-order ladder board),
util/directional-couplers/dir_cpl_2.4g_20dB.trc:18-20 (coupled-line section), SURVEY.md 8d (specs).
"""
import numpy as np

from . import refbind as R


class RefWorkload:
    def __init__(self, name, elems, rs, rl, f, specs, tols, hist, seed):
        self.name, self.elems, self.rs, self.rl = name, elems, rs, rl
        self.f, self.specs, self.tols, self.hist, self.seed = np.ascontiguousarray(f), list(specs), list(tols), hist, seed


def seed_for(cfg):
    return 0x5EED010000000000 + cfg


def _cheby11(fc):
    """11th-order 0.1 dB Chebyshev, series first, 50 Ohm, with the config-2 parasitic model (Q 60, SRF 30 fc / 0.1 Ohm, SRF 50 fc)."""
    return R.elems_to_list(R.ladder_lpf(R.cheby_g(11, 0.1), fc, 50.0, True, (60.0, 30.0, 0.1, 50.0)))


def _lc_tols(items, tol_l, tol_c, first_var=0):
    out = []
    for i, (kind, _p) in enumerate(items):
        if kind in (R.SER_L, R.SHUNT_L):
            out.append((i, 0, first_var + len(out), R.TOL_REL, tol_l))
        elif kind in (R.SER_C, R.SHUNT_C):
            out.append((i, 0, first_var + len(out), R.TOL_REL, tol_c))
    return out


def cfg2(nf=4096):
    fc = 10e6
    items = _cheby11(fc)
    f = R.grid_log(fc / 2.5, fc * 6.25, nf)
    specs = [(R.SPEC_S21_MIN_DB, 0.0, 0.95 * fc, -2.0), (R.SPEC_S21_MAX_DB, 1.3 * fc, 1e99, -49.0)]
    return RefWorkload("cfg2-cheby11-lpf-1e6x4096", items, 50.0, 50.0, f, specs, _lc_tols(items, 0.05, 0.02),
                       dict(hist_bins=256, hist_spec=0, hist_lo=-4.0, hist_hi=0.0), seed_for(2))


def cfg5(nf=4096):
    items = [(R.CPL_THRU, [55.2771, 45.2267, 95.4225, 95.4225, 2.4e9, 50.0])] + _cheby11(3e9)
    tols = [(0, 0, 0, R.TOL_REL, 0.02), (0, 1, 1, R.TOL_REL, 0.02), (0, 2, 2, R.TOL_REL, 0.01), (0, 3, 2, R.TOL_REL, 0.01)]
    tols += _lc_tols(items, 0.05, 0.02, first_var=3)
    f = R.grid_lin(70e6, 4000e6, nf)
    specs = [(R.SPEC_S21_MIN_DB, 2.3e9, 2.5e9, -1.4), (R.SPEC_S21_MAX_DB, 3.9e9, 1e99, -48.0)]
    k24 = int(np.argmin(np.abs(f - 2.4e9)))
    specs.append((R.SPEC_S21_MIN_DB, f[k24], f[k24], -1e9))
    return RefWorkload("cfg5-coupler+cheby11-1e8x4096", items, 50.0, 50.0, f, specs, tols,
                       dict(hist_bins=256, hist_spec=2, hist_lo=-3.0, hist_hi=0.0), seed_for(5))


def get(name, nf=4096):
    return cfg5(nf) if name.startswith("cfg5") else cfg2(nf)


def run(wl, n_samples, sample_offset=0, nthreads=1):
    """One oracle Monte-Carlo pass over [sample_offset, sample_offset + n_samples) -> counters dict."""
    cfg = R.mc_cfg(wl.seed, n_samples, wl.tols, sample_offset=sample_offset, **wl.hist)
    return R.mc_run(R.make_elems(wl.elems), wl.rs, wl.rl, wl.f, wl.specs, cfg, nthreads=nthreads)
