"""The BASELINE.json configurations as (network, grid, specs, tolerances, histogram) bundles
(SURVEY.md §8d).  Element values of the rf-tools filters are the ones printed in the reference
SVGs (util/if-bandpass-filter/schematic.svg:191-213, util/gpsdo-ouput-filters/10M/
schematic.svg:197-231, docs/gpsdo-filters/*.svg:195-241); when the reference tree is mounted
they can equally be loaded with Net.from_rftools_svg -- tests check both give the same list.
"""
import numpy as np

from . import (CPL_MS, CPL_THRU, MCORN, MLIN, MOPEN, MTEE, SER_L, SER_LC_SER, SHUNT_C, SHUNT_LC_PAR, SHUNT_LC_SER,
               SPEC_S21_MAX_DB, SPEC_S21_MIN_DB, SUBST, TOL_ABS, TOL_REL, Net, grid_lin, grid_log, lc_tolerances)


def seed_for(cfg):
    return 0x5EED010000000000 + cfg


class Workload:
    def __init__(self, name, net, f, specs=(), tols=(), hist=None, n_samples=0, seed=0):
        self.name, self.net, self.f = name, net, np.ascontiguousarray(f)
        self.specs, self.tols = list(specs), list(tols)
        self.hist = hist or dict(hist_bins=0, hist_spec=0, hist_lo=0.0, hist_hi=1.0)
        self.n_samples, self.seed = n_samples, seed


def if_bpf_net():
    return Net.from_elements([(SER_LC_SER, [33e-9, 4.7e-12]), (SHUNT_LC_PAR, [4.7e-9, 33e-12]),
                              (SER_LC_SER, [33e-9, 4.7e-12])], 50.0, 50.0)


def cfg1():
    """util/if-bandpass-filter nominal, 1024-point log grid 300e6/3.5 .. 500e6*3.5."""
    return Workload("cfg1-if-bpf-nominal", if_bpf_net(), grid_log(300e6 / 3.5, 500e6 * 3.5, 1024), seed=seed_for(1))


def cheby11(fc):
    return Net.cheby_lpf(11, 0.1, fc, 50.0, True).add_parasitics(fc, 60.0, 30.0, 0.1, 50.0)


def cfg2(n_samples=1000000, nf=4096):
    """pcb/generic-filter 11th-order 0.1 dB Chebyshev LPF, fc 10 MHz, ESR/SRF parasitics, +-5 % L, +-2 % C."""
    fc = 10e6
    net = cheby11(fc)
    f = grid_log(fc / 2.5, fc * 6.25, nf)
    specs = [(SPEC_S21_MIN_DB, 0.0, 0.95 * fc, -2.0), (SPEC_S21_MAX_DB, 1.3 * fc, 1e99, -49.0)]
    return Workload("cfg2-cheby11-lpf-1e6x4096", net, f, specs, lc_tolerances(net, 0.05, 0.02),
                    dict(hist_bins=256, hist_spec=0, hist_lo=-4.0, hist_hi=0.0), n_samples, seed_for(2))


def pa_lpf_net():
    """The cascade of util/pa-lpf-simulation/pa-lpf-simulation.sch."""
    zw = 0.75e-3

    def ml(length, w=zw):
        return (MLIN, [w, float("%se-3" % repr(length))])      # the correctly rounded decimal, as the loader reads it
    co = (MCORN, [zw])
    items = [(SUBST, [4.5, 0.6e-3, 34.79e-6, 0.045, 1.68e-8, 0.15e-6]),
             ml(1.65), co, ml(1.40415), co, ml(1.65), co, ml(1.40415), co, ml(0.95),
             (MTEE, [zw, zw, zw]), ml(0.15), ml(4.35, 3e-3), (MOPEN, [3e-3]),
             ml(0.95), co, ml(2.5), co, ml(3.45), co, ml(1.55), co, ml(0.95),
             (MTEE, [zw, zw, zw]), ml(2.5), ml(3.0, 3e-3), (MOPEN, [3e-3]),
             ml(0.95), co, ml(0.8), co, ml(1.6), co, ml(1.75), co, ml(5.95)]
    return Net.from_elements(items, 50.0, 50.0)


def cfg3(n_samples=10000000):
    """PA microstrip LPF harmonic-rejection yield at 2.4/4.8/7.2 GHz; one draw per board:
    er +-0.2 abs, h +-10 %, etch delta +-0.05 mm abs on every width, t +-20 %."""
    net = pa_lpf_net()
    tols = [(0, 0, 0, TOL_ABS, 0.2), (0, 1, 1, TOL_REL, 0.10), (0, 2, 3, TOL_REL, 0.20)]
    for i, (k, _p) in enumerate(net.elements):
        if k in (MLIN, MCORN, MOPEN):
            tols.append((i, 0, 2, TOL_ABS, 0.05e-3))
        elif k == MTEE:
            tols += [(i, 0, 2, TOL_ABS, 0.05e-3), (i, 1, 2, TOL_ABS, 0.05e-3), (i, 2, 2, TOL_ABS, 0.05e-3)]
    f = np.array([2.4e9, 4.8e9, 7.2e9])
    specs = [(SPEC_S21_MIN_DB, 2.3e9, 2.5e9, -1.0), (SPEC_S21_MAX_DB, 4.7e9, 4.9e9, -22.0),
             (SPEC_S21_MAX_DB, 7.1e9, 7.3e9, -8.5)]
    return Workload("cfg3-pa-lpf-microstrip-yield", net, f, specs, tols,
                    dict(hist_bins=64, hist_spec=1, hist_lo=-30.0, hist_hi=-15.0), n_samples, seed_for(3))


def cfg3b(n_samples=10000000):
    """Config 3 as BASELINE.json words it ("PA LPF with ESR/SRF parasitics, harmonic-rejection yield at 2.4/4.8/7.2 GHz"): the
    reference tree holds the PA low-pass only as the microstrip layout of cfg3(); this is its lumped twin -- a
    7th-order 0.1 dB Chebyshev LC low-pass, fc = 2.9 GHz, with the config-2 parasitic model, +-5 % L / +-2 % C, evaluated at the
    carrier and its two harmonics only (3 points per sample: per-sample work dominates)."""
    fc = 2.9e9
    net = Net.cheby_lpf(7, 0.1, fc, 50.0, True).add_parasitics(fc, 60.0, 30.0, 0.1, 50.0)
    f = np.array([2.4e9, 4.8e9, 7.2e9])
    specs = [(SPEC_S21_MIN_DB, 2.3e9, 2.5e9, -0.70), (SPEC_S21_MAX_DB, 4.7e9, 4.9e9, -43.0), (SPEC_S21_MAX_DB, 7.1e9, 7.3e9, -71.5)]
    return Workload("cfg3b-pa-lpf-lumped-twin-yield", net, f, specs, lc_tolerances(net, 0.05, 0.02),
                    dict(hist_bins=64, hist_spec=1, hist_lo=-50.0, hist_hi=-38.0), n_samples, seed_for(3))


def gpsdo_bank():
    """(name, net, fc): 10M Chebyshev 100/50 Ohm + 15M/40M/60M elliptic 50/50 Ohm."""
    def ell(vals):
        l1, c2, l2, l3, c4, l4, l5, c6, l6, l7 = vals
        return Net.from_elements([(SER_L, [l1]), (SHUNT_LC_SER, [l2, c2]), (SER_L, [l3]), (SHUNT_LC_SER, [l4, c4]),
                                  (SER_L, [l5]), (SHUNT_LC_SER, [l6, c6]), (SER_L, [l7])], 50.0, 50.0)
    n10 = Net.from_elements([(SHUNT_C, [430e-12]), (SER_L, [1.3e-6]), (SHUNT_C, [620e-12]), (SER_L, [1.3e-6]),
                             (SHUNT_C, [560e-12]), (SER_L, [1.1e-6]), (SHUNT_C, [240e-12])], 100.0, 50.0)
    n15 = ell([560e-9, 270e-12, 82e-9, 820e-9, 180e-12, 470e-9, 680e-9, 180e-12, 330e-9, 390e-9])
    n40 = ell([220e-9, 100e-12, 33e-9, 270e-9, 68e-12, 180e-9, 270e-9, 68e-12, 120e-9, 150e-9])
    n60 = ell([150e-9, 68e-12, 22e-9, 180e-9, 47e-12, 120e-9, 180e-9, 47e-12, 82e-9, 100e-9])
    return [("10M", n10, 10e6), ("15M", n15, 15e6), ("40M", n40, 40e6), ("60M", n60, 60e6)]


def cfg4(n_samples=65536, nf=4096):
    """GPSDO output-filter bank, FULL_S (HBM-write-bound) mode; one Workload per filter."""
    out = []
    for name, net, fc in gpsdo_bank():
        out.append(Workload("cfg4-gpsdo-%s-full-s" % name, net, grid_log(fc / 2.5, fc * 6.25, nf), [],
                            lc_tolerances(net, 0.05, 0.05), None, n_samples, seed_for(4)))
    return out


def cfg5(n_samples=100000000, nf=4096):
    """Coupled-line through section (dir_cpl_2.4g_20dB.trc:18-20) + the cfg-2 ladder rescaled to fc = 3 GHz;
    ladder +-5 % L / +-2 % C, coupler Z0e, Z0o +-2 %, theta +-1 % (one draw for both mode angles)."""
    cpl = Net.from_elements([(CPL_THRU, [55.2771, 45.2267, 95.4225, 95.4225, 2.4e9, 50.0])], 50.0, 50.0)
    net = cpl.concat(cheby11(3e9))
    tols = [(0, 0, 0, TOL_REL, 0.02), (0, 1, 1, TOL_REL, 0.02), (0, 2, 2, TOL_REL, 0.01), (0, 3, 2, TOL_REL, 0.01)]
    tols += [(e, p, v + 3, m, t) for (e, p, v, m, t) in lc_tolerances(net, 0.05, 0.02)]
    f = grid_lin(70e6, 4000e6, nf)
    specs = [(SPEC_S21_MIN_DB, 2.3e9, 2.5e9, -1.4), (SPEC_S21_MAX_DB, 3.9e9, 1e99, -48.0)]
    k24 = int(np.argmin(np.abs(f - 2.4e9)))
    specs.append((SPEC_S21_MIN_DB, f[k24], f[k24], -1e9))     # histogram variable: |S21| dB at the grid point nearest 2.4 GHz
    return Workload("cfg5-coupler+cheby11-1e8x4096", net, f, specs, tols,
                    dict(hist_bins=256, hist_spec=2, hist_lo=-3.0, hist_hi=0.0), n_samples, seed_for(5))


def cfg5p(n_samples=100000000, nf=4096):
    """Config 5 with the coupler described PHYSICALLY: the 2.4 GHz 20 dB coupled microstrip of
    util/directional-couplers/dir_cpl_2.4g_20dB.trc:6-17 (W 1.69218 mm, S 0.991476 mm, L 20 mm on Er 3.5, H 0.762 mm,
    T 35 um, cover 20 cm, analysed at 2.4 GHz) with manufacturing tolerances -- etch +-0.03 mm on W (and -/+ on S: one
    draw widens the strips and narrows the gap), H +-5 %, Er +-0.1 -- in front of the cfg-2 ladder at fc = 3 GHz."""
    sub = (SUBST, [3.5, 0.762e-3, 35e-6, 0.0, 0.0, 0.0])
    cpl = (CPL_MS, [1.69218e-3, 0.991476e-3, 20e-3, 0.2, 2.4e9, 50.0])
    lad = cheby11(3e9)
    net = Net.from_elements([sub, cpl] + lad.elements, 50.0, 50.0)
    tols = [(1, 0, 0, TOL_ABS, 0.03e-3), (1, 1, 0, TOL_ABS, -0.03e-3), (0, 1, 1, TOL_REL, 0.05), (0, 0, 2, TOL_ABS, 0.1)]
    tols += [(e + 2, p, v + 3, m, t) for (e, p, v, m, t) in lc_tolerances(lad, 0.05, 0.02)]
    f = grid_lin(70e6, 4000e6, nf)
    specs = [(SPEC_S21_MIN_DB, 2.3e9, 2.5e9, -1.4), (SPEC_S21_MAX_DB, 3.9e9, 1e99, -48.0)]
    k24 = int(np.argmin(np.abs(f - 2.4e9)))
    specs.append((SPEC_S21_MIN_DB, f[k24], f[k24], -1e9))
    return Workload("cfg5p-physical-coupler+cheby11", net, f, specs, tols,
                    dict(hist_bins=256, hist_spec=2, hist_lo=-3.0, hist_hi=0.0), n_samples, seed_for(5))


def pa_bias_netlist():
    """The reference's 5-port bias network, util/pa-bias-simulation/pa-bias-simulation.sch:19-72, read by hand (component line
    numbers in comments) -> (branches, node count, ports).  Nodes:
    1 P1/R17/VCVS in+ | 2 VCVS SRC3 out+ | 3 R16-C12 | 4 rail (610,670) | 5 SPfile far side (890,400) | 6 C9-R1 |
    7 P3 | 8 C2-R2 | 9 C3-R3 | 10 R9-C1 | 11 P2 | 12 C4-R4 ; second sub-circuit: 13 P6(port 4)/R19/VCVS in+ |
    14 SRC4 out+ | 15 R18-C13 | 16 rail | 17 C6-R6 | 18 C7-R7 | 19 R10-C5 | 20 P5 | 21 C8-R8."""
    R, C, V, S = 1, 3, 4, 5                  # QO_NB_R, QO_NB_C, QO_NB_VCVS, QO_NB_SBLOCK
    br = [
        (R, [1, 0], [50.0]),                 # R17 :48
        (V, [1, 2, 0, 0], [1.0, 0.0]),       # SRC3 :40
        (R, [2, 3], [2.6]),                  # R16 :41
        (C, [3, 4], [26.5e-12, 0, 0]),       # C12 :42
        (S, [4, 5, 0], [0, 1, 50.0]),        # L_11SQ39N :39
        (C, [5, 6], [100e-6, 0, 0]),         # C9 :24
        (R, [6, 0], [10.0]),                 # R1 :35
        (R, [5, 7], [1000.0]),               # R11 :36
        (C, [4, 8], [2.2e-12, 0, 0]),        # C2 :28
        (R, [8, 0], [3.0]),                  # R2 :32
        (C, [4, 9], [1.8e-12, 0, 0]),        # C3 :29
        (R, [9, 0], [3.7]),                  # R3 :33
        (R, [4, 10], [0.6]),                 # R9 :34
        (C, [10, 11], [12e-12, 0, 0]),       # C1 :19
        (C, [11, 12], [1.2e-12, 0, 0]),      # C4 :20
        (R, [12, 0], [5.5]),                 # R4 :27
        (R, [13, 0], [50.0]),                # R19 :67
        (V, [13, 14, 0, 0], [1.0, 0.0]),     # SRC4 :59
        (R, [14, 15], [2.6]),                # R18 :60
        (C, [15, 16], [26.5e-12, 0, 0]),     # C13 :61
        (C, [16, 17], [2.2e-12, 0, 0]),      # C6 :53
        (R, [17, 0], [3.0]),                 # R6 :57
        (C, [16, 18], [1.8e-12, 0, 0]),      # C7 :54
        (R, [18, 0], [3.7]),                 # R7 :58 (sic: R7 3.7)
        (R, [16, 19], [0.6]),                # R10 :56
        (C, [19, 20], [12e-12, 0, 0]),       # C5 :49
        (C, [20, 21], [1.2e-12, 0, 0]),      # C8 :50
        (R, [21, 0], [5.5]),                 # R8 :52
    ]
    ports = [(1, 50.0), (11, 50.0), (7, 50.0), (13, 50.0), (20, 50.0)]     # Pac numbers 1..5 (P6 carries number 4)
    return br, 21, ports


def pa_bias_nodal(Q, block_f, block_s, z0=50.0):
    """pa_bias_netlist() as a Nodal object; the SPfile inductor (pa-bias-simulation.sch:39) from its measured points
    (block_s[:, 0..3] = S11, S21, S12, S22).  -> (nodal, branches, tolerances: every R +-1 %, every C +-5 %)."""
    br, nn, ports = pa_bias_netlist()
    nd = Q.Nodal(nn)
    idx = nd.add_sblock(Q.SBlock.from_arrays(block_f, block_s[:, 0], block_s[:, 1], block_s[:, 2], block_s[:, 3], z0))
    for kind, nodes, p in br:
        nd.add_branch(kind, nodes, [idx, p[1], p[2]] if kind == 5 else p)
    for node, z in ports:
        nd.add_port(node, z)
    tols = [(i, 0, v, TOL_REL, 0.05 if b[0] == 3 else 0.01) for v, (i, b) in enumerate((i, b) for i, b in enumerate(br) if b[0] in (1, 3))]
    return nd, br, tols
