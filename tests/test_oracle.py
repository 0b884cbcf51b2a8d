"""The oracle against every golden vector the reference tree holds for this path
 -- CPU only.  If these fail the oracle cannot be trusted as a checker."""
import numpy as np
import pytest

from conftest import relerr


def _pa_lpf(R, golden_nets):
    n = golden_nets["util/pa-lpf-simulation/pa-lpf-simulation.sch"]
    return R.make_elems([(k, p) for k, p in n["elements"]]), n["rs"], n["rl"]


def test_grid_lin_bit_exact_vs_dat(R, golden_dat):
    # pa-lpf-simulation.sch:59 (.SP lin 10 MHz..10 GHz, 5000) vs .dat:6-5005
    assert np.array_equal(R.grid_lin(1e7, 1e10, 5000), golden_dat["frequency"])


def test_pa_lpf_dat_parity(R, golden_dat, golden_nets):
    # util/pa-lpf-simulation/pa-lpf-simulation.dat:5007-35017, criteria of SURVEY §8c
    e, rs, rl = _pa_lpf(R, golden_nets)
    s11, s21, s12, s22 = R.sweep(e, rs, rl, golden_dat["frequency"])
    assert relerr(s21, golden_dat["S21"]) <= 1e-9
    assert relerr(s12, golden_dat["S12"]) <= 1e-9
    for got, ref in ((s11, golden_dat["S11"]), (s22, golden_dat["S22"])):
        assert np.all(np.abs(got - ref) <= 1e-9 * np.maximum(np.abs(ref), 0.02))
    assert np.max(np.abs(20 * np.log10(np.abs(s21)) - golden_dat["S21_dB"])) <= 1e-9
    assert np.max(np.abs(20 * np.log10(np.abs(s11)) - golden_dat["S11_dB"])) <= 1e-6
    # tighter: what the restatement actually achieves (guards against silent model drift)
    assert relerr(s21, golden_dat["S21"]) <= 1e-11


def test_pa_lpf_dpl_markers(R, golden_nets):
    # pa-lpf-simulation.dpl:25-28 -- markers read to 3 digits on neighbouring grid points
    e, rs, rl = _pa_lpf(R, golden_nets)
    f = R.grid_lin(1e7, 1e10, 5000)
    s21db = 20 * np.log10(np.abs(R.sweep(e, rs, rl, f)[1]))
    for fm, val in ((2.39009e9, -0.952), (7.20424e9, -8.85), (4.79817e9, -22.8), (3.06355e9, -2.82)):
        k = int(np.argmin(np.abs(f - fm)))
        assert abs(s21db[k] - val) < 0.006 * max(1.0, abs(val) / 3)


def test_quasi_static_spot_values(R):
    Z, E, Weff = R.ms_quasi(0.75e-3, 0.6e-3, 34.79e-6, 4.5)
    assert abs(Z - 61.27242653861835) < 1e-12 and abs(E - 3.214058807174604) < 1e-13
    assert abs(Weff - 0.7876421546866908e-3) < 1e-17


def test_philox_known_answers(R):
    # SURVEY App. C (Salmon et al. SC'11 Random123 kat vectors)
    assert R.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert R.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert R.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_stream_contract(R):
    seed, s = 0x0123456789abcdef, (5 << 32) | 7
    o = R.philox([7, 5, 3, 0], [0x89abcdef, 0x01234567])
    u6 = ((o[1] << 32 | o[0]) >> 11) * 2.0 ** -53
    u7 = ((o[3] << 32 | o[2]) >> 11) * 2.0 ** -53
    assert R.lib().ref_uniform01(seed, s, 6) == u6 and R.lib().ref_uniform01(seed, s, 7) == u7
    assert R.lib().ref_perturb_factor(seed, s, 6, 0, 0.05) == 1.0 + 0.05 * (2 * u6 - 1) or \
        abs(R.lib().ref_perturb_factor(seed, s, 6, 0, 0.05) - (1.0 + 0.05 * (2 * u6 - 1))) < 3e-16


def test_norminv_and_log(R):
    from scipy.stats import norm
    p = np.concatenate([np.linspace(1e-6, 1 - 1e-6, 2001), [0.02425, 0.97575, 0.5, 1e-12]])
    z = np.array([R.lib().ref_norminv(float(x)) for x in p])
    assert np.max(np.abs(z - norm.ppf(p)) / np.maximum(1.0, np.abs(norm.ppf(p)))) < 2e-9
    x = 10.0 ** np.linspace(-300, 300, 4001)
    lg = np.array([R.lib().ref_log_det(float(v)) for v in x])
    assert np.max(np.abs(lg - np.log(x)) / np.maximum(1.0, np.abs(np.log(x)))) < 4e-16
    g = np.array([R.lib().ref_variate(11, i, 3, 1) for i in range(20000)])
    assert np.all(np.abs(g) <= 1.0) and abs(g.std() - 1 / 3) < 0.01 and abs(g.mean()) < 0.01


@pytest.mark.parametrize("case", ["if_bpf", "gpsdo_10m", "gpsdo_15m", "gpsdo_40m", "gpsdo_60m", "lol_hpf",
                                  "cheby11_ideal", "cfg2_nominal", "coupler_20db", "cfg5_nominal"])
def test_appendix_b_goldens(R, golden_b, case):
    # 40-digit mpmath values of the ladder / coupled-line equations (tools/make_golden.py)
    c = golden_b[case]
    e = R.make_elems([(k, p) for k, p in c["elements"]])
    f = np.array([r["f"] for r in c["rows"]])
    s11, s21, _, _ = R.sweep(e, c["rs"], c["rl"], f)
    ref21 = np.array([complex(float(r["s21"][0]), float(r["s21"][1])) for r in c["rows"]])
    ref11 = np.array([complex(float(r["s11"][0]), float(r["s11"][1])) for r in c["rows"]])
    assert relerr(s21, ref21) < 2e-11      # conditioning of an 11-element chain in double
    assert np.max(np.abs(s11 - ref11)) < 1e-11


def test_synthesis(R, golden_b):
    g = R.cheby_g(11, 0.1)
    assert np.max(np.abs(g - np.array([float(x) for x in golden_b["cheby11_g"]]))) < 2e-15
    assert np.max(np.abs(R.butter_g(11) - np.array([float(x) for x in golden_b["butter11_g"]]))) < 1e-15
    g7 = R.cheby_g(7, 0.1)      # textbook table (Matthaei, Young, Jones)
    assert np.allclose(g7[:4], [1.1811783, 1.4228062, 2.0966713, 1.5734011], atol=2e-7)
    lad = R.ladder_lpf(g, 10e6, 50.0, True, (60, 30, 0.1, 50))
    el = R.elems_to_list(lad)
    assert abs(el[0][1][0] - 957.38778813045415e-9) < 1e-21 and abs(el[0][1][1] - 1.0025741472757387) < 1e-14
    assert abs(el[0][1][2] - 0.29397464207207637e-12) < 1e-27 and abs(el[1][1][2] - 0.21917752199026542e-9) < 1e-24


def test_trc_analysis(R, golden_nets):
    # util/directional-couplers/*.trc:6-20: physical -> (Z0e, Z0o, Ang_l) to the files' 6 digits
    for key, t in golden_nets.items():
        if not key.endswith(".trc"):
            continue
        ze, zo, ae, ao = R.cpl_analyze(t["w"], t["s"], t["h"], t["t"], t["er"], t["ht"], t["f0"], t["l"])
        assert abs(ze / t["z0e"] - 1) < 5e-6 and abs(zo / t["z0o"] - 1) < 5e-6
        assert abs(np.sqrt(ae * ao) / t["ang"] - 1) < 5e-6


def test_mc_oracle_properties(R):
    g = R.cheby_g(11, 0.1)
    lad = R.ladder_lpf(g, 10e6, 50.0, True, (60, 30, 0.1, 50))
    f = R.grid_log(4e6, 62.5e6, 257)
    tols = [(i, 0, i, 0, 0.05 if i % 2 == 0 else 0.02) for i in range(11)]
    specs = [(1, 0, 9.5e6, -2.0), (2, 13e6, 1e99, -49.0)]
    a = R.mc_run(lad, 50, 50, f, specs, R.mc_cfg(1, 300, tols, hist_bins=16, hist_spec=0, hist_lo=-4, hist_hi=0))
    b = R.mc_run(lad, 50, 50, f, specs, R.mc_cfg(1, 300, tols, hist_bins=16, hist_spec=0, hist_lo=-4, hist_hi=0), nthreads=4)
    assert a["n_total"] == 300 and a["n_pass"] == b["n_pass"] and np.array_equal(a["hist"], b["hist"])
    assert int(a["hist"].sum()) == 300 and 0 < a["n_pass"] < 300
    # sharding invariance: [0,300) == [0,120) + [120,300)
    c = R.mc_run(lad, 50, 50, f, specs, R.mc_cfg(1, 120, tols, hist_bins=16, hist_spec=0, hist_lo=-4, hist_hi=0))
    d = R.mc_run(lad, 50, 50, f, specs, R.mc_cfg(1, 180, tols, sample_offset=120, hist_bins=16, hist_spec=0, hist_lo=-4, hist_hi=0))
    assert c["n_pass"] + d["n_pass"] == a["n_pass"] and np.array_equal(c["hist"] + d["hist"], a["hist"])
    # zero tolerance => every sample is the nominal design
    z = R.mc_run(lad, 50, 50, f, specs, R.mc_cfg(1, 10, [(0, 0, 0, 0, 0.0)]))
    assert z["n_pass"] in (0, 10)


def test_oracle_line_inside_and_measured_block_behind_vs_numpy(R, golden_s2p):
    """The cascade family of the compiled chain kernel's GPU test (a ladder with a transmission line INSIDE it, the measured inductor
    of util/pa-bias-simulation/pa-bias-simulation.sch:39 and a resistor BEHIND it, unequal terminations), oracle against an
    independent numpy restatement (plain 2x2 products, SPfile interpolation and S -> ABCD from conftest): the checker of that GPU
    test is itself pinned on this topology.  Lossy L: R + jwL in parallel with Cp; lossy C: ESR + 1/(jwC) + jw Ls."""
    from conftest import np_s_to_abcd, np_spfile
    fd, sd, z0 = golden_s2p["11SQ39N_f"], golden_s2p["11SQ39N_s"], float(golden_s2p["11SQ39N_z0"])
    R.sblock_clear()
    R.sblock_register(0, fd, sd[:, 0], sd[:, 1], sd[:, 2], sd[:, 3], z0)
    SER_R, SER_L, SHUNT_C, TLINE, SBLOCK = R.SER_R, R.SER_L, R.SHUNT_C, R.TLINE, 18          # include/qo100net.h: QO_SBLOCK = 18
    L1, R1, Cp1 = 1.2e-6, 0.9, 0.4e-12
    C1, E1, Ls1 = 330e-12, 0.12, 0.7e-9
    L2, C2 = 0.8e-6, 270e-12
    items = [(SER_L, [L1, R1, Cp1]), (SHUNT_C, [C1, E1, Ls1]), (TLINE, [75.0, 20.0, 10e6]), (SER_L, [L2, 0.0, 0.0]),
             (SHUNT_C, [C2, 0.0, 0.0]), (SBLOCK, [0.0, 1.0]), (SER_R, [2.2])]
    f = np.geomspace(2e6, 3e8, 400)
    rs, rl = 50.0, 75.0
    o = R.sweep(R.make_elems(items), rs, rl, f)
    w = 2 * np.pi * f
    one, zero = np.ones_like(w, dtype=complex), np.zeros_like(w, dtype=complex)

    def ser(z): return np.stack([one, z, zero, one], 1)
    def sh(y): return np.stack([one, zero, y, one], 1)
    def mul(a, b): return np.stack([a[:, 0] * b[:, 0] + a[:, 1] * b[:, 2], a[:, 0] * b[:, 1] + a[:, 1] * b[:, 3],
                                    a[:, 2] * b[:, 0] + a[:, 3] * b[:, 2], a[:, 2] * b[:, 1] + a[:, 3] * b[:, 3]], 1)
    th = np.deg2rad(20.0) * f / 10e6
    zl1 = 1.0 / (1.0 / (R1 + 1j * w * L1) + 1j * w * Cp1)
    parts = [ser(zl1), sh(1.0 / (E1 + 1.0 / (1j * w * C1) + 1j * w * Ls1)),
             np.stack([np.cos(th) + 0j, 1j * 75.0 * np.sin(th), 1j * np.sin(th) / 75.0, np.cos(th) + 0j], 1),
             ser(1j * w * L2), sh(1j * w * C2), np_s_to_abcd(np_spfile(f, fd, sd, True), z0), ser(np.full_like(w, 2.2, dtype=complex))]
    M = parts[0]
    for p_ in parts[1:]:
        M = mul(M, p_)
    den = M[:, 0] * rl + M[:, 1] + M[:, 2] * rs * rl + M[:, 3] * rs
    s21 = 2 * np.sqrt(rs * rl) / den
    s11 = (M[:, 0] * rl + M[:, 1] - M[:, 2] * rs * rl - M[:, 3] * rs) / den
    s22 = (-M[:, 0] * rl + M[:, 1] - M[:, 2] * rs * rl + M[:, 3] * rs) / den
    assert relerr(o[1], s21) < 1e-10 and relerr(o[2], s21) < 1e-10           # the measured inductor is reciprocal to its print precision
    assert np.max(np.abs(o[0] - s11)) < 1e-10 and np.max(np.abs(o[3] - s22)) < 1e-10
    assert 20 * np.log10(np.abs(s21)).min() < -30 and 20 * np.log10(np.abs(s21)).max() > -3      # a real low-pass: both regimes are exercised
    R.sblock_clear()
