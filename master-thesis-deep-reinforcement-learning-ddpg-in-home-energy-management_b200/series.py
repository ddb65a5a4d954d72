"""Input series for shems_LU1: CSV ingest and the synthetic ChargerID98-shaped generator.

The env reads 8 of the 21 columns of `data/ChargerXX_all_{train,eval,test}_fix.csv`
(shems_LU1.jl:251-260, 268-279; schema from Data_preparation_v2.ipynb cells 8-45).  The
real files are not public, so benchmarks use `synth_charger98`, whose statistics follow
SURVEY.md §8(d) (measured on the Charger98 test series reconstructed from the MPC result
CSV under `SHEMS python/single_building/results/`).

Everything here returns the host layout of the C ABI: float32 [8][nrows] with rows
(soc_ev, h_countdown, electkwh, PV_generation, p_buy, hour_cos, hour_sin, season).
"""
import csv
import math

import numpy as np

COLS = ("soc_ev", "h_countdown", "electkwh", "PV_generation", "p_buy", "hour_cos", "hour_sin", "season")
SPLIT_ROWS = {"train": 4320, "eval": 1440, "test": 3000, "year": 8761}

# Charger98 test-series profile: mean load by hour, monthly load means,
# mean PV by hour (04h..19h), monthly PV means relative to the annual mean.
_LOAD_BY_HOUR = np.array([1.17, 1.41, 1.24, 1.22, 1.45, 2.20, 1.84, 1.31, 0.97, 1.03, 1.25, 1.28, 1.13, 0.97, 1.06,
                          1.13, 1.49, 1.77, 1.69, 1.58, 1.64, 1.55, 1.20, 1.08])
_LOAD_BY_MONTH = np.array([2.64, 2.43, 1.77, 1.65, 1.20, 0.59, 0.57, 0.45, 0.69, 1.05, 1.37, 2.29])
_PV_BY_HOUR = np.zeros(24)
_PV_BY_HOUR[4:20] = [0.05, 0.68, 2.37, 4.53, 6.53, 8.26, 9.59, 10.52, 10.01, 9.15, 7.59, 5.47, 3.88, 2.48, 0.87, 0.14]
_PV_BY_MONTH = np.array([0.30, 1.2, 2.9, 4.6, 6.0, 6.72, 6.4, 5.4, 3.8, 2.1, 0.9, 0.72])


def season_of_month(m):
    """Data_preparation_v2.ipynb cell 17: spring 3-5 -> 1, summer 6-8 -> 2, autumn 9-11 -> 3, else 4."""
    return 1 if 3 <= m <= 5 else 2 if 6 <= m <= 8 else 3 if 9 <= m <= 11 else 4


def synth_charger98(nrows=4320, seed=98, interpolate_soc=True):
    """Seeded synthetic hourly series with Charger98-like load/PV/EV-session statistics.

    h_countdown counts down by 1 to 0 at the last connected hour and is -1 while the EV is
    absent, with >= 1 absent hour between sessions (notebook cell 39); soc_ev = 1 when absent;
    with interpolate_soc the connected rows ramp linearly from the arrival SOC to 1.0 as the
    train split does (cell 40/45); p_buy = 0.4 (cell 34); hour_cos/sin = cos/sin(2π·hour/23)
    (cell 15).  The calendar starts 2020-11-01 00h like the real data.
    """
    rng = np.random.default_rng(seed)
    hours = np.arange(nrows) % 24
    day = np.arange(nrows) // 24
    # 30.4-day months starting in November
    month = ((10 + (day // 30.4375).astype(int)) % 12) + 1
    load = _LOAD_BY_HOUR[hours] * (_LOAD_BY_MONTH[month - 1] / _LOAD_BY_MONTH.mean())
    load = load * rng.lognormal(mean=-0.18, sigma=0.6, size=nrows)
    load = np.clip(load, 0.195, 8.47)
    cloud = np.clip(rng.beta(2.0, 1.6, size=nrows) * 1.45, 0.0, 2.2)
    daily = np.repeat(np.clip(rng.beta(2.5, 1.5, size=nrows // 24 + 1) * 1.5, 0.05, 1.6), 24)[:nrows]
    pv = _PV_BY_HOUR[hours] * (_PV_BY_MONTH[month - 1] / _PV_BY_MONTH.mean()) * cloud * daily
    pv = np.clip(pv * 1.3, 0.0, 22.4)
    pv[pv < 0.02] = 0.0

    cd = -np.ones(nrows)
    soc = np.ones(nrows)
    t = int(rng.integers(1, 30))
    while t < nrows - 2:
        # arrival hour mostly 14-23h or 1-5h
        target_h = int(rng.choice(np.r_[14:24, 1:6]))
        t += (target_h - (t % 24)) % 24
        if t >= nrows - 2:
            break
        u = rng.random()
        length = int(rng.integers(9, 24)) if u < 0.6 else int(rng.integers(39, 72)) if u < 0.9 else int(rng.integers(1, 9))
        length = min(length, nrows - 2 - t)
        if length < 1:
            break
        arrival = float(np.clip(rng.normal(0.45, 0.18), 0.04, 0.78))
        for j in range(length + 1):
            cd[t + j] = length - j
            soc[t + j] = arrival + (1.0 - arrival) * j / length if interpolate_soc else arrival
        t += length + 1
        # gap: >= 1 absent hour, median ~42h, tail to ~390h
        t += 1 + int(min(rng.exponential(55.0), 390))
    out = np.zeros((8, nrows), np.float32)
    out[0] = soc
    out[1] = cd
    out[2] = load
    out[3] = pv
    out[4] = 0.4
    out[5] = np.cos(hours / 23.0 * 2 * math.pi)
    out[6] = np.sin(hours / 23.0 * 2 * math.pi)
    out[7] = [season_of_month(int(m)) for m in month]
    return out


def load_csv(path):
    """Parse a `ChargerXX_all_*_fix.csv` (21-column schema, by header name) once, with the library's native parser
    (shems_series_from_csv; csrc/series.cu).  `load_csv_python` is the independent pure-Python restatement used to test it."""
    import ctypes as C
    from . import _lib as L
    n = C.c_int32()
    L.check(L.lib().shems_series_from_csv(str(path).encode(), None, 0, C.byref(n)))
    out = np.empty((8, n.value), np.float32)
    L.check(L.lib().shems_series_from_csv(str(path).encode(), out.ctypes.data_as(L.PF), n.value, C.byref(n)))
    return out


def load_csv_python(path):
    """Parse a `ChargerXX_all_*_fix.csv` (21-column schema, by header name) once.

    Values are converted Float64 -> Float32 exactly as `env.state.x = df[idx, :col]` does
    (shems_LU1.jl:251-260).  Bool/missing columns are ignored; only the 8 env columns are read.
    """
    with open(path, newline="") as f:
        rd = csv.reader(f)
        header = next(rd)
        pos = [header.index(c) for c in COLS]
        rows = [[float(r[p]) for p in pos] for r in rd if r]
    return np.ascontiguousarray(np.array(rows, dtype=np.float64).T.astype(np.float32))


def from_mpc_results(path, ev_capacity=35.816):
    """Rebuild the env input columns from an MPC result CSV of the reference's Python benchmark.

    Flow balances of SHEMS_optimizer_cost.py:55-57: electkwh = PV_DE + B_DE + GR_DE and
    PV_generation = PV_DE + PV_B + PV_GR + PV_EV; h_countdown = C_EV; soc_ev = Soc_Ev/capacity
    while connected (the MPC trajectory stands in for the data column), 1 when absent.
    """
    with open(path, newline="") as f:
        rd = csv.DictReader(f)
        rows = list(rd)
    n = len(rows)
    out = np.zeros((8, n), np.float64)
    for i, r in enumerate(rows):
        g = lambda k: float(r[k])
        cd = g("C_EV")
        out[0, i] = 1.0 if cd < 0 else min(1.0, g("Soc_Ev") / ev_capacity)
        out[1, i] = cd
        out[2, i] = g("PV_DE") + g("B_DE") + g("GR_DE")
        out[3, i] = g("PV_DE") + g("PV_B") + g("PV_GR") + g("PV_EV")
        out[4, i] = 0.4
        h = g("hour")
        out[5, i] = math.cos(h / 23.0 * 2 * math.pi)
        out[6, i] = math.sin(h / 23.0 * 2 * math.pi)
        out[7, i] = season_of_month(int(g("month")))
    return np.ascontiguousarray(out.astype(np.float32))
