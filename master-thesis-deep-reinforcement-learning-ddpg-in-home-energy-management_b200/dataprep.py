"""f2 of SURVEY §8: series ingest — the data-preparation rules of RL-SHEMS/Data_preparation_v2.ipynb restated on numpy columns.

The reference prepares each charger's year of 15-minute measurements in a Julia notebook and writes the three files the
environment reads (`data/ChargerXX_all_{train,eval,test}_fix.csv`, shems_LU1.jl:217, :265).  The raw company data is not public,
so these functions are pinned by what the notebook itself prints (cell outputs, see tests/test_dataprep.py) and by the rules'
own invariants.  A "frame" here is a dict of equally long numpy columns.

    resample_hourly            cell 6   15 min -> 1 h: energies summed, h_countdown max, soc_ev min, countdown fix-ups
    add_features               cells 8-17, 34: month/day/hour, nday, d_res, hour/month cos/sin, season flags, fixed prices
    split_all_data_advanced_v2 cell 26  test 10 d / eval 5 d / train 15 d blocks, split points moved off charging sessions
    check_and_update_h_countdown  cell 39: a departure (0) must be followed by an absent hour (-1)
    interpolate_soc_ev         cell 40  linear SOC ramp from arrival to 1.0 inside each session (train split only, cell 45)
    write_fix_csv / columns    cell 42  the 21-column CSV schema (header order of cell 35's frame)
"""
import math

import numpy as np

CSV_COLUMNS = ("electkwh", "PV_generation", "chargekwh", "h_countdown", "soc_ev", "month", "day", "hour", "nday", "d_res", "hour_cos",
               "hour_sin", "month_cos", "month_sin", "spring", "summer", "autumn", "winter", "season", "p_buy", "p_sell")
SPLIT_LIMITS = {"train": 4320, "eval": 1440, "test": 3000}   # cell 26 build_sets limits; cell 36 prints exactly these lengths


def resample_hourly(e_consumption, e_production, e_charger, h_countdown, soc_ev, per_hour=4):
    """cell 6 `resample`: groups of `per_hour` consecutive samples (timestamps floored to the hour).

    e_charger may contain NaN for `missing` (coalesce(x, 0)).  Returns a frame with hourly e_consumption, e_production,
    e_charger (sums), h_countdown (group maximum, then floor + the two first-value fix-ups), soc_ev (group minimum, 1 when absent).
    """
    n = len(e_consumption) // per_hour
    g = lambda x: np.asarray(x, np.float64)[:n * per_hour].reshape(n, per_hour)
    out = dict(e_consumption=g(e_consumption).sum(1), e_production=g(e_production).sum(1),
               e_charger=np.nan_to_num(g(e_charger), nan=0.0).sum(1), h_countdown=g(h_countdown).max(1), soc_ev=g(soc_ev).min(1))
    cd, soc = out["h_countdown"], out["soc_ev"]
    for i in range(1, n):                                   # Julia 2:nrow
        if cd[i] > -1:
            cd[i] = math.floor(cd[i])
            if cd[i] == cd[i - 1]:
                cd[i - 1] += 1                              # "Change first countdown value: +1"
            elif cd[i] == 0 and cd[i - 1] == -1:
                cd[i - 1] = 1                               # a one-hour session gets its arrival row
                soc[i - 1] = soc[i]
        if cd[i] == -1 and soc[i] < 1:
            soc[i] = 1
    return out


def add_features(hourly, month, day, hour, p_buy=0.4, p_sell=0.08):
    """cells 8-17 and 34: the columns of the `_fix` files from the hourly frame and its calendar columns."""
    month, day, hour = (np.asarray(x, np.int64) for x in (month, day, hour))
    cd = np.asarray(hourly["h_countdown"], np.float64)
    chargekwh = np.asarray(hourly["e_charger"], np.float64).copy()
    chargekwh[cd == -1] = np.nan                             # cell 8: chargekwh is `missing` while the EV is absent
    f = dict(electkwh=np.asarray(hourly["e_consumption"], np.float64), PV_generation=np.asarray(hourly["e_production"], np.float64),
             chargekwh=chargekwh, h_countdown=cd, soc_ev=np.asarray(hourly["soc_ev"], np.float64), month=month, day=day, hour=hour)
    n = len(cd)
    f["nday"] = np.arange(1, n + 1)                          # cell 12 (a row counter despite its name)
    f["d_res"] = f["electkwh"] + np.nan_to_num(chargekwh, nan=0.0) - f["PV_generation"]   # cell 13
    f["hour_cos"] = np.cos(hour / hour.max() * 2 * math.pi)  # cell 15: hour ./ maximum(hour) = hour / 23
    f["hour_sin"] = np.sin(hour / hour.max() * 2 * math.pi)
    f["month_cos"] = np.cos(month / month.max() * 2 * math.pi)
    f["month_sin"] = np.sin(month / month.max() * 2 * math.pi)
    f["spring"] = (month >= 3) & (month <= 5)                # cell 17
    f["summer"] = (month >= 6) & (month <= 8)
    f["autumn"] = (month >= 9) & (month <= 11)
    f["winter"] = (month >= 12) | (month <= 2)
    f["season"] = np.where(f["spring"], 1, np.where(f["summer"], 2, np.where(f["autumn"], 3, 4)))
    f["p_buy"] = np.full(n, p_buy)                           # cell 34
    f["p_sell"] = np.full(n, p_sell)
    return f


def _take(frame, lo, hi):
    return {k: v[lo:hi].copy() for k, v in frame.items()}


def _vcat(a, b):
    return b if a is None else {k: np.concatenate([a[k], b[k]]) for k in a}


def split_all_data_advanced_v2(frame):
    """cell 26: walk the year in (test 10 d, eval 5 d, train 15 d) blocks; a block may not end inside a charging session
    (h_countdown of the block's LAST row must be -1), otherwise it grows by whole days; what a set gained is taken back
    from its next block (at most 4 days); sets are capped at 4320 / 1440 / 3000 rows (build_sets)."""
    cd = frame["h_countdown"]
    nrow = len(cd)
    pattern = [("test", 24 * 10), ("eval", 5 * 24), ("train", 15 * 24)]
    adj = {"train": 0, "eval": 0, "test": 0}
    sets = {"train": None, "eval": None, "test": None}
    pi, i = 0, 0
    while i < nrow:
        name, row_count = pattern[pi]
        take_back = min(adj[name], 4 * 24)
        row_count -= take_back
        adj[name] -= take_back
        while cd[min(i + row_count, nrow) - 1] > -1:        # Julia Input_df[min(i+row_count, nrow), :h_countdown], 1-based
            if i + row_count - 1 > 10000:
                break
            if (i + row_count + 24 - 1) > (nrow - 1):
                break
            row_count += 24
            adj[name] += 24
        rows = _take(frame, i, min(i + row_count, nrow))     # Julia i+1 : min(i+row_count, nrow)
        merged = _vcat(sets[name], rows)
        limit = SPLIT_LIMITS[name]
        length = len(merged["h_countdown"])
        if length > limit:                                   # build_sets: "limit exceeded"
            row_count -= length - limit
            merged = _take(merged, 0, limit)
        sets[name] = merged
        i += row_count
        pi = (pi + 1) % len(pattern)
    return sets["train"], sets["eval"], sets["test"]


def check_and_update_h_countdown(frame):
    """cell 39 (in place): after a departure hour (0) the next hour must be absent: h_countdown = -1, soc_ev = 1."""
    cd, soc = frame["h_countdown"], frame["soc_ev"]
    fixed = []
    for i in range(len(cd) - 1):
        if cd[i] == 0 and cd[i + 1] != -1:
            cd[i + 1] = -1
            soc[i + 1] = 1.0
            fixed.append(i + 2)                              # the notebook prints the 1-based row
    return fixed


def interpolate_soc_ev(frame):
    """cell 40 (in place): inside each session (first row with countdown > 0 after an absent row ... the row with countdown 0)
    soc_ev ramps linearly from the arrival value to 1.0."""
    cd, soc = frame["h_countdown"], frame["soc_ev"]
    start = None
    for i in range(len(cd)):
        if cd[i] > 0 and (i == 0 or cd[i - 1] == -1):
            start = i
        if cd[i] == 0:
            end = i
            if start is not None:
                s0 = soc[start]
                for j in range(start, end + 1):
                    soc[j] = s0 + (1.0 - s0) * (j - start) / (end - start)
                start = None


def write_fix_csv(path, frame):
    """cell 42: CSV.write of the 21-column frame (Bool columns as true/false, `missing` as an empty field)."""
    def fmt(k, v):
        if k in ("spring", "summer", "autumn", "winter"):
            return "true" if v else "false"
        if k in ("month", "day", "hour", "nday", "season"):
            return str(int(v))
        if isinstance(v, float) and math.isnan(v):
            return ""
        return repr(float(v))
    n = len(frame["h_countdown"])
    with open(path, "w") as f:
        f.write(",".join(CSV_COLUMNS) + "\n")
        for i in range(n):
            f.write(",".join(fmt(k, frame[k][i]) for k in CSV_COLUMNS) + "\n")


def prepare_charger(e_consumption, e_production, e_charger, h_countdown, soc_ev, month, day, hour, interpolate_train=True):
    """The notebook end to end for one charger: 15-minute columns (+ hourly calendar columns) -> (train, eval, test) frames."""
    hourly = resample_hourly(e_consumption, e_production, e_charger, h_countdown, soc_ev)
    frame = add_features(hourly, month, day, hour)
    train, ev, test = split_all_data_advanced_v2(frame)
    for part in (train, test, ev):
        check_and_update_h_countdown(part)
    if interpolate_train:
        interpolate_soc_ev(train)
    return train, ev, test
