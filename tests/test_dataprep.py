"""f2: series ingest.  CPU tests of the data-preparation rules restated from RL-SHEMS/Data_preparation_v2.ipynb —
pinned by the notebook's OWN printed outputs where it prints any (reference output that exists in /root/reference) — and of the
native CSV parser of the library (host code, no GPU needed) against the independent pure-Python parser."""
import ctypes as C
import math
import os

import numpy as np
import pytest


def test_d_res_matches_notebook_cell13_output(sb):
    """cells 10 + 13 print the first rows of Charger09's hourly frame and their residual demand `electkwh + chargekwh - PV`;
    the printed values are reproduced (to 1e-12: the frame prints 0.145 for a sum of four quarter-hour readings, so the last
    digits of the notebook's 0.11400000000000002 belong to inputs it does not show)."""
    dp = sb.dataprep
    electkwh = [0.106, 0.125, 0.126, 0.129, 0.129, 0.13, 0.145, 0.579]     # cell 10 output, rows 1-8
    pv = [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.031, 0.226]
    hourly = dict(e_consumption=np.array(electkwh), e_production=np.array(pv), e_charger=np.zeros(8), h_countdown=-np.ones(8), soc_ev=np.ones(8))
    f = dp.add_features(hourly, month=[11] * 8, day=[1] * 8, hour=list(range(8)))
    want = [0.106, 0.125, 0.126, 0.129, 0.129, 0.13, 0.11400000000000002, 0.3529999999999999]   # cell 13 output
    np.testing.assert_allclose(f["d_res"], want, rtol=0, atol=1e-12)
    assert np.isnan(f["chargekwh"]).all()                    # cell 8: `missing` while the EV is absent
    assert f["season"].tolist() == [3] * 8 and f["autumn"].all() and not f["winter"].any()
    assert (f["p_buy"] == 0.4).all() and (f["p_sell"] == 0.08).all()    # cell 34


def test_split_lengths_match_notebook_cell36_output(sb):
    """cell 36 prints train 4320 / eval 1440 / test 3000 for the 8760-hour year.  Without sessions at the split points the
    10 d / 5 d / 15 d pattern gives exactly these lengths; the first split points are the ones cell 35 prints (240, 360, 720, ...)."""
    dp = sb.dataprep
    n = 8760
    hour = np.arange(n) % 24
    frame = dp.add_features(dict(e_consumption=np.ones(n), e_production=np.zeros(n), e_charger=np.zeros(n), h_countdown=-np.ones(n),
                                 soc_ev=np.ones(n)), month=np.ones(n, int), day=np.ones(n, int), hour=hour)
    train, ev, test = dp.split_all_data_advanced_v2(frame)
    assert (len(train["nday"]), len(ev["nday"]), len(test["nday"])) == (4320, 1440, 3000)
    assert test["nday"][0] == 1 and test["nday"][239] == 240 and ev["nday"][0] == 241 and train["nday"][0] == 361 and test["nday"][240] == 721
    # a session across a split point pushes it to the next day and the set gives the day back at its next block
    cd = -np.ones(n)
    cd[230:250] = np.arange(19, -1, -1)                       # connected over row 240 (the first test block's last row)
    frame["h_countdown"] = cd
    train2, ev2, test2 = dp.split_all_data_advanced_v2(frame)
    assert test2["nday"][263] == 264 and ev2["nday"][0] == 265            # block grew by one day
    assert test2["h_countdown"][263] == -1
    assert test2["nday"][264] == 745                                       # next test block starts after eval 120 + train 360
    assert len(test2["nday"]) == 3000 and len(train2["nday"]) == 4320 and len(ev2["nday"]) == 1440


def test_resample_countdown_and_soc_rules(sb):
    """cell 6: energies are summed over the 4 quarter hours, h_countdown takes the group maximum and is floored, soc_ev the group
    minimum; the hour before a repeated countdown value is bumped, a one-hour session gets an arrival row, soc_ev is 1 when absent."""
    dp = sb.dataprep
    q = lambda *hours: np.repeat(np.array(hours, float), 4)
    cons = np.arange(32, dtype=float) / 10
    charger = np.full(32, np.nan); charger[8:20] = 1.5
    #             h0   h1   h2    h3    h4    h5   h6   h7
    cd15 = q(-1, -1, 2.5, 1.75, 0.75, -1, -1, -1)
    cd15[8:12] = [2.5, 2.5, 2.25, 2.0]
    soc15 = q(1, 1, 0.4, 0.5, 0.6, 0.9, 1, 1)
    h = dp.resample_hourly(cons, np.zeros(32), charger, cd15, soc15)
    assert np.allclose(h["e_consumption"], cons.reshape(8, 4).sum(1)) and np.allclose(h["e_charger"], [0, 0, 6, 6, 6, 0, 0, 0])
    assert h["h_countdown"].tolist() == [-1, -1, 2, 1, 0, -1, -1, -1]      # floor of the group maxima
    assert h["soc_ev"].tolist() == [1, 1, 0.4, 0.5, 0.6, 1, 1, 1]          # absent hour with soc 0.9 -> 1
    # repeated value: [.., 1, 1, 0] -> the earlier hour is bumped to 2
    h2 = dp.resample_hourly(np.ones(20), np.zeros(20), np.zeros(20), q(-1, 1.2, 1.0, 0.3, -1), q(1, 0.3, 0.5, 0.8, 1))
    assert h2["h_countdown"].tolist() == [-1, 2, 1, 0, -1]
    # one-hour session: [-1, 0] -> [1, 0] and the arrival row inherits the session's soc
    h3 = dp.resample_hourly(np.ones(12), np.zeros(12), np.zeros(12), q(-1, 0.5, -1), q(1, 0.7, 1))
    assert h3["h_countdown"].tolist() == [1, 0, -1] and h3["soc_ev"].tolist() == [0.7, 0.7, 1]


def test_countdown_fix_and_soc_interpolation(sb):
    """cells 39-40 (and 45: `inserted -1 at: 1741` is the 1-based row the fix-up reports)."""
    dp = sb.dataprep
    f = dict(h_countdown=np.array([-1, 3, 2, 1, 0, 5, 4, -1, 2, 1, 0, -1.0]), soc_ev=np.array([1, .2, .2, .2, .2, .5, .5, 1, .6, .6, .6, 1.0]))
    assert dp.check_and_update_h_countdown(f) == [6]
    assert f["h_countdown"].tolist() == [-1, 3, 2, 1, 0, -1, 4, -1, 2, 1, 0, -1] and f["soc_ev"][5] == 1.0
    dp.interpolate_soc_ev(f)
    assert np.allclose(f["soc_ev"][1:5], [0.2, 0.2 + 0.8 / 3, 0.2 + 1.6 / 3, 1.0])
    assert np.allclose(f["soc_ev"][8:11], [0.6, 0.8, 1.0])
    assert f["soc_ev"][6] == 0.5                               # row 7 is no session start (previous row is not absent... it is: -1) -> starts, never ends
    # time features, cell 15: hour / maximum(hour) = hour / 23
    g = dp.add_features(dict(e_consumption=np.ones(24), e_production=np.zeros(24), e_charger=np.zeros(24), h_countdown=-np.ones(24),
                             soc_ev=np.ones(24)), month=[12] * 24, day=[1] * 24, hour=list(range(24)))
    assert g["hour_cos"][23] == math.cos(2 * math.pi) and abs(g["hour_sin"][6] - math.sin(6 / 23 * 2 * math.pi)) < 1e-15
    assert g["winter"].all() and (g["season"] == 4).all()


def test_native_csv_parser_matches_python_and_handles_schema(sb, tmp_path, train_series):
    """shems_series_from_csv (csrc/series.cu): the 21-column `_fix` schema with Bool columns as true/false and `missing` chargekwh;
    selected by column name, Float64 text -> Float32 like `env.state.x = df[idx, :col]`."""
    dp = sb.dataprep
    n = 300
    ser = train_series[:, :n].astype(np.float64)
    hour = np.arange(n) % 24
    frame = dp.add_features(dict(e_consumption=ser[2], e_production=ser[3], e_charger=np.where(ser[1] >= 0, 1.25, 0.0), h_countdown=ser[1],
                                 soc_ev=ser[0] + 1e-9), month=np.full(n, 11), day=1 + np.arange(n) // 24, hour=hour)
    path = str(tmp_path / "Charger98_all_train_fix.csv")
    dp.write_fix_csv(path, frame)
    head = open(path).readline().strip().split(",")
    assert tuple(head) == dp.CSV_COLUMNS and len(head) == 21
    assert ",true," in open(path).read() and ",," in open(path).read()       # Bool text and an empty (`missing`) field are present
    a = sb.series.load_csv(path)
    b = sb.series.load_csv_python(path)
    assert a.dtype == np.float32 and a.shape == (8, n)
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(a[1], ser[1].astype(np.float32))
    np.testing.assert_array_equal(a[0], (ser[0] + 1e-9).astype(np.float32))  # Float64 text rounded once to Float32
    # CRLF line ends, BOM, quoted header names, shuffled column order
    lines = open(path).read().strip().split("\n")
    order = list(range(21))[::-1]
    shuffled = ["﻿" + ",".join('"%s"' % head[j] for j in order)] + [",".join(l.split(",")[j] for j in order) for l in lines[1:]]
    p2 = str(tmp_path / "shuffled.csv")
    open(p2, "w", newline="").write("\r\n".join(shuffled) + "\r\n")
    np.testing.assert_array_equal(sb.series.load_csv(p2), a)
    # errors: a missing column is a KeyError (Julia: ArgumentError on df[idx, :col]); an unreadable field is reported with its line
    p3 = str(tmp_path / "nocol.csv")
    open(p3, "w").write("\n".join(",".join(x for j, x in enumerate(l.split(",")) if j != 4) for l in lines) + "\n")
    with pytest.raises(sb.ShemsKeyError):
        sb.series.load_csv(p3)
    p4 = str(tmp_path / "bad.csv")
    bad = lines[:]
    cells = bad[5].split(","); cells[3] = "abc"; bad[5] = ",".join(cells)
    open(p4, "w").write("\n".join(bad) + "\n")
    with pytest.raises(sb.ShemsError, match="line 6"):
        sb.series.load_csv(p4)
    with pytest.raises(sb.ShemsError):
        sb.series.load_csv(str(tmp_path / "does_not_exist.csv"))
