"""Writes the INPUT files of julia/make_golden.jl under tests/golden/julia_inputs/ (committed, so the Julia side needs nothing
but the reference checkout and this directory):

  Charger98_all_test_fix.csv  the Charger98 test series (tests/golden/charger98_test_series.npz) in the 21-column schema the
                              reference reads with CSV.read (Data_preparation_v2.ipynb cell 35); only the 8 columns the env reads
                              carry data, the others are fillers
  lu1_cases.csv               seeded (state, idx, action, track) cases reaching every leaf of step! (tests/lu1_cases.py)
  lu1_tape.csv                72 x 2 target actions in [0,1] for a closed-loop DRL episode (track = 1)
(the replay() problem of make_golden.jl needs no input file: weights and memory come from a counter-based generator written
out on both sides, tests/julia_golden_spec.py <-> julia/make_golden.jl)

Every number is written as the shortest decimal that round-trips the Float64 value of the float32 (so CSV.jl's Float64 parse
followed by the reference's own Float32 conversion reproduces the float32 exactly).  Run in the build container:
    python tests/golden/make_julia_inputs.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import lu1_cases as Cs  # noqa: E402

OUT = os.path.join(HERE, "julia_inputs")
N_CASES = 10000
HEADER21 = ["electkwh", "PV_generation", "chargekwh", "h_countdown", "soc_ev", "month", "day", "hour", "nday", "d_res", "hour_cos",
            "hour_sin", "month_cos", "month_sin", "spring", "summer", "autumn", "winter", "season", "p_buy", "p_sell"]


def num(x):
    return repr(float(x))


def write_series(path, ser):
    soc, cd, load, pv, pb, hc, hs, season = ser
    with open(path, "w") as f:
        f.write(",".join(HEADER21) + "\n")
        for i in range(ser.shape[1]):
            s = int(season[i])
            row = [num(load[i]), num(pv[i]), "0.0", str(int(cd[i])), num(soc[i]), "1", "1", str(i % 24), "1", num(load[i] - pv[i]),
                   num(hc[i]), num(hs[i]), "0.0", "0.0", "true" if s == 1 else "false", "true" if s == 2 else "false",
                   "true" if s == 3 else "false", "true" if s == 4 else "false", str(s), num(pb[i]), "0.08"]
            f.write(",".join(row) + "\n")


def main():
    os.makedirs(OUT, exist_ok=True)
    ser = np.load(os.path.join(HERE, "charger98_test_series.npz"))["series"]
    write_series(os.path.join(OUT, "Charger98_all_test_fix.csv"), ser)
    cs = Cs.make_cases(ser, N_CASES)
    names = "Soc_b Soc_ev c_ev d_e g_e p_buy h_cos h_sin season".split()
    with open(os.path.join(OUT, "lu1_cases.csv"), "w") as f:
        f.write(",".join(["case"] + names + ["idx", "a1", "a2", "track"]) + "\n")
        for i in range(N_CASES):
            f.write(",".join([str(i)] + [num(v) for v in cs["state"][i]] + [str(int(cs["idx"][i])), num(cs["a"][i, 0]), num(cs["a"][i, 1]),
                              num(cs["track"][i])]) + "\n")
    tape = np.random.default_rng(5).uniform(0, 1, (72, 2)).astype(np.float32)
    with open(os.path.join(OUT, "lu1_tape.csv"), "w") as f:
        f.write("a1,a2\n")
        for t in range(72):
            f.write(num(tape[t, 0]) + "," + num(tape[t, 1]) + "\n")
    print("written", OUT)


if __name__ == "__main__":
    main()
