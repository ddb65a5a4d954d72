"""Generates the committed fixtures under tests/golden/ (run in the build container only).

  charger98_test_series.npz  the Charger98 *test* input series (2999 h) rebuilt from the only real
                             Charger98 data the reference ships: the MPC benchmark result
                             /root/reference/SHEMS python/single_building/results/260724_results_2999_…_Charger98.csv
                             (flow balances of SHEMS_optimizer_cost.py:55-57).  float32 [8][2999].
  kat_appendix_b.json        the hand-derived known-answer vectors K1-K6 of SURVEY.md Appendix B
                             (surveyor's values, NOT reference output) next to this repo's oracle values.
  oracle_rule_based_charger98.npz
                             oracle outputs (NOT reference outputs: Julia cannot run here) of the
                             rule-based inference on that series: per-column sums of the 23-column trace
                             and the first/last rows — a regression pin for oracle and CUDA path alike.

The reference holds no golden vector for shems_LU1 (everything under RL-SHEMS/out/ is a git-LFS
pointer), so these fixtures pin regressions and the surveyor's hand derivations, not reference output.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

import shems_b200 as sb  # noqa: E402
from oracle import oracle as O  # noqa: E402

MPC = "/root/reference/SHEMS python/single_building/results/260724_results_2999_2999_0-2999_1_5_10.0_all_test_fix_Charger98.csv"


def main():
    ser = sb.series.from_mpc_results(MPC)
    np.savez_compressed(os.path.join(HERE, "charger98_test_series.npz"), series=ser)

    P = O.params_for_charger(98)
    # Appendix B: state = [Soc_b, Soc_ev, c_ev, d_e, g_e, p_buy]; surveyor's expected values
    kats = [
        dict(name="K1", state=[3.0, 0.5, 5, 1.0, 4.0, 0.4], a=[0.8, 0.9], track=0,
             survey=dict(B=-2.99991012, EV=11, Soc_b=9.0357935e-05, Soc_ev=0.80712527, reward=-2.060034382, PV_DE=1, PV_EV=3, B_EV=2.849914122, GR_EV=5.150085878)),
        dict(name="K2", state=[2.0, 1.0, -1, 0.5, 5.0, 0.4], a=[0.9, 1.0], track=0,
             survey=dict(B=3.29999995, EV=0, Soc_b=5.29984093, Soc_ev=1.0, reward=0.082105266, PV_DE=0.5, PV_B=3.29999995, PV_GR=1.026315796)),
        dict(name="K3", state=[2.0, 1.0, -1, 1.5, 0.0, 0.4], a=[0.3, 0.5], track=0,
             survey=dict(B=-1.99994004, EV=0, Soc_b=0.42104000, Soc_ev=1.0, reward=-0.050000001, B_DE=1.5, penalty=0.05),
             note="survey Soc_b' 0.42104000 evaluates (B_DE+B_EV)/eta in Float64; in Julia both addends are Float32 here, so the sum and "
                  "the division stay Float32 (shems_LU1.jl:432) -> 0.42103994 (difference 1.4e-7 relative)"),
        dict(name="K4", state=[0.0, 0.6, 0, 0.5, 0.0, 0.4], a=[0.0, 0.7], track=0,
             survey=dict(B=0, EV=3.58159900, Soc_b=0, Soc_ev=1.0, reward=-14.930560858, discomfort=30.0000019, GR_DE=0.5, GR_EV=3.581599, EX_EV=10.74480057)),
        dict(name="K5", state=[1.0, 0.4, 3, 0.8, 6.0, 0.4], a=None, track=-0.5,
             survey=dict(B=-0.99997002, EV=11, Soc_b=3.0099443e-05, Soc_ev=0.70712531, reward=-1.940011548, PV_DE=0.8, PV_EV=5.2, B_EV=0.949971393, GR_EV=4.850028798)),
        dict(name="K6", state=[0.0005, 0.2, 10, 2.0, 0.5, 0.4], a=[0.5, 1.0], track=0,
             survey=dict(B=0, EV=11, Soc_b=4.99985e-04, Soc_ev=0.50712532, reward=-5.000000075, PV_DE=0.5, GR_DE=1.5, GR_EV=11)),
    ]
    names = "index c_ev EV_target EV Soc_ev rewards profit discomfort penalty PV_DE B_DE GR_DE PV_B PV_GR PV_EV B_EV GR_EV EX_EV GR_B B_GR B B_tar Soc_b".split()
    for k in kats:
        st = np.array(k["state"] + [1, 0, 1], np.float32)
        c = k["state"][2]
        ser3 = np.zeros((8, 3), np.float32)
        ser3[0] = 1
        ser3[1] = [c, c - 1 if c > 0 else -1, -1]
        ser3[2], ser3[3], ser3[4], ser3[5], ser3[7] = 0.7, 0.1, 0.4, 1, 1
        if k["a"] is None:
            a = np.zeros(2, np.float32)
            O.lib().oracle_action_rule(P, O._fp(st), O._fp(a))
        else:
            a = np.array(k["a"], np.float32)
        r, s2, i2, tr = O.step_single(P, ser3, st, 1, a, k["track"])
        k["series3"] = ser3.tolist()
        k["action_used"] = [float(a[0]), float(a[1])]
        k["oracle"] = dict(reward=r, Soc_b=float(s2[0]), Soc_ev=float(s2[1]), state2=[float(x) for x in s2],
                           trace={n: float(v) for n, v in zip(names, tr)})
    with open(os.path.join(HERE, "kat_appendix_b.json"), "w") as f:
        json.dump(kats, f, indent=1)

    env = O.OracleEnv(P, ser, 2998, 1)
    env.reset(mode=0)
    out = env.rollout(0, 2998, want_trace=True)
    tr = out["trace"][:, :, 0]
    np.savez_compressed(os.path.join(HERE, "oracle_rule_based_charger98.npz"), colsum=tr.sum(0), first=tr[:5], last=tr[-5:],
                        ep_return=out["ep_return"], final_state=env.obs[:, 0])
    print("rule-based return on Charger98 test:", out["ep_return"][0], "profit", tr[:, 6].sum(), "(MPC upper benchmark -369.537)")


if __name__ == "__main__":
    main()
