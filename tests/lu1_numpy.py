"""An INDEPENDENT second restatement of RL-SHEMS/RL_environments/envs/shems_LU1.jl (test infrastructure).

Unlike oracle/shems_oracle.c (C, hand-tagged Int/Float32/Float64 values) this file is a statement-for-statement Python
transliteration that leaves the typing to numpy scalars: with numpy >= 2 (NEP 50) a Python int/float literal is "weak"
exactly like a Julia Int literal next to a Float32, `np.float32 (op) np.float64 -> np.float64` is Julia's promotion rule,
and every operation rounds in the promoted type.  Julia Float64 literals (0.01, 0.001, 0.99, 0.95, 0.5) are written as
np.float64, Float32 literals (1f-3, 1f-7) as np.float32.  The two restatements share no code; tests/test_kat_leaves.py
requires them to agree BIT FOR BIT on every leaf of the flow dispatch (:362-449) and records which leaf each case took.

Line numbers in comments are shems_LU1.jl's.  Nothing here is imported by the product.
"""
import numpy as np

F32, F64 = np.float32, np.float64
assert tuple(int(x) for x in np.__version__.split(".")[:2]) >= (2, 0), "needs NEP 50 scalar promotion (numpy >= 2)"

# capacities (:47-59): (ev.soc_max::Float32, b.soc_max::Float32 = cap * 0.9f0, b.rate_max::Float64)
CAPACITIES = {
    1: (F32(48.250), F32(7.5) * F32(0.9), F64(3.3)), 2: (F32(36.271), F32(10) * F32(0.9), F64(3.3)),
    3: (F32(45.508), F32(10) * F32(0.9), F64(3.3)), 4: (F32(78.993), F32(11) * F32(0.9), F64(4.6)),
    5: (F32(37.207), F32(10) * F32(0.9), F64(4.6)), 6: (F32(35.816), F32(15) * F32(0.9), F64(4.6)),
    7: (F32(36.521), F32(12) * F32(0.9), F64(3.3)), 8: (F32(45.728), F32(10) * F32(0.9), F64(3.3)),
    9: (F32(21.935), F32(7.5) * F32(0.9), F64(3.3)), 98: (F32(35.816), F32(7.5) * F32(0.9), F64(3.3)),
    97: (F32(78.993), F32(11) * F32(0.9), F64(4.6)),
}


# shems_LU7.jl:42-55: ev_capacities::Dict{Int, Float64} of Float32 literals (ids 1-9, 98, 99)
LU7_EV_CAPACITIES = {k: F64(v[0]) for k, v in CAPACITIES.items() if k != 97}
LU7_EV_CAPACITIES[99] = F64(F32(35.816))


class Consts:
    """pv, b, ev, m and penalty_weight of the module (:40-43, :67-99) for one charger id (JOB_ID digits, :45).
    variant: "LU1" (shems_LU1.jl), "LU7" (shems_LU7.jl) or "INPUT0607" (shems_LU1_input0607.jl with fourth ternary digit 0) — the
    sibling files differ in these constants and in the reward line only (reward_line below)."""

    def __init__(self, charger_id=98, variant="LU1"):
        self.variant = variant
        self.pv_eta = F32(1)                              # PV(1f0) :92
        self.ev_soc_min, self.ev_rate_max, self.b_loss = F32(0), F32(11), F32(0.00003)
        self.b_eta, self.b_soc_min = F32(0.95), F32(0)
        if variant == "LU7":
            # Battery(0.95f0, 0f0, 10f0, 4.6f0, 0.00003f0) shems_LU7.jl:91 (rate_max::Float64 <- 4.6f0); Market(0.3f0, 1) :94
            self.b_soc_max, self.b_rate_max = F32(10), F64(F32(4.6))
            self.ev_soc_max = F32(LU7_EV_CAPACITIES[charger_id])      # ::Float32 field <- the Float64 dict value
            self.sell_discount = F64(F32(0.3))
            self.discomfort_weight_ev = F64(1)            # DISCOMFORT_WEIGHT_EV = 1 (Int) -> Float64 field
            self.disc_pot = None                          # the struct has no disc_pot: the reward line is linear (:465-468)
            self.penalty_weight = F64(0.1)                # penalty_weight = 0.1 (Float64) :35
            return
        cap = CAPACITIES[charger_id]                      # KeyError like the reference (:95)
        if variant == "INPUT0607" and charger_id == 97:
            raise KeyError(97)                            # shems_LU1_input0607.jl:57-68 has no charger 97
        self.b_soc_max, self.b_rate_max = cap[1], cap[2]  # Battery(0.95f0, 0f0, cap, rate, 0.00003f0) :95
        self.ev_soc_max = cap[0]                          # :97
        self.sell_discount = F64(F32(0.2))                # Market(0.2f0, DISCOMFORT_WEIGHT_EV, DISC_POT) with Float64 fields :85-89, :99
        if variant == "INPUT0607":
            self.discomfort_weight_ev = F64(F32(0.1))     # fourth ternary digit 0 (shems_LU1_input0607.jl:38-47)
            self.disc_pot = F64(F32(1))                   # DISC_POT = 1f0 (:49)
            self.penalty_weight = F64(0.1)                # penalty_weight = 0.1 (Float64) (:52)
        else:
            self.discomfort_weight_ev = F64(F32(0.01))
            self.disc_pot = F64(F32(2))
            self.penalty_weight = F32(0.1)                # :43

    def discomfort_term(self, discomfort):
        """the discomfort part of the reward line: shems_LU1.jl:467-470 w * discomfort^pot; shems_LU7.jl:465-468 discomfort * w;
        shems_LU1_input0607.jl:481-484 (discomfort * w)^pot.  discomfort is the Int 0 or a Float32."""
        if self.variant == "LU7":
            return discomfort * self.discomfort_weight_ev                     # Int/Float32 * Float64 -> Float64
        d = F64(0) if isinstance(discomfort, int) else F64(discomfort)        # promotion of Int / Float32 next to a Float64
        if self.variant == "INPUT0607":
            return (d * self.discomfort_weight_ev) ** self.disc_pot
        return self.discomfort_weight_ev * (d ** self.disc_pot)


def jl_min(x, y):
    """Base.min after promotion (no NaNs on this path)."""
    t = np.result_type(x, y)
    x, y = t.type(x), t.type(y)
    return y if y < x else x


def jl_clamp(x, lo, hi):
    """Base.clamp(x, lo, hi): ifelse(x > hi, hi, ifelse(x < lo, lo, x)) converted to the promoted type."""
    t = np.result_type(x, lo, hi) if not all(isinstance(v, int) for v in (x, lo, hi)) else None
    conv = (lambda v: t.type(v)) if t is not None else (lambda v: v)
    return conv(hi) if x > hi else (conv(lo) if x < lo else conv(x))


def to_f32_pair(B, EV):
    """Float32.([B, EV]) (:315, :339)"""
    return F32(B), F32(EV)


def action_drl(K, state, a, tags=None):
    """action(env, a::ShemsAction) :283-316"""
    Soc_b, Soc_ev, c_ev, d_e, g_e = state[:5]
    B_target, EV_target = a
    Soc_b_perc = (Soc_b - K.b_soc_min) / (K.b_soc_max - K.b_soc_min)               # :288
    if c_ev > -1 and Soc_ev < EV_target:                                            # :292
        EV = jl_min(K.ev_rate_max, (EV_target - Soc_ev) * (K.ev_soc_max - K.ev_soc_min))
        ev_tag = "ev_on"
    else:
        EV = 0
        ev_tag = "ev_off"
    pv_ = g_e - d_e - EV                                                            # :301
    if pv_ > 0 and Soc_b_perc < B_target:                                           # :304
        B_target_value = B_target * (K.b_soc_max - K.b_soc_min) + K.b_soc_min
        B = jl_clamp(pv_, 0, jl_min(K.b_rate_max, (B_target_value - Soc_b + K.b_loss)))
        b_tag = "b_charge"
    elif Soc_b > F32(1e-3):                                                         # :309
        B = -jl_min(K.b_rate_max, ((1 - K.b_loss) * Soc_b))
        b_tag = "b_discharge"
    else:
        B = 0
        b_tag = "b_idle"
    if tags is not None:
        tags.update(act_ev=ev_tag, act_b=b_tag)
    return to_f32_pair(B, EV)


def action_rule(K, state, tags=None):
    """action(env, track) :318-340"""
    Soc_b, Soc_ev, c_ev, d_e, g_e = state[:5]
    EV = jl_min(K.ev_rate_max, (1 - Soc_ev) * (K.ev_soc_max - K.ev_soc_min))         # :323
    pv_ = g_e - d_e - EV
    if pv_ > 0 and Soc_b < (F64(0.95) * K.b_soc_max):                                # :330
        B = jl_clamp(pv_, 0, jl_min(K.b_rate_max, K.b_soc_max - Soc_b + K.b_loss))
        b_tag = "b_charge"
    elif Soc_b > F32(1e-3):
        B = -jl_min(K.b_rate_max, ((1 - K.b_loss) * Soc_b))
        b_tag = "b_discharge"
    else:
        B = 0
        b_tag = "b_idle"
    if tags is not None:
        tags.update(act_ev="rule", act_b=b_tag)
    return to_f32_pair(B, EV)


SERIES_ORDER = ("soc_ev", "h_countdown", "electkwh", "PV_generation", "p_buy", "hour_cos", "hour_sin", "season")


def step(K, series, state, idx, a, track=0):
    """step!(env, s, a; track) :343-485 incl. next_state! :264-281.
    series: float array [8][nrows] in SERIES_ORDER (df columns); state: 9 np.float32 (env.state); idx: env.idx (1-based).
    returns (env.reward::Float64, Vector{Float32}(env.state), env.idx, results[23]::Float64, tags)"""
    tags = {}
    st = [F32(v) for v in state]
    Soc_b, Soc_ev, c_ev, d_e, g_e, p_buy, h_cos, h_sin, season = st                 # :344
    if idx + 1 > series.shape[1]:
        raise IndexError("BoundsError: row idx+1 > nrow(df) (:266-268)")
    if track >= 0:                                                                  # :346
        B_target, EV_target = F32(a[0]), F32(a[1])
        B, EV = action_drl(K, st, (B_target, EV_target), tags)
    else:                                                                           # :350
        B_target, EV_target = F32(0), F32(0)
        B, EV = F32(a[0]), F32(a[1])
        tags.update(act_ev="given", act_b="given")
    pv_, BD, BC = F64(0), F64(0), F64(0)                                            # zeros(8) :356
    PV_DE = PV_B = PV_EV = PV_GR = B_DE = B_EV = B_GR = GR_DE = GR_EV = GR_B = EX_EV = F64(0)   # zeros(11) :357
    if B < F64(-0.01):                                                              # :362
        BD = jl_clamp(-B, F64(0.001), jl_min(K.b_rate_max, ((1 - K.b_loss - F32(1e-7)) * Soc_b)))
    tags["discharge"] = bool(BD > 0)
    if (g_e * K.pv_eta) > d_e:                                                      # :368
        PV_DE = d_e
        pv_ = (g_e * K.pv_eta) - PV_DE
        if pv_ > EV:                                                                # :371
            PV_EV = EV
            pv_ = pv_ - PV_EV
            tags["flow"] = "A1"
        elif pv_ <= EV:
            PV_EV = pv_
            pv_ = 0
            if BD > (EV - PV_EV) / K.b_eta:                                         # :377
                B_EV = (EV - PV_EV)
                BD = BD - B_EV / K.b_eta
                tags["flow"] = "A2a"
            elif BD <= (EV - PV_EV) / K.b_eta:
                B_EV = BD * K.b_eta
                BD = 0
                GR_EV = (EV - PV_EV) - B_EV
                tags["flow"] = "A2b"
    elif (g_e * K.pv_eta) <= d_e:                                                   # :388
        PV_DE = g_e * K.pv_eta
        pv_ = 0
        d_e = d_e - PV_DE
        if BD > (d_e / K.b_eta):                                                    # :392
            B_DE = d_e
            BD = BD - B_DE / K.b_eta
            if BD > (EV / K.b_eta):
                B_EV = EV
                BD = BD - B_EV / K.b_eta
                tags["flow"] = "B1a"
            elif BD <= (EV / K.b_eta):
                B_EV = BD * K.b_eta
                BD = 0
                GR_EV = EV - B_EV
                tags["flow"] = "B1b"
        elif BD <= (d_e / K.b_eta):                                                 # :403
            B_DE = BD * K.b_eta
            BD = 0
            GR_DE = d_e - B_DE
            GR_EV = EV
            tags["flow"] = "B2"
    tags["charge"] = "none"
    if B > F64(0.01):                                                               # :412
        BC = jl_clamp(B, F64(0.001), jl_min(K.b_rate_max, K.b_soc_max - Soc_b))
        if pv_ > (BC / K.b_eta):
            PV_B = BC
            pv_ = pv_ - (BC / K.b_eta)
            tags["charge"] = "c1"
        elif pv_ <= (BC / K.b_eta):
            PV_B = pv_ * K.b_eta
            pv_ = 0
            GR_B = 0
            tags["charge"] = "c2"
    PV_GR = pv_                                                                     # :424
    B_GR = 0
    new_Soc_b = F32((1 - K.b_loss) * (Soc_b + PV_B + GR_B - ((B_DE + B_EV + B_GR) / K.b_eta)))   # :432 (store converts to Float32)
    new_Soc_ev = F32(Soc_ev + (PV_EV + B_EV + GR_EV) / (K.ev_soc_max - K.ev_soc_min))            # :435
    discomfort = 0
    penalty = 0
    EX_EV = 0
    tags["tail"] = "none"
    if c_ev == 0 and new_Soc_ev < 1:                                                # :442
        discomfort = (1 - new_Soc_ev) * 100
        EX_EV = (1 - new_Soc_ev) * (K.ev_soc_max - K.ev_soc_min)
        new_Soc_ev = F32(1)
        tags["tail"] = "departure"
    elif c_ev < 0 and EV_target < F64(0.99):                                        # :447
        penalty = (1 - EV_target) * K.penalty_weight
        tags["tail"] = "penalty"
    # next_state! :264-281 (df[row, :col] with 1-based rows)
    col = {n: series[k] for k, n in enumerate(SERIES_ORDER)}
    nidx = idx + 1
    new_c_ev = F32(col["h_countdown"][nidx - 1])
    tags["arrival"] = False
    if new_c_ev >= 0 and col["h_countdown"][idx - 1] == -1:                         # :270
        new_Soc_ev = F32(col["soc_ev"][nidx - 1])
        tags["arrival"] = True
    new_state = [new_Soc_b, new_Soc_ev, new_c_ev, F32(col["electkwh"][nidx - 1]), F32(col["PV_generation"][nidx - 1]),
                 F32(col["p_buy"][nidx - 1]), F32(col["hour_cos"][nidx - 1]), F32(col["hour_sin"][nidx - 1]),
                 F32(col["season"][nidx - 1])]
    # reward :464-471 (p_buy is the pre-step price)
    profit = (K.sell_discount * p_buy * (PV_GR + B_GR)) - (p_buy * (GR_DE + GR_B + GR_EV + EX_EV))
    dterm = K.discomfort_term(discomfort)                                           # 0::Int ^ 2.0 -> 0.0; Float32 ^ Float64 promotes
    if track < 0:
        reward = profit - dterm
        penalty = 0
    else:
        reward = profit - dterm - penalty
    results = [nidx, c_ev, EV_target, EV, Soc_ev, reward, profit, discomfort, penalty, PV_DE, B_DE, GR_DE,
               PV_B, PV_GR, PV_EV, B_EV, GR_EV, EX_EV, GR_B, B_GR, B, B_target, Soc_b]                    # :476-478
    return F64(reward), np.array(new_state, F32), nidx, np.array([F64(v) for v in results], F64), tags


def reset_state(K, series, maxsteps, rng_is_minus1, idx_draw=None, socb_draw=None):
    """reset_state! :216-262 with the two MersenneTwister draws (:224-225) supplied by the caller."""
    cd = series[1]
    nrow = series.shape[1]
    if rng_is_minus1:
        Soc_b = F32(F64(0.5) * (K.b_soc_min + K.b_soc_max))
        idx = 1
    else:
        Soc_b = F32(socb_draw)
        idx = int(idx_draw)
        c_ev_end = cd[idx + maxsteps - 1]
        counter = 0
        while c_ev_end > -1 and idx < (nrow - maxsteps):
            idx += int(c_ev_end + 1)
            if idx > (nrow - maxsteps):
                idx = int(idx_draw)                      # same seed -> the same draw again (:236)
            c_ev_end = cd[idx + maxsteps - 1]
            counter += 1
            if counter > 100:
                break
    row = [F32(series[k][idx - 1]) for k in range(8)]
    return np.array([Soc_b] + row, F32), idx
