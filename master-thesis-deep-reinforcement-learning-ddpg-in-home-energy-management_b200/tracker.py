"""f1/f3 of SURVEY §8: the result artefacts around the hot path, in the reference's formats.

* `write_results_csv` — the per-step 23-column tracker CSV of `write_to_results_file`
  (RL-SHEMS/src/memory_plotting_saving.jl:167-190; column order = the `results` row, shems_LU1.jl:476-478).
* `save_checkpoint` / `load_checkpoint` — the reference saves the actor only (BSON, :263-270) and can therefore not resume
  training; here the whole learner state (4 nets in Flux layout, normalisation constants) goes into one .npz.  The Julia
  shim turns the actor arrays back into a Flux Chain (same memory layout as Dense.W / Dense.b) for saveBSON.
"""
import numpy as np

from . import _lib as L

TRACE_HEADER = ["index", "c_ev", "EV_target", "EV", "Soc_ev", "rewards", "profit", "discomfort", "penalty", "PV_DE", "B_DE",
                "GR_DE", "PV_B", "PV_GR", "PV_EV", "B_EV", "GR_EV", "EX_EV", "GR_B", "B_GR", "B", "B_tar", "Soc_b"]


def write_results_csv(path, trace, instance=0):
    """trace: [T][23][N] float64 (device tensor or array) of an inference rollout -> CSV for one instance."""
    tr = trace.cpu().numpy() if hasattr(trace, "cpu") else np.asarray(trace)
    rows = tr[:, :, instance]
    with open(path, "w") as f:
        f.write(",".join(TRACE_HEADER) + "\n")
        for r in rows:
            f.write(",".join(repr(float(v)) for v in r) + "\n")
    return dict(rewards=float(rows[:, 5].sum()), profit=float(rows[:, 6].sum()), discomfort=float(rows[:, 7].sum()),
                penalty=float(rows[:, 8].sum()))  # the sums write_to_tracker_file appends (:208-210)


def save_checkpoint(path, learner, s_min=None, s_max=None, **scores):
    out = {}
    for net, name in ((L.NET_ACTOR, "actor"), (L.NET_CRITIC, "critic"), (L.NET_ACTOR_TARGET, "actor_target"),
                      (L.NET_CRITIC_TARGET, "critic_target")):
        for k in range(3):
            i, o = learner.layer_shape(net, k)
            w, b = learner.get_layer(net, k)
            out[f"{name}_W{k + 1}"] = w.reshape(i, o).T.copy()  # Flux Dense.W: out×in
            out[f"{name}_b{k + 1}"] = b
    if s_min is not None:
        out["s_min"], out["s_max"] = np.asarray(s_min, np.float32), np.asarray(s_max, np.float32)
    for k, v in scores.items():
        out[k] = np.asarray(v)
    np.savez(path, **out)


def load_checkpoint(path, learner):
    z = np.load(path)
    for net, name in ((L.NET_ACTOR, "actor"), (L.NET_CRITIC, "critic"), (L.NET_ACTOR_TARGET, "actor_target"),
                      (L.NET_CRITIC_TARGET, "critic_target")):
        for k in range(3):
            W = z[f"{name}_W{k + 1}"]
            learner.set_layer(net, k, np.ascontiguousarray(W.T).ravel(), z[f"{name}_b{k + 1}"])
    if "s_min" in z:
        learner.set_norm(z["s_min"], z["s_max"])
    return z
