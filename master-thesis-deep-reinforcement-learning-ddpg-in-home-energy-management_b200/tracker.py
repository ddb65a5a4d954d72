"""f1/f3 of SURVEY §8: the result artefacts around the hot path, in the reference's formats.

* `write_results_csv` — the per-step 23-column tracker CSV of `write_to_results_file`
  (RL-SHEMS/src/memory_plotting_saving.jl:167-190; column order = the `results` row, shems_LU1.jl:476-478).
* `save_checkpoint` / `load_checkpoint` — the reference saves the actor only (BSON, :263-270) and can therefore not resume
  training; here the whole state goes into one .npz: the 4 nets in Flux layout, both optimisers (ADAM moments, β powers, update
  counter), normalisation constants, OU noise state, the replay memories — for every learner of a population handle.  The
  Julia shim turns the actor arrays back into a Flux Chain (same memory layout as Dense.W / Dense.b) for saveBSON.
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


NETS = ((L.NET_ACTOR, "actor"), (L.NET_CRITIC, "critic"), (L.NET_ACTOR_TARGET, "actor_target"), (L.NET_CRITIC_TARGET, "critic_target"))


def save_checkpoint(path, learner, s_min=None, s_max=None, memories=None, **scores):
    """The WHOLE learner state in one .npz, so a run can be resumed where it stopped (the reference cannot: saveBSON keeps the
    actor and the score arrays only, memory_plotting_saving.jl:263-270): for every learner l of the handle the four nets
    (Flux layout, `l{l}_actor_W1` ... — learner 0 also without the prefix, which is what the Julia shim turns into a Flux
    Chain), ADAM moments, β powers, update counter and normalisation constants (`l{l}_state`, `l{l}_opt`: ddpg_get_state),
    OUNoise.X when OU episodes ran, and — with `memories` (one Replay per learner) — the replay memories in logical order."""
    out = {}
    P = learner.population
    keep = getattr(learner, "_sel", 0)
    for l in range(P):
        if P > 1:
            learner.select(l)
        st, opt = learner.get_state()
        out[f"l{l}_state"], out[f"l{l}_opt"] = st, opt
        for net, name in NETS:
            for k in range(3):
                i, o = learner.layer_shape(net, k)
                w, b = learner.get_layer(net, k)
                out[f"l{l}_{name}_W{k + 1}"] = w.reshape(i, o).T.copy()  # Flux Dense.W: out×in
                out[f"l{l}_{name}_b{k + 1}"] = b
                if l == 0:
                    out[f"{name}_W{k + 1}"], out[f"{name}_b{k + 1}"] = out[f"l0_{name}_W{k + 1}"], b
    if P > 1:
        learner.select(keep)
    ou = learner.get_ou_state()
    if ou is not None:
        out["ou_x"] = ou
    out["population"] = np.int64(P)
    if memories is not None:
        mems = list(memories) if isinstance(memories, (list, tuple)) else [memories]
        for l, m in enumerate(mems):
            S, A, R, S2, D = m.get()
            out[f"mem{l}_s"], out[f"mem{l}_a"], out[f"mem{l}_r"], out[f"mem{l}_s2"], out[f"mem{l}_done"] = S, A, R, S2, D
            out[f"mem{l}_capacity"] = np.int64(m.capacity)
    if s_min is not None:
        out["s_min"], out["s_max"] = np.asarray(s_min, np.float32), np.asarray(s_max, np.float32)
    for k, v in scores.items():
        out[k] = np.asarray(v)
    np.savez(path, **out)


def load_checkpoint(path, learner, memories=None):
    """Restores what save_checkpoint wrote: every learner's full state (a resumed run continues bit-identically,
    tests/test_resume_gpu.py) and, into the given EMPTY Replay objects, the replay memories.  Files holding only net
    arrays (weights-only checkpoints) restore the nets of learner 0."""
    import torch
    z = np.load(path)
    P = learner.population
    if "l0_state" in z:
        assert int(z["population"]) == P, "checkpoint holds %d learners, the handle %d" % (int(z["population"]), P)
        keep = getattr(learner, "_sel", 0)
        for l in range(P):
            if P > 1:
                learner.select(l)
            learner.set_state(z[f"l{l}_state"], z[f"l{l}_opt"])
        if P > 1:
            learner.select(keep)
        if "ou_x" in z:
            learner.set_ou_state(z["ou_x"])
    else:
        for net, name in NETS:
            for k in range(3):
                W = z[f"{name}_W{k + 1}"]
                learner.set_layer(net, k, np.ascontiguousarray(W.T).ravel(), z[f"{name}_b{k + 1}"])
        if "s_min" in z:
            learner.set_norm(z["s_min"], z["s_max"])
    if memories is not None:
        mems = list(memories) if isinstance(memories, (list, tuple)) else [memories]
        for l, m in enumerate(mems):
            assert len(m) == 0 and m.capacity == int(z[f"mem{l}_capacity"]), "restore into an empty memory of the saved capacity"
            dev = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=m._dev)
            if z[f"mem{l}_r"].size:
                m.push(dev(z[f"mem{l}_s"]), dev(z[f"mem{l}_a"]), dev(z[f"mem{l}_r"]), dev(z[f"mem{l}_s2"]), dev(z[f"mem{l}_done"]))
    return z
