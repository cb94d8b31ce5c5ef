"""Simulation entry point -- host-side mirror of gillespie_sim (src/DiscretePOMP.jl:134-152, src/hmm_sim.jl:55-102).

The Doob-Gillespie trajectories are simulated on the device by the trajectory kernels of the MBP layer
(dpomp_mbp_iterate: one thread per trajectory, events recorded in HBM), `n_sims` trajectories per call; the host only
rebuilds the reference's result structs (SimResults: Particle with its event list, the state after every event, the
observations filled in by the model's obs_function).
"""
from __future__ import annotations

import os
from decimal import Decimal
from typing import List, Union

import numpy as np

from .mbp_ibis import MbpParticles
from .particle_filter import device_model, get_private_model
from .structs import DPOMPModel, Event, HiddenMarkovModel, Observation, Particle, SimResults

C_DEFAULT_OBS_PROP = 1.0  # src/hmm_sim.jl:74


def generate_observations(tmax: float, num_obs: int, n_states: int) -> List[Observation]:
    """generate_observations (src/hmm_sim.jl:75-82): equally spaced blank observations, obs_id = 1."""
    step = tmax / num_obs
    return [Observation(step * (i + 1), 1, C_DEFAULT_OBS_PROP, np.zeros(n_states, dtype=np.int64)) for i in range(num_obs)]


def _results(model: HiddenMarkovModel, theta: np.ndarray, ptcls: MbpParticles, p: int, y: List[Observation],
             states: np.ndarray) -> SimResults:
    fc, times, types, ll = ptcls.get_particle(p)
    ic = np.asarray(model.fn_initial_condition(), dtype=np.int64)
    traj = [Event(float(t), int(e)) for t, e in zip(times, types)]
    pop_v, x = [], ic.copy()
    for e in types:  # push!(pop_v, copy(p.final_condition)) after every event (src/hmm_sim.jl:66)
        x = x + model.fn_transition(int(e))
        pop_v.append(x.copy())
    obs = [Observation(o.time, o.obs_id, o.prop, o.val.copy()) for o in y]
    for i, o in enumerate(obs):  # observe && model.obs_function(y[i], p.final_condition, theta) (src/hmm_sim.jl:98)
        model.obs_function(o, states[i].copy(), theta)
    prior = model.prior.logpdf(theta) if hasattr(model.prior, "logpdf") else 0.0
    part = Particle(theta.copy(), ic.copy(), fc.copy(), traj, float(prior), ll.copy())
    return SimResults(model.model_name, part, pop_v, obs)


def gillespie_sim(model: DPOMPModel, parameters, tmax: float = 100.0, num_obs: int = 5, n_sims: int = 1, seed: int = 1,
                  max_traj: int = 196000, verbose: bool = True) -> Union[SimResults, List[SimResults]]:
    """gillespie_sim(model, parameters; tmax = 100.0, num_obs = 5, n_sims = 1) (src/DiscretePOMP.jl:134-152).
    Returns a SimResults, or a list of them when n_sims > 1.  `max_traj` is the reference's MAX_TRAJ event capacity."""
    theta = np.asarray(parameters, dtype=np.float64)
    y = generate_observations(tmax, num_obs, len(model.initial_condition))
    mdl = get_private_model(model, y)
    if verbose:
        print(f"Running: {model.model_name} DGA for θ := {theta.tolist()}" + (f" x {n_sims}" if n_sims > 1 else ""), end="")
    ptcls = MbpParticles(device_model(mdl), n_sims, max_traj, seed)  # the store grows on demand up to max_traj
    th = np.tile(theta[:, None], (1, n_sims))
    c = len(model.initial_condition)
    states = np.zeros((n_sims, num_obs, c), dtype=np.int64)
    for i in range(1, num_obs + 1):
        ptcls.iterate(th, i, fresh=(i == 1))
        states[:, i - 1, :] = ptcls.final_conditions()
    out = [_results(mdl, theta, ptcls, p + 1, y, states[p]) for p in range(n_sims)]
    if verbose:
        print(" - finished.")
    return out[0] if n_sims == 1 else out


def _jl_float(x: float) -> str:
    """Julia's `string(::Float64)`: shortest round-trip digits, positional for 1e-4 <= |x| < 1e6, else d.ddde±n."""
    x = float(x)
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Inf" if x > 0 else "-Inf"
    if x == 0.0:
        return "-0.0" if str(x).startswith("-") else "0.0"
    sign, digits, exp = Decimal(repr(x)).as_tuple()  # repr: shortest digits that round-trip, like Ryu
    digits = "".join(map(str, digits)).rstrip("0") or "0"
    e10 = len("".join(map(str, Decimal(repr(x)).as_tuple().digits))) + exp - 1  # decimal exponent of the first digit
    neg = "-" if sign else ""
    if -4 <= e10 < 6:
        if e10 >= 0:
            whole = digits[: e10 + 1].ljust(e10 + 1, "0")
            frac = digits[e10 + 1:] or "0"
            return f"{neg}{whole}.{frac}"
        return f"{neg}0.{'0' * (-e10 - 1)}{digits}"
    return f"{neg}{digits[0]}.{digits[1:] or '0'}e{e10}"


def save_to_file(results: SimResults, dpath: str, literal: bool = False) -> None:
    """save_to_file(results::SimResults, dpath) (src/hmm_utils.jl:35-72): writes `<dpath>sim.csv` (one row per event:
    time, event type, state after the event) and `<dpath>obs.csv` (time, id, observed values); like the reference, `dpath`
    is used as a prefix (`string(dpath, "sim.csv")`), so it normally ends with a path separator.
    The reference's sim.csv header is `time, event,1` and every row carries the whole state as ONE Julia array literal
    (`population` is a Vector of Vectors, so `size(results.population, 2) == 1`): `literal=True` reproduces those bytes;
    the default writes one column per compartment, which is what the header suggests and what a CSV reader can parse."""
    if dpath and not os.path.isdir(dpath):
        os.makedirs(dpath, exist_ok=True)
    traj = results.particle.trajectory
    n_comp = len(results.population[0]) if len(results.population) else len(results.particle.initial_condition)
    with open(f"{dpath}sim.csv", "w") as f:
        f.write("time, event")
        f.write(",1" if literal else "".join(f",{p + 1}" for p in range(n_comp)))
        for i, ev in enumerate(traj):
            f.write(f"\n{_jl_float(ev.time)},{int(ev.event_type)}")
            x = results.population[i]
            f.write(",[" + ", ".join(str(int(v)) for v in x) + "]" if literal else "".join(f",{int(v)}" for v in x))
    with open(f"{dpath}obs.csv", "w") as f:
        f.write("time,id")
        f.write("".join(f",{p + 1}" for p in range(len(results.observations[0].val))))
        for o in results.observations:
            f.write(f"\n{_jl_float(o.time)},{int(o.obs_id)}" + "".join(f",{int(v)}" for v in o.val))
