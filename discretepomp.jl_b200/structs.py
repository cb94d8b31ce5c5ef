"""Public and private types of the path: the unchanged data contract of DiscretePOMP.jl.

Mirrors `src/hmm_structs.jl` and `src/cmn_structs.jl` of the reference:
Event (:12-15), Observation (:30-35), Particle (:51-58), SimResults (:88-93), DPOMPModel (:107-116),
HiddenMarkovModel (:119-130), MCMCSample (:147-152); RejectionSample (cmn_structs.jl:13-17),
ImportanceSample (cmn_structs.jl:34-41).  Field names, order and meaning are the reference's.
Arrays follow Julia's layout: `theta` is (n_theta, n_samples), populations are (n_particles, n_compartments).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Callable, List, Optional

import numpy as np


@dataclass(order=False)
class Event:
    time: float
    event_type: int  # 1-based, indexes the rate function and the transition matrix

    def __lt__(self, other: "Event") -> bool:  # isless(a::Event, b::Event) (src/DiscretePOMP.jl:75)
        return self.time < other.time


@dataclass
class Observation:
    time: float
    obs_id: int  # <1 if not a resampling step
    prop: float  # df: 1.0
    val: np.ndarray  # Array{Int64,1}

    def __post_init__(self):
        self.val = np.asarray(self.val, dtype=np.int64).copy()

    def __lt__(self, other: "Observation") -> bool:  # src/DiscretePOMP.jl:76
        return self.time < other.time


@dataclass
class Particle:
    theta: np.ndarray
    initial_condition: np.ndarray
    final_condition: np.ndarray
    trajectory: List[Event]
    prior: float  # log prior
    log_like: np.ndarray  # [full log like g(x), latest marginal]


@dataclass
class SimResults:
    model_name: str
    particle: Particle
    population: List[np.ndarray]
    observations: List[Observation]


@dataclass
class DPOMPModel:
    """Public model (src/hmm_structs.jl:107-116).  Mutable, like the reference's `mutable struct`."""

    model_name: str
    rate_function: Callable  # rate_function(output, parameters, population) -> None, in place
    initial_condition: np.ndarray
    m_transition: np.ndarray  # (n_events, n_compartments)
    obs_function: Callable  # simulation only
    obs_model: Callable  # obs_model(y, population, theta) -> log likelihood
    prior: Any  # object with logpdf(theta) and rand(n) -> (n_theta, n)
    t0_index: int = 0  # 1-based index of the initial-time parameter, 0 if fixed at 0.0

    def __post_init__(self):
        self.initial_condition = np.asarray(self.initial_condition, dtype=np.int64)
        self.m_transition = np.atleast_2d(np.asarray(self.m_transition, dtype=np.int64))


@dataclass
class HiddenMarkovModel:
    """Private model (src/hmm_structs.jl:119-130), built by get_private_model (src/DiscretePOMP.jl:96-99)."""

    model_name: str
    n_events: int
    rate_function: Callable
    fn_initial_condition: Callable
    fn_transition: Callable
    obs_function: Callable
    obs_model: Callable
    obs_data: List[Observation]
    prior: Any
    t0_index: int
    # device side of the model (not in the reference): the compiled rate table handle, created lazily
    _device_model: Optional[Any] = field(default=None, repr=False, compare=False)
    _public: Optional[DPOMPModel] = field(default=None, repr=False, compare=False)


@dataclass
class RejectionSample:
    theta: np.ndarray  # (n_theta, iterations, chains)
    mu: np.ndarray
    cv: np.ndarray


@dataclass
class ImportanceSample:
    mu: np.ndarray
    cv: np.ndarray
    theta: np.ndarray  # (n_theta, n)
    weight: np.ndarray
    run_time: int  # ns
    bme: np.ndarray  # -log evidence estimates


@dataclass
class MCMCSample:
    samples: RejectionSample
    adapt_period: int
    sre: np.ndarray
    run_time: int
