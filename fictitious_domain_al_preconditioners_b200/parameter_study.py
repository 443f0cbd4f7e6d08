"""The experimental choice of gamma of ``elliptic_interface`` ("Do parameter study",
elliptic_interface.cc:1086-1128): solve the same assembled problem once per sampled value of
gamma (background and immersed gamma set to the same value), record the outer iteration counts and keep the first
value with the fewest.  The host side (deal.II in the reference, ``synthetic`` here) rebuilds what depends on gamma —
the augmented AMG matrices ``A1 + gamma1 Ct C`` and ``A2 + gamma2 M`` — and the library gets a fresh context per
value, exactly as the reference calls ``solve()`` anew; every solve starts from a zero solution vector
(``system_solution_block = 0``, :1109)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Sequence

import numpy as np

from .context import ALContext, NoConvergence


def linspace(start: float, end: float, num: int) -> list[float]:
    """``linspace`` of utilities.h:333-346: ``start + i * (end - start) / (num - 1)``; start < end required."""
    if not start < end:
        raise ValueError("Invalid range given. Check start and stop values.")  # the reference's AssertThrow
    if num < 2:
        raise ValueError("at least two sample values")
    step = (end - start) / (num - 1)
    return [start + step * i for i in range(num)]


@dataclass
class ParameterStudy:
    gammas: list[float]
    outer_iterations: list[int]
    inner_iterations: list[int] = field(default_factory=list)
    solve_ms: list[float] = field(default_factory=list)

    @property
    def min_index(self) -> int:
        # std::min_element: the first of several minima (elliptic_interface.cc:1112-1114)
        return int(np.argmin(np.asarray(self.outer_iterations)))

    @property
    def best_gamma(self) -> float:
        return self.gammas[self.min_index]


def gamma_parameter_study(make_problem: Callable[[float], object], gammas: Sequence[float],
                          make_context: Callable[[object], ALContext] | None = None,
                          build_hierarchies: Callable[[object], dict] | None = None) -> ParameterStudy:
    """``make_problem(gamma)`` returns the assembled problem for gamma_AL_background = gamma_AL_immersed = gamma
    (``synthetic.elliptic_interface(gamma_fluid=g, gamma_solid=g, ...)``).  ``make_context(config)``: the library
    context (default: the CUDA ``ALContext``; the tests also pass the oracle's).  A solve that does not converge counts
    with the maximal step number of its outer control, as ``SolverControl::last_step()`` would report."""
    from . import synthetic as syn

    make_context = make_context or (lambda cfg: ALContext(cfg))
    build_hierarchies = build_hierarchies or syn.build_hierarchies
    study = ParameterStudy(gammas=[float(g) for g in gammas], outer_iterations=[])
    for g in study.gammas:
        prob = make_problem(g)
        H = build_hierarchies(prob) if prob.amg_matrix else {}
        ctx = make_context(prob.config)
        oracle = type(ctx).__name__ == "OracleContext"
        syn.setup_context(ctx, prob, H, oracle=oracle)
        rhs = ctx.augment_rhs(prob.rhs) if prob.augment_rhs else prob.rhs
        try:
            _, info = ctx.solve(rhs)
            study.outer_iterations.append(int(info.outer_iterations))
            study.inner_iterations.append(int(info.inner_iterations))
            study.solve_ms.append(float(info.solve_ms))
        except NoConvergence:
            study.outer_iterations.append(int(prob.config.outer.max_steps))
            study.inner_iterations.append(-1)
            study.solve_ms.append(float("nan"))
        finally:
            ctx.close()
    return study
