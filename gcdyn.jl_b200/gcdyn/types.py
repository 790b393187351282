"""Host-side mirror of gcdyn.jl's data types (reference: src/types.jl).

Julia is not available in the build environment, so this Python module plays the
role of the Julia wrapper package: same names, same argument meaning, same
validation rules and error behaviour.  Julia exception types map to the classes
below (all are subclasses of the closest Python builtin).
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "EVENTS", "EVENT_CODE", "TreeNode",
    "AbstractBranchingProcess", "SigmoidalBranchingProcess",
    "ConstantBranchingProcess", "DiscreteBranchingProcess",
    "ArgumentError", "DimensionMismatch", "BoundsError", "DomainError",
]


class ArgumentError(ValueError):
    """Julia `ArgumentError`."""


class DimensionMismatch(ValueError):
    """Julia `DimensionMismatch`."""


class BoundsError(IndexError):
    """Julia `BoundsError`."""


class DomainError(ValueError):
    """Julia `DomainError` (e.g. `log` of a negative real)."""


#: The possible events of a node (types.jl:6).  The position in this tuple is the
#: event code used in the flattened struct-of-arrays and by the C ABI.
EVENTS = ("root", "birth", "sampled_death", "unsampled_death", "type_change",
          "sampled_survival", "unsampled_survival")
EVENT_CODE = {name: code for code, name in enumerate(EVENTS)}


class TreeNode:
    """`TreeNode(event, time, type)` (types.jl:23-37).

    A tree is a node with event ``"root"`` plus children.  ``time`` is stored as a
    float (Julia field `time::Float64`); ``type`` keeps whatever Python type it was
    given (Julia `TreeNode{T}`); for a ``type_change`` node, ``type`` is the *new*
    type.  ``children`` is ordered; ``up`` is the parent or ``None``.
    """

    __slots__ = ("event", "time", "type", "children", "up")

    def __init__(self, event, time, type):
        if event not in EVENT_CODE:
            raise ArgumentError(f"Event must be one of {EVENTS}")
        self.event = event
        self.time = float(time)
        self.type = type
        self.children = []
        self.up = None

    def __repr__(self):  # treenode.jl:249
        return f"TreeNode: {self.event} event at time {self.time} with type {self.type}"


def _uniform_gamma(n, γ):
    # types.jl:98-100 — `ones(n, n) / (n-1) * γ` with `-γ` on the diagonal.  n == 1
    # divides by zero exactly as the reference does (Inf/NaN entries, rejected later).
    with np.errstate(divide="ignore", invalid="ignore"):
        Γ = np.ones((n, n)) / np.float64(n - 1) * γ
    Γ[np.diag_indices(n)] = -γ
    return Γ


class AbstractBranchingProcess:
    """`AbstractBranchingProcess{T}` (types.jl:44).  Immutable after construction."""

    __slots__ = ("μ", "δ", "Γ", "ρ", "σ", "type_space", "_index_cache")

    def _init_common(self, μ, δ, Γ, ρ, σ, type_space):
        # Validation order and messages follow types.jl:78-90 (= :131-143 = :183-195).
        if ρ < 0 or ρ > 1:
            raise ArgumentError("ρ must be between 0 and 1")
        elif σ < 0 or σ > 1:
            raise ArgumentError("σ must be between 0 and 1")
        Γ = np.asarray(Γ, dtype=np.float64)
        ts = np.array(list(type_space) if not isinstance(type_space, np.ndarray) else type_space,
                      dtype=np.float64).reshape(-1)
        if Γ.ndim != 2:
            raise DimensionMismatch("The rate matrix must be two-dimensional.")
        if len(ts) != Γ.shape[0]:
            raise DimensionMismatch("The number of types in the type space must match the number of rows in the rate matrix.")
        elif len(ts) != Γ.shape[1]:
            raise DimensionMismatch("The number of types in the type space must match the number of columns in the rate matrix.")
        # `any(≉(0; atol=1e-10), sum(Γ, dims=2))` — isapprox with rtol = 0; NaN is not ≈ 0.
        rowsum = Γ.sum(axis=1)
        if np.any(~(np.abs(rowsum) <= 1e-10)):
            raise ArgumentError("The transition rate matrix must contain only rows that sum to 0.")
        elif np.any(np.diag(Γ) > 0) and len(ts) > 1:
            raise ArgumentError("The transition rate matrix must contain only negative or zero values on the diagonal.")
        object.__setattr__(self, "μ", μ * 1.0)
        object.__setattr__(self, "δ", δ * 1.0)
        object.__setattr__(self, "Γ", Γ)
        object.__setattr__(self, "ρ", float(ρ))
        object.__setattr__(self, "σ", float(σ))
        object.__setattr__(self, "type_space", ts)
        object.__setattr__(self, "_index_cache", {})

    def __setattr__(self, name, value):
        raise AttributeError(f"{type(self).__name__} is immutable (Julia `struct`)")

    # ASCII aliases for the Greek field names.
    mu = property(lambda self: self.μ)
    delta = property(lambda self: self.δ)
    Gamma = property(lambda self: self.Γ)
    rho = property(lambda self: self.ρ)
    sigma = property(lambda self: self.σ)

    def birth_rates(self):
        """λ evaluated on the whole type space (what `dp_dt!` recomputes per RHS call, mbp.jl:267-268)."""
        raise NotImplementedError


class SigmoidalBranchingProcess(AbstractBranchingProcess):
    """types.jl:65-103.

    ``SigmoidalBranchingProcess(λ_xscale, λ_xshift, λ_yscale, λ_yshift, μ, γ, ρ, σ, type_space)`` or
    ``SigmoidalBranchingProcess(λ_xscale, λ_xshift, λ_yscale, λ_yshift, μ, δ, Γ, ρ, σ, type_space)``.
    """

    __slots__ = ("λ_xscale", "λ_xshift", "λ_yscale", "λ_yshift")

    def __init__(self, *args):
        if len(args) == 9:
            xs, xsh, ys, ysh, μ, γ, ρ, σ, type_space = args
            Γ = _uniform_gamma(len(type_space), γ)
            δ = 1
        elif len(args) == 10:
            xs, xsh, ys, ysh, μ, δ, Γ, ρ, σ, type_space = args
        else:
            raise TypeError("SigmoidalBranchingProcess takes 9 or 10 positional arguments")
        self._init_common(μ, δ, Γ, ρ, σ, type_space)
        for name, v in (("λ_xscale", xs), ("λ_xshift", xsh), ("λ_yscale", ys), ("λ_yshift", ysh)):
            object.__setattr__(self, name, v * 1.0)

    def birth_rates(self):
        from .multitype_branching_process import sigmoid
        return np.array([sigmoid(x, self.λ_xscale, self.λ_xshift, self.λ_yscale, self.λ_yshift)
                         for x in self.type_space], dtype=np.float64)


class ConstantBranchingProcess(AbstractBranchingProcess):
    """types.jl:121-155.  ``(λ, μ, γ, ρ, σ, type_space)`` or ``(λ, μ, δ, Γ, ρ, σ, type_space)``."""

    __slots__ = ("λ",)

    def __init__(self, *args):
        if len(args) == 6:
            λ, μ, γ, ρ, σ, type_space = args
            Γ = _uniform_gamma(len(type_space), γ)
            δ = 1
        elif len(args) == 7:
            λ, μ, δ, Γ, ρ, σ, type_space = args
        else:
            raise TypeError("ConstantBranchingProcess takes 6 or 7 positional arguments")
        self._init_common(μ, δ, Γ, ρ, σ, type_space)
        object.__setattr__(self, "λ", λ * 1.0)

    def birth_rates(self):
        return np.full(len(self.type_space), self.λ, dtype=np.float64)


class DiscreteBranchingProcess(AbstractBranchingProcess):
    """types.jl:173-207.  ``λ`` is a vector indexed like ``type_space``."""

    __slots__ = ("λ",)

    def __init__(self, *args):
        if len(args) == 6:
            λ, μ, γ, ρ, σ, type_space = args
            Γ = _uniform_gamma(len(type_space), γ)
            δ = 1
        elif len(args) == 7:
            λ, μ, δ, Γ, ρ, σ, type_space = args
        else:
            raise TypeError("DiscreteBranchingProcess takes 6 or 7 positional arguments")
        self._init_common(μ, δ, Γ, ρ, σ, type_space)
        object.__setattr__(self, "λ", np.asarray(λ, dtype=np.float64).reshape(-1).copy())

    def birth_rates(self):
        # The reference indexes `model.λ[i]` lazily (mbp.jl:38) and never checks the length.
        if len(self.λ) < len(self.type_space):
            raise BoundsError("λ has fewer entries than type_space")
        return self.λ[: len(self.type_space)].copy()
