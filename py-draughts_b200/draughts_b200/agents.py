"""The agent surface of the reference's ML helpers (draughts/__init__.py:52-55: Agent, BaseAgent, AgentEngine, Engine):
what a policy needs to be plugged where the reference expects an engine.  The search engines themselves (alpha-beta,
Hub bridge) are callers of the hot path and stay out of scope; these four classes are the interface the
reference's test_ai_features.py exercises together with copy / features / to_tensor / legal_moves_mask."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Protocol, runtime_checkable

from .move import Move


class Engine(ABC):
    """Anything that answers `get_best_move(board)`; with_evaluation=True -> (move, score)."""
    name: str = "Engine"

    @abstractmethod
    def get_best_move(self, board, with_evaluation: bool = False):
        ...


@runtime_checkable
class Agent(Protocol):
    """Structural type: any object with `select_move(board) -> Move`."""

    def select_move(self, board) -> Move:
        ...


class BaseAgent(ABC):
    """Base class for agents: a name (default: the class name) and `as_engine()`."""

    def __init__(self, name: str | None = None) -> None:
        self.name = name if name is not None else type(self).__name__

    @abstractmethod
    def select_move(self, board) -> Move:
        ...

    def as_engine(self) -> "AgentEngine":
        return AgentEngine(self)


class AgentEngine(Engine):
    """Adapter: an agent behind the Engine interface.  Agents give no evaluation, so the score is 0.0; `nodes`
    counts the positions the agent was asked about."""

    def __init__(self, agent, name: str | None = None) -> None:
        self.agent = agent
        self.name = name if name is not None else getattr(agent, "name", type(agent).__name__)
        self.nodes = 0

    @property
    def inspected_nodes(self) -> int:
        return self.nodes

    def get_best_move(self, board, with_evaluation: bool = False):
        move = self.agent.select_move(board)
        self.nodes += 1
        return (move, 0.0) if with_evaluation else move
