"""draughts_b200 -- B200-native batched draughts move generation behind the py-draughts API.

Drop-in names of the reference (draughts/__init__.py:36-66): Board (= StandardBoard), the eight
variant boards, Move, Color, Figure; plus the batched surface (BoardBatch, DeviceBoardBatch,
perft, rollout drivers).  Importing this package loads nothing from CUDA; the first compute call
binds the device and fails loudly if the sm_100a library or a GPU is missing.
"""
from .agents import Agent, AgentEngine, BaseAgent, Engine
from .batch import BoardBatch, DeviceBoardBatch, random_positions
from .boards import (BOARDS, AmericanBoard, AntidraughtsBoard, BaseBoard, BoardFeatures, BrazilianBoard,
                     BreakthroughBoard, FrisianBoard, FryskBoard, RussianBoard, StandardBoard, get_board)
from .models import Color, Figure
from .move import Move
from .drivers import perft, perft_divide, rollout

Board = StandardBoard
__version__ = "0.1.0"

__all__ = ["Board", "BaseBoard", "StandardBoard", "AmericanBoard", "FrisianBoard", "RussianBoard", "BrazilianBoard",
           "AntidraughtsBoard", "BreakthroughBoard", "FryskBoard", "BoardFeatures", "Move", "Color", "Figure",
           "Agent", "BaseAgent", "AgentEngine", "Engine", "BoardBatch", "DeviceBoardBatch", "random_positions", "perft", "perft_divide", "rollout", "get_board",
           "BOARDS"]
