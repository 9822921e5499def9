"""game_engine_b200 — batched referee/phase-step simulator for B200 (sm_100a).

Host side of the drop-in for the reference's BotBehaviorNode/PhaseNode/RefereeNode step
(reference agent/game_agent_v2.py:468-617, 619-803, 987-1241).  The compute path is the CUDA
library `csrc/` behind the C-ABI of `include/game_engine_b200.h`; there is no CPU fallback.
"""
from .compiler import compile_game, CompiledGame, DSLCompileError  # noqa: F401

__version__ = "0.1.0"
