"""microrts_b200 -- host-side mirror of the reference's simulation API over libmicrorts_cuda.so.

Names follow the reference (rts.units.UnitTypeTable, rts.PhysicalGameState, rts.GameState, rts.UnitAction); the
batched state is the new class the reference would gain as rts.cuda.BatchedGameState.
"""
from .api import (ACTIONS_RAW, ACTIONS_VECTOR, BatchedGameState, MicroRTSError, PhysicalGameState, UnitAction,
                  UnitTypeTable, POLICY_EXTERNAL, POLICY_LIGHT_RUSH, POLICY_PASSIVE, POLICY_RANDOM_BIASED,
                  POLICY_WORKER_RUSH, POLICY_HEAVY_RUSH, POLICY_RANGED_RUSH, POLICY_WORKER_DEFENSE, POLICY_LIGHT_DEFENSE,
                  POLICY_HEAVY_DEFENSE, POLICY_RANGED_DEFENSE, POLICY_PO_WORKER_RUSH, POLICY_PO_LIGHT_RUSH, POLICY_PO_HEAVY_RUSH,
                  POLICY_PO_RANGED_RUSH, POLICY_WORKER_RUSH_PP, POLICY_CRUSH_V1, POLICY_CRUSH_V2, POLICY_EMR_DETERMINISTICO, PF_ASTAR, PF_BFS, PF_GREEDY, PF_FLOODFILL, DTYPE_U8, DTYPE_I32)

__all__ = ["UnitTypeTable", "PhysicalGameState", "BatchedGameState", "UnitAction", "MicroRTSError", "ACTIONS_RAW",
           "ACTIONS_VECTOR", "POLICY_EXTERNAL", "POLICY_PASSIVE", "POLICY_RANDOM_BIASED", "POLICY_WORKER_RUSH",
           "POLICY_LIGHT_RUSH", "POLICY_HEAVY_RUSH", "POLICY_RANGED_RUSH", "POLICY_WORKER_DEFENSE", "POLICY_LIGHT_DEFENSE",
           "POLICY_HEAVY_DEFENSE", "POLICY_RANGED_DEFENSE", "POLICY_PO_WORKER_RUSH", "POLICY_PO_LIGHT_RUSH", "POLICY_PO_HEAVY_RUSH",
           "POLICY_PO_RANGED_RUSH", "POLICY_WORKER_RUSH_PP", "POLICY_CRUSH_V1", "POLICY_CRUSH_V2", "POLICY_EMR_DETERMINISTICO", "PF_ASTAR", "PF_BFS", "PF_GREEDY", "PF_FLOODFILL", "DTYPE_U8", "DTYPE_I32", "rewards", "vec_client", "trace", "maps", "search"]
from . import maps, rewards, search, trace, vec_client  # noqa: E402,F401  (JNIGridnetVecClient facade + src/ai/reward functions)
