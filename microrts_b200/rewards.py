"""Reward functions of the reference (src/ai/reward/*.java) over a whole batch.

Each class keeps the reference's name and constants.  The reference computes a reward from the TraceEntry of the step (the
PlayerActions as issueSafe left them + the PhysicalGameState before the cycle) and the GameState after the cycle; here the
step kernel reduces those to a few integers per player and game (mrts_batch_set_info_output, include/microrts_cuda.h) and
`compute` turns them into the reference's (reward, done) pair for every game at once, in float64 like the reference.

  info    : int32 [n][12]  step facts of the maximising player (layout in include/microrts_cuda.h)
  results : int32 [n][4]   time, winner (-1 none), gameover, error bits after the cycle
"""
import numpy as np


class RewardFunctionInterface:
    # what `done` depends on, for the in-kernel auto-reset of JNIGridnetVecClient (mrts_batch_set_vec_autoreset): 1 = the game is
    # over, 2 = no Resource unit holds resources, 3 = never done; None (a custom class) makes the client reset from the host
    DONE_MODE = None

    def compute(self, info, results, maxplayer):
        raise NotImplementedError


class WinLossRewardFunction(RewardFunctionInterface):
    """src/ai/reward/WinLossRewardFunction.java:24-32 (a draw at game over counts as a loss for both players)."""

    DONE_MODE = 1

    def compute(self, info, results, maxplayer):
        over = (results[:, 2] & 1) != 0
        mp = np.broadcast_to(np.asarray(maxplayer), over.shape)
        reward = np.where(over, np.where(results[:, 1] == mp, 1.0, -1.0), 0.0)
        return reward, over


class ResourceGatherRewardFunction(RewardFunctionInterface):
    """src/ai/reward/ResourceGatherRewardFunction.java:25-46: +1 per HARVEST and per RETURN issued; done when no Resource
    unit holds resources any more."""
    RESOURCE_RETURN_REWARD = 1.0
    RESOURCE_HARVEST_REWARD = 1.0
    DONE_MODE = 2

    def compute(self, info, results, maxplayer):
        reward = info[:, 0] * self.RESOURCE_HARVEST_REWARD + info[:, 1] * self.RESOURCE_RETURN_REWARD
        return reward.astype(np.float64), info[:, 10] == 0


class AttackRewardFunction(RewardFunctionInterface):
    """src/ai/reward/AttackRewardFunction.java:23-40: +1 per attack on a cell held by the opponent, -1 on an own unit."""
    ATTACK_REWARD = 1.0
    DONE_MODE = 3

    def compute(self, info, results, maxplayer):
        return (info[:, 2] - info[:, 3]) * np.float64(self.ATTACK_REWARD), np.zeros(len(info), dtype=bool)


class ProduceWorkerRewardFunction(RewardFunctionInterface):
    """src/ai/reward/ProduceWorkerRewardFunction.java:23-33."""
    WORKER_PRODUCE_REWARD = 1.0
    DONE_MODE = 3

    def compute(self, info, results, maxplayer):
        return info[:, 4] * np.float64(self.WORKER_PRODUCE_REWARD), np.zeros(len(info), dtype=bool)


class ProduceBuildingRewardFunction(RewardFunctionInterface):
    """src/ai/reward/ProduceBuildingRewardFunction.java:23-33 (Barracks or Base)."""
    BUILDING_PRODUCE_REWARD = 1.0
    DONE_MODE = 3

    def compute(self, info, results, maxplayer):
        return info[:, 5] * np.float64(self.BUILDING_PRODUCE_REWARD), np.zeros(len(info), dtype=bool)


class ProduceCombatUnitRewardFunction(RewardFunctionInterface):
    """src/ai/reward/ProduceCombatUnitRewardFunction.java:23-33 (Light, Heavy or Ranged)."""
    COMBAT_UNITS_PRODUCE_REWARD = 1.0
    DONE_MODE = 3

    def compute(self, info, results, maxplayer):
        return info[:, 6] * np.float64(self.COMBAT_UNITS_PRODUCE_REWARD), np.zeros(len(info), dtype=bool)


class CloserToEnemyBaseRewardFunction(RewardFunctionInterface):
    """src/ai/reward/CloserToEnemyBaseRewardFunction.java:23-71: old - new Euclidean distance from the opponent's first Base to
    the player's closest Worker/Light/Heavy/Ranged (2000000000 when there is none; 0 when the opponent has no Base)."""
    NONE = 2000000000.0
    DONE_MODE = 3

    def compute(self, info, results, maxplayer):
        old = np.where(info[:, 8] >= 0, np.sqrt(np.maximum(info[:, 8], 0).astype(np.float64)), self.NONE)
        new = np.where(info[:, 9] >= 0, np.sqrt(np.maximum(info[:, 9], 0).astype(np.float64)), self.NONE)
        return np.where(info[:, 7] != 0, old - new, 0.0), np.zeros(len(info), dtype=bool)


class CloserToEnemyUnitRewardFunction(CloserToEnemyBaseRewardFunction):
    """src/ai/reward/CloserToEnemyUnitRewardFunction.java is a verbatim copy of CloserToEnemyBaseRewardFunction in the reference."""
