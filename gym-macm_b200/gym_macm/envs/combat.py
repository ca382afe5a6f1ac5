"""`TDM`: team death-match host (gym_macm/envs/combat.py:56-264) -- placeholder until the TDM
kernel lands in this round; constructing it raises instead of silently doing something else."""


class TDM(object):
    name = "Team Deathmatch"

    def __init__(self, *args, **kwargs):
        raise NotImplementedError("TDM is not built yet")
