from gym_macm.envs.mvmnt import Flock
from gym_macm.envs.combat import TDM
