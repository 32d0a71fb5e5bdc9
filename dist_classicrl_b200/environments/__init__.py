from dist_classicrl_b200.environments.custom_env import DeviceVecEnv, DistClassicRLEnv
from dist_classicrl_b200.environments.hash_mdp import HashMDPVecEnv
from dist_classicrl_b200.environments.rigged_two_armed_bandit import RiggedTwoArmedBanditVecEnv, make_bandit_vec_env
from dist_classicrl_b200.environments.tiktaktoe_mod import TicTacToeVecEnv

__all__ = ["DeviceVecEnv", "DistClassicRLEnv", "HashMDPVecEnv", "RiggedTwoArmedBanditVecEnv", "TicTacToeVecEnv",
           "make_bandit_vec_env"]
