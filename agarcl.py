"""`import agarcl` — the reference's Python module name (environment/bindings.cpp:94), served by the compiled pybind11
module agarcl_b200/agarcl.<abi>.so (agarcl_b200/csrc/pybind_agarcl.cpp) over libagarcl_b200.so.  This file only makes the
module importable by its reference name from the repository root; build with `python -m agarcl_b200.build`."""
from agarcl_b200.agarcl import *  # noqa: F401,F403
from agarcl_b200.agarcl import (CloneInfo, FoodInfo, GlobalState, GoBiggerEnvironment, GridEnvironment, Location, PlayerState,  # noqa: F401
                                PlayerStates, SporeInfo, VirusInfo, has_screen_env)
