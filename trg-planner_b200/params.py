"""Parameter sets of the reference's YAML configs (values only; the YAML files live in the
read-only reference tree, which does not exist on the GPU box).

config/indoor.yaml:14-21 and config/mountain.yaml:14-21; order follows TRG::TRG, trg.h:51-59.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, asdict


class CParams(ctypes.Structure):
    """Mirror of `TrgParams` (include/trg_b200.h) == `OrcParams` (oracle/trg_oracle.h)."""
    _fields_ = [
        ("is_verbose", ctypes.c_int),
        ("expand_dist", ctypes.c_float),
        ("robot_size", ctypes.c_float),
        ("sample_num", ctypes.c_int),
        ("height_threshold", ctypes.c_float),
        ("collision_threshold", ctypes.c_float),
        ("update_collision_threshold", ctypes.c_float),
        ("safety_factor", ctypes.c_float),
        ("goal_tolerance", ctypes.c_float),
    ]


@dataclass(frozen=True)
class TrgParams:
    is_verbose: bool = False
    expand_dist: float = 0.5
    robot_size: float = 0.5
    sample_num: int = 20
    height_threshold: float = 0.5
    collision_threshold: float = 0.5
    update_collision_threshold: float = 0.5
    safety_factor: float = 3.0
    goal_tolerance: float = 0.2

    def to_c(self) -> CParams:
        d = asdict(self)
        d["is_verbose"] = int(d["is_verbose"])
        return CParams(**d)


INDOOR = TrgParams(False, 0.4, 0.3, 15, 0.15, 0.1, 0.1, 3.0, 0.8)
MOUNTAIN = TrgParams(False, 0.6, 0.3, 7, 0.16, 0.1, 0.5, 3.0, 0.8)
