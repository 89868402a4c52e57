"""The reference's task yamls (cfg/allegro_hand_*.yaml) as literals: the `env:` scalars and the whole `sim:` block, per task.  A task class's
DEFAULT_CFG is the entry below -- a task runs with the configuration its yaml states unless the caller passes another cfg.  Pinned field
for field to the yamls by tests/test_task_cfg_cpu.py (fixture tests/golden/task_cfg.json, written by oracle/gen_golden_cfg.py)."""

TASK_CFG = {
    "BlockAssemblyGraspSim": {   # cfg/allegro_hand_block_assembly_grasp_sim.yaml
        "env": {
            "actionPenaltyScale": -0.0, "actionsMovingAverage": 1.0, "aggregateMode": 1, "asymmetric_observations": True,
            "controlFrequencyInv": 1, "distRewardScale": -1, "dofSpeedScale": 1.0, "enableDebugVis": False,
            "enable_camera_sensors": False, "envSpacing": 1.25, "env_name": "allegro_hand_block_assembly_grasp_sim",
            "episodeLength": 150, "fallDistance": 0.4, "fallPenalty": 0.0, "forceLimitScale": 1.0,
            "handAgentIndex": "[[0, 1, 2, 3, 4, 5]]", "handResetStep": 0, "maxConsecutiveSuccesses": 0, "numEnvs": 2048,
            "objectType": "egg", "observationType": "partial_contact", "printNumSuccesses": False, "reachGoalBonus": 250,
            "resetDofPosRandomInterval": 0.0, "resetDofVelRandomInterval": 0.0, "resetPositionNoise": 0.0,
            "resetRotationNoise": 0.0, "rotEps": 0.1, "rotRewardScale": 1.0, "startPositionNoise": 0.0,
            "startRotationNoise": 0.0, "stiffnessScale": 1.0, "successTolerance": 0.1, "useRelativeControl": False,
        },
        "sim": {"substeps": 2,
                "physx": {"bounce_threshold_velocity": 0.002, "contact_collection": 0, "contact_offset": 0.002, "default_buffer_size_multiplier": 15.0, "max_depenetration_velocity": 1000.0, "num_position_iterations": 16, "num_threads": 64, "num_velocity_iterations": 0, "rest_offset": 0.0, "solver_type": 1},
                "flex": {"num_inner_iterations": 20, "num_outer_iterations": 5, "relaxation": 0.75, "warm_start": 0.8}},
        "task": {"randomize": False},
    },
    "BlockAssemblyInsertSim": {   # cfg/allegro_hand_block_assembly_insert_sim.yaml
        "env": {
            "actionPenaltyScale": -0.0, "actionsMovingAverage": 1.0, "aggregateMode": 1, "asymmetric_observations": True,
            "controlFrequencyInv": 1, "distRewardScale": -1, "dofSpeedScale": 1.0, "enableDebugVis": False,
            "enable_camera_sensors": False, "envSpacing": 1.25, "env_name": "allegro_hand_block_assembly_insert_sim",
            "episodeLength": 125, "fallDistance": 0.4, "fallPenalty": 0.0, "forceLimitScale": 1.0,
            "handAgentIndex": "[[0, 1, 2, 3, 4, 5]]", "handResetStep": 0, "maxConsecutiveSuccesses": 0, "numEnvs": 2048,
            "objectType": "egg", "observationType": "partial_contact", "printNumSuccesses": False, "reachGoalBonus": 250,
            "resetDofPosRandomInterval": 0.0, "resetDofVelRandomInterval": 0.0, "resetPositionNoise": 0.0,
            "resetRotationNoise": 0.0, "rotEps": 0.1, "rotRewardScale": 1.0, "startPositionNoise": 0.0,
            "startRotationNoise": 0.0, "stiffnessScale": 1.0, "successTolerance": 0.1, "useRelativeControl": False,
        },
        "sim": {"substeps": 2,
                "physx": {"bounce_threshold_velocity": 0.002, "contact_collection": 0, "contact_offset": 0.002, "default_buffer_size_multiplier": 15.0, "max_depenetration_velocity": 1000.0, "num_position_iterations": 16, "num_threads": 64, "num_velocity_iterations": 0, "rest_offset": 0.0, "solver_type": 1},
                "flex": {"num_inner_iterations": 20, "num_outer_iterations": 5, "relaxation": 0.75, "warm_start": 0.8}},
        "task": {"randomize": False},
    },
    "BlockAssemblyOrient": {   # cfg/allegro_hand_block_assembly_orient.yaml
        "env": {
            "actionPenaltyScale": -0.0, "actionsMovingAverage": 0.2, "aggregateMode": 1, "asymmetric_observations": True,
            "controlFrequencyInv": 1, "distRewardScale": -1, "dofSpeedScale": 1.0, "enableDebugVis": False,
            "enable_camera_sensors": False, "envSpacing": 1.25, "env_name": "allegro_hand_block_assembly_orient",
            "episodeLength": 75, "fallDistance": 0.4, "fallPenalty": 0.0, "forceLimitScale": 1.0,
            "handAgentIndex": "[[0, 1, 2, 3, 4, 5]]", "handResetStep": 0, "maxConsecutiveSuccesses": 0, "numEnvs": 2048,
            "objectType": "egg", "observationType": "partial_contact", "printNumSuccesses": False, "reachGoalBonus": 250,
            "resetDofPosRandomInterval": 0.0, "resetDofVelRandomInterval": 0.0, "resetPositionNoise": 0.0,
            "resetRotationNoise": 0.0, "rotEps": 0.1, "rotRewardScale": 1.0, "startPositionNoise": 0.0,
            "startRotationNoise": 0.0, "stiffnessScale": 1.0, "successTolerance": 0.1, "useRelativeControl": False,
        },
        "sim": {"substeps": 2,
                "physx": {"bounce_threshold_velocity": 0.002, "contact_collection": 1, "contact_offset": 0.02, "default_buffer_size_multiplier": 15.0, "max_depenetration_velocity": 1000.0, "num_position_iterations": 16, "num_threads": 64, "num_velocity_iterations": 0, "rest_offset": 0.0, "solver_type": 1},
                "flex": {"num_inner_iterations": 20, "num_outer_iterations": 5, "relaxation": 0.75, "warm_start": 0.8}},
        "task": {"randomize": False},
    },
    "BlockAssemblySearch": {   # cfg/allegro_hand_block_assembly_search.yaml
        "env": {
            "actionPenaltyScale": -0.0, "actionsMovingAverage": 0.6, "aggregateMode": 1, "asymmetric_observations": True,
            "controlFrequencyInv": 1, "distRewardScale": -1, "dofSpeedScale": 5.0, "enableDebugVis": False,
            "enable_camera_sensors": True, "envSpacing": 1.25, "env_name": "allegro_hand_block_assembly_search",
            "episodeLength": 75, "fallDistance": 0.4, "fallPenalty": 0.0, "forceLimitScale": 1.0,
            "handAgentIndex": "[[0, 1, 2, 3, 4, 5]]", "handResetStep": 45, "maxConsecutiveSuccesses": 0, "numEnvs": 1,
            "objectType": "egg", "observationType": "partial_contact", "printNumSuccesses": False, "reachGoalBonus": 250,
            "resetDofPosRandomInterval": 0.0, "resetDofVelRandomInterval": 0.0, "resetPositionNoise": 0.0,
            "resetRotationNoise": 0.0, "rotEps": 0.1, "rotRewardScale": 1.0, "startPositionNoise": 0.0,
            "startRotationNoise": 0.0, "stiffnessScale": 1.0, "successTolerance": 0.1, "useRelativeControl": False,
        },
        "sim": {"substeps": 2,
                "physx": {"bounce_threshold_velocity": 0.002, "contact_collection": 1, "contact_offset": 0.02, "default_buffer_size_multiplier": 15.0, "max_depenetration_velocity": 1000.0, "num_position_iterations": 16, "num_threads": 64, "num_velocity_iterations": 0, "rest_offset": 0.0, "solver_type": 1},
                "flex": {"num_inner_iterations": 20, "num_outer_iterations": 5, "relaxation": 0.75, "warm_start": 0.8}},
        "task": {"randomize": False},
    },
    "ToolPositioningGrasp": {   # cfg/allegro_hand_tool_positioning_grasp.yaml
        "env": {
            "actionPenaltyScale": -0.0, "actionsMovingAverage": 1.0, "aggregateMode": 1, "asymmetric_observations": True,
            "controlFrequencyInv": 1, "distRewardScale": -1, "dofSpeedScale": 1.0, "enableDebugVis": False,
            "enable_camera_sensors": False, "envSpacing": 1.25, "env_name": "allegro_hand_tool_positioning_grasp",
            "episodeLength": 150, "fallDistance": 0.4, "fallPenalty": 0.0, "forceLimitScale": 1.0,
            "handAgentIndex": "[[0, 1, 2, 3, 4, 5]]", "handResetStep": 0, "maxConsecutiveSuccesses": 0, "numEnvs": 2048,
            "objectType": "egg", "observationType": "partial_contact", "printNumSuccesses": False, "reachGoalBonus": 250,
            "resetDofPosRandomInterval": 0.0, "resetDofVelRandomInterval": 0.0, "resetPositionNoise": 0.0,
            "resetRotationNoise": 0.0, "rotEps": 0.1, "rotRewardScale": 1.0, "startPositionNoise": 0.0,
            "startRotationNoise": 0.0, "stiffnessScale": 1.0, "successTolerance": 0.1, "useRelativeControl": False,
        },
        "sim": {"substeps": 2,
                "physx": {"bounce_threshold_velocity": 0.002, "contact_collection": 0, "contact_offset": 0.002, "default_buffer_size_multiplier": 15.0, "max_depenetration_velocity": 1.0, "num_position_iterations": 16, "num_threads": 64, "num_velocity_iterations": 0, "rest_offset": 0.0, "solver_type": 1},
                "flex": {"num_inner_iterations": 20, "num_outer_iterations": 5, "relaxation": 0.75, "warm_start": 0.8}},
        "task": {"randomize": False},
    },
    "ToolPositioningOrient": {   # cfg/allegro_hand_tool_positioning_orient.yaml
        "env": {
            "actionPenaltyScale": -0.0, "actionsMovingAverage": 1.0, "aggregateMode": 1, "asymmetric_observations": True,
            "controlFrequencyInv": 1, "distRewardScale": -1, "dofSpeedScale": 1.0, "enableDebugVis": False,
            "enable_camera_sensors": False, "envSpacing": 1.25, "env_name": "allegro_hand_tool_positioning_orient",
            "episodeLength": 125, "fallDistance": 0.4, "fallPenalty": 0.0, "forceLimitScale": 1.0,
            "handAgentIndex": "[[0, 1, 2, 3, 4, 5]]", "handResetStep": 0, "maxConsecutiveSuccesses": 0, "numEnvs": 2048,
            "objectType": "egg", "observationType": "partial_contact", "printNumSuccesses": False, "reachGoalBonus": 250,
            "resetDofPosRandomInterval": 0.0, "resetDofVelRandomInterval": 0.0, "resetPositionNoise": 0.0,
            "resetRotationNoise": 0.0, "rotEps": 0.1, "rotRewardScale": 1.0, "startPositionNoise": 0.0,
            "startRotationNoise": 0.0, "stiffnessScale": 1.0, "successTolerance": 0.1, "useRelativeControl": False,
        },
        "sim": {"substeps": 2,
                "physx": {"bounce_threshold_velocity": 0.002, "contact_collection": 0, "contact_offset": 0.002, "default_buffer_size_multiplier": 15.0, "max_depenetration_velocity": 1.0, "num_position_iterations": 16, "num_threads": 64, "num_velocity_iterations": 0, "rest_offset": 0.0, "solver_type": 1},
                "flex": {"num_inner_iterations": 20, "num_outer_iterations": 5, "relaxation": 0.75, "warm_start": 0.8}},
        "task": {"randomize": False},
    },
}


def default_cfg(task):
    """a deep copy of the task's yaml-stated configuration"""
    import copy
    return copy.deepcopy(TASK_CFG[task])


def scene_from_cfg(task, cfg=None, seed=22, **overrides):
    """the Scene a task runs on, from its cfg dict (missing keys = the yaml's values): the ONE place yaml keys become solver
    parameters -- task classes, bench.py and chain.py all come through here.  dt is Isaac Gym's 1/60 (utils/config.py CFG:188)."""
    from ..scene import Scene
    d = TASK_CFG[task]
    cfg = cfg or d
    env, sim = cfg.get("env", {}), cfg.get("sim", {})
    physx, dphysx = sim.get("physx", {}), d["sim"]["physx"]
    kw = dict(task=task, seed=seed, dt=1.0 / 60.0, substeps=int(sim.get("substeps", d["sim"]["substeps"])),
              iters=int(physx.get("num_position_iterations", dphysx["num_position_iterations"])),
              contact_offset=float(physx.get("contact_offset", dphysx["contact_offset"])),
              max_depen_vel=float(physx.get("max_depenetration_velocity", dphysx["max_depenetration_velocity"])),
              episode_length=int(env.get("episodeLength", d["env"]["episodeLength"])),
              act_moving_average=float(env.get("actionsMovingAverage", d["env"]["actionsMovingAverage"])))
    kw.update(overrides)
    return Scene(**kw)
