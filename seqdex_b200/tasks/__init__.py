from .block_assembly_grasp_sim import BlockAssemblyGraspSim  # noqa: F401
