from .block_assembly_grasp_sim import BlockAssemblyGraspSim  # noqa: F401
from .block_assembly_orient import BlockAssemblyOrient  # noqa: F401
from .block_assembly_search import BlockAssemblySearch  # noqa: F401
from .block_assembly_insert_sim import BlockAssemblyInsertSim  # noqa: F401
from .tool_positioning import ToolPositioningChain, ToolPositioningGrasp, ToolPositioningOrient  # noqa: F401
