"""``from hugectr.tools import DataGenerator, DataGeneratorParams`` (reference: hugectr.tools submodule,
HugeCTR/include/pybind/data_generator_wrapper.hpp:29-70)."""
from hugectr_b200.data.generator import DataGenerator, DataGeneratorParams  # noqa: F401
from hugectr_b200.tools import criteo2raw, planner, workspace_calculator  # noqa: F401
