"""``from hugectr.data import DataSourceParams, DataSource`` (reference: hugectr.data submodule,
HugeCTR/include/pybind/data_source_wrapper.hpp:27-35)."""
from hugectr_b200.data import DataSource  # noqa: F401
from hugectr_b200.solver import DataSourceParams  # noqa: F401
