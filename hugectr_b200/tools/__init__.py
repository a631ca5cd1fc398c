"""hugectr.tools: DataGenerator, sharding planner, workspace calculator, format converters."""


def __getattr__(name):
    import importlib
    if name in ("DataGenerator", "DataGeneratorParams"):
        m = importlib.import_module("hugectr_b200.data.generator")
        return getattr(m, name)
    if name in ("planner", "workspace_calculator", "criteo2raw", "convert_to_raw", "embedding_gen", "criteo2parquet",
                "criteo2predict"):
        return importlib.import_module(f"hugectr_b200.tools.{name}")
    raise AttributeError(name)
