"""B200-native (sm_100a) two-tower training step — drop-in for the hot path of zoongahn/jodalroB-twoTower.

Import as `jodalrob_twotower_b200` (the hyphenated directory name is not a Python identifier; the
underscore package next to it forwards here).
"""
from .kjt import KeyedJaggedTensor  # noqa: F401
from .schema import (categorical_keys_from_metadata, keys_from_categorical_config, load_schema_json,  # noqa: F401
                     vocab_sizes_from_metadata)


def __getattr__(name):
    # modules/engine need the CUDA library; resolve lazily so `import jodalrob_twotower_b200` works for
    # the host-side helpers (KJT, schema) everywhere, while any compute entry fails loudly without the .so
    if name in ("CategoricalEmbedder", "BaseTower", "NoticeTower", "CompanyTower", "TwoTowerModel", "TwoTowerTrainTask",
                "create_two_tower_model", "create_two_tower_train_task", "FusedAdam", "FusedWarmupScheduler", "GraphedFusedStep"):
        from . import modules
        return getattr(modules, name)
    if name == "TwoTowerEvaluator":
        from . import evaluator
        return evaluator.TwoTowerEvaluator
    if name == "FeatureProjector":
        from . import projector
        return projector.FeatureProjector
    if name in ("DeviceFeatureStore", "GpuBatchAssembler"):
        from . import loader
        return getattr(loader, name)
    if name == "PooledTables":
        from . import tables
        return tables.PooledTables
    if name in ("TwoTowerEngine", "TwoTowerConfig"):
        from . import engine
        return getattr(engine, name)
    if name in ("native", "ops", "modules", "engine", "dist", "tables", "evaluator", "loader", "projector"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
